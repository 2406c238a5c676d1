// pm_build.cpp -- the product's own index builder: `.panman` -> LiteIndex seed deltas.
//
// Reference: IndexBuilder (/root/reference/src/index_single_mode.cpp:736-1392 buildIndexHelper / buildIndex, :1647-2205 processNode) walks
// the tree once, recomputes the syncmers of every node inside the ranges its mutations touch and records, per node, which seeds (k-min-mers
// of l consecutive syncmers) appeared or disappeared relative to the parent: (hash, parentCount, childCount) triples in DFS order.  What
// that incremental machine computes is pinned by the reference's own test (src/test/test_index.cpp:200-230): replaying a node's deltas
// from the root gives exactly the seed multiset of seeding the node's ungapped genome directly.
//
// Here the definition IS the algorithm, and the seeding is the data-parallel part, so it runs on the GPU with the kernels of the read path:
//   1. pm_panman.cpp: one depth-first walk hands out every node's ungapped genome (host, sequential, ~1 us per kb);
//   2. batches of genomes -> seed lists on the device (one lane per genome in the syncmer kernels, one warp per genome for the k-min-mers:
//      the same launchSeedList the read path's list utilities use, sequences of tens of kilobases instead of 150 bases);
//   3. per node: sort the seed list (host threads, one node each), merge against the parent's sorted list -> the node's deltas, sorted by
//      hash.  Nodes are met in pre-order, so a stack of the sorted lists on the current root path is all that is kept.
// The result is delta-for-delta the LiteIndex the reference builds with --flank-mask 0 (tests/test_index_build.py: every node of rsv_4K,
// extended_mammoth and sars_20000).  With the reference's default --flank-mask 250 its index is NOT a function of the node genomes: masked
// positions are neither added nor deleted while the mask bounds move from node to node (index_single_mode.cpp:1770-1780, 1850-1925), so a
// node keeps seeds that an ancestor happened to have (measured on rsv_4K: leaves with 8,137 indexed seeds where the genome has 4,504).
// A genome-defined builder cannot and should not reproduce that history; flank_mask > 0 is refused here with that explanation, and the
// reference-built default index stays readable through pm_host_index_read.
#include "pm_internal.h"

#include <algorithm>
#include <thread>

using namespace pm;
using namespace pm::host;

extern "C" int seedListImpl(int device, const char* seqs, const uint64_t* off, uint64_t n, const pm_seed_params* sp, int trimStart, int trimEnd,
                            int mode, uint64_t* outHash, uint8_t* outRev, int64_t* outPos, uint64_t* outCount);

namespace {

// seeding::hpcCompress (seeding.cpp:286-306): runs of the same letter (case-insensitive) collapse to their first character
void hpcCollapse(std::string& s) {
    size_t o = 0;
    for (size_t i = 0; i < s.size(); ++i)
        if (i == 0 || std::toupper((unsigned char)s[i]) != std::toupper((unsigned char)s[i - 1])) s[o++] = s[i];
    s.resize(o);
}

struct Batch {
    std::vector<uint32_t> node;
    std::string bases;
    std::vector<uint64_t> off{0};
    void clear() { node.clear(); bases.clear(); off.assign(1, 0); }
};

// child list vs parent list (both sorted, duplicates = multiplicity) -> deltas of the child, by ascending hash
void diffSorted(const std::vector<uint64_t>& par, const std::vector<uint64_t>& chi, HostIndex& H) {
    size_t i = 0, j = 0;
    auto clamp16 = [](size_t c) { return (int16_t)std::min<size_t>(c, 32767); };
    while (i < par.size() || j < chi.size()) {
        const uint64_t h = (j >= chi.size() || (i < par.size() && par[i] < chi[j])) ? par[i] : chi[j];
        size_t pc = 0, cc = 0;
        while (i < par.size() && par[i] == h) { ++i; ++pc; }
        while (j < chi.size() && chi[j] == h) { ++j; ++cc; }
        if (pc != cc) { H.hash.push_back(h); H.parentCount.push_back(clamp16(pc)); H.childCount.push_back(clamp16(cc)); }
    }
}

}  // namespace

extern "C" {

/* every node's ungapped genome, concatenated in pre-order (== the index's DFS order) */
int pm_panman_genomes(const char* panman_path, char** bases, uint64_t** offsets, uint32_t** parent_index, char** ids_joined, uint32_t** coords, uint64_t* n_nodes) {
    if (!panman_path || !bases || !offsets || !n_nodes) return fail(PM_ERR_INVALID, "null argument");
    *bases = nullptr; *offsets = nullptr; *n_nodes = 0;
    if (parent_index) *parent_index = nullptr;
    if (ids_joined) *ids_joined = nullptr;
    if (coords) *coords = nullptr;
    return guarded([&]() -> int {
        PanmanTree T;
        readPanman(panman_path, T);
        std::string all; std::vector<uint64_t> off(1, 0);
        std::vector<uint32_t> order;
        std::vector<uint32_t> allCoords;
        walkPanmanGenomesCoords(T, coords != nullptr, [&](uint32_t v, const std::string& g, const std::vector<uint32_t>& c) {
            order.push_back(v); all += g; off.push_back(all.size());
            if (coords) allCoords.insert(allCoords.end(), c.begin(), c.end());
        });
        for (size_t i = 0; i < order.size(); ++i) if (order[i] != i) throw std::runtime_error("panman: walk left the newick pre-order");
        *bases = static_cast<char*>(std::malloc(all.size() + 1)); *offsets = static_cast<uint64_t*>(std::malloc(off.size() * sizeof(uint64_t)));
        if (!*bases || !*offsets) throw std::bad_alloc();
        std::memcpy(*bases, all.data(), all.size()); (*bases)[all.size()] = 0;
        std::memcpy(*offsets, off.data(), off.size() * sizeof(uint64_t));
        if (coords) {
            *coords = static_cast<uint32_t*>(std::malloc((allCoords.size() + 1) * sizeof(uint32_t)));
            if (!*coords) throw std::bad_alloc();
            std::memcpy(*coords, allCoords.data(), allCoords.size() * sizeof(uint32_t));
        }
        if (parent_index) {
            *parent_index = static_cast<uint32_t*>(std::malloc(T.nodes.size() * sizeof(uint32_t)));
            if (!*parent_index) throw std::bad_alloc();
            for (size_t i = 0; i < T.nodes.size(); ++i) (*parent_index)[i] = T.nodes[i].parent == kNoNode ? 0u : T.nodes[i].parent;
        }
        if (ids_joined) {
            std::string ids;
            for (const PanmanNode& n : T.nodes) { ids += n.id; ids.push_back('\n'); }
            *ids_joined = static_cast<char*>(std::malloc(ids.size() + 1));
            if (!*ids_joined) throw std::bad_alloc();
            std::memcpy(*ids_joined, ids.data(), ids.size()); (*ids_joined)[ids.size()] = 0;
        }
        *n_nodes = T.nodes.size();
        return PM_OK;
    });
}

int pm_index_build(const char* panman_path, const pm_seed_params* sp, int flank_mask, int device, pm_host_index** out) {
    if (!panman_path || !sp || !out) return fail(PM_ERR_INVALID, "null argument");
    *out = nullptr;
    if (deviceCountNoThrow() <= device || device < 0) return fail(PM_ERR_NO_DEVICE, "no usable CUDA device (the builder seeds the genomes on the GPU; there is no CPU fallback)");
    return guarded([&]() -> int {
        if (flank_mask != 0)
            throw Unsupported("flank_mask > 0: the reference's flank masking makes its index depend on the traversal history (masked positions are neither "
                              "added nor deleted while the mask moves with every node's gap map), which a genome-defined builder does not reproduce; build with "
                              "flank_mask = 0 (== panmap --flank-mask 0, delta for delta) or read a reference-built .idx");
        if (sp->l < 0 || sp->l > 64) throw std::runtime_error("unsupported l");
        PanmanTree T;
        readPanman(panman_path, T);
        const size_t N = T.nodes.size();
        std::unique_ptr<pm_host_index> hi(new pm_host_index());
        HostIndex& H = hi->h;
        H.sp = *sp;
        H.nodeOffsets.assign(N + 1, 0); H.parentIndex.assign(N, 0); H.nodeIds.resize(N); H.identicalToParent.assign(N, 0);
        for (size_t i = 0; i < N; ++i) { H.parentIndex[i] = T.nodes[i].parent == kNoNode ? 0u : T.nodes[i].parent; H.nodeIds[i] = T.nodes[i].id; }

        // pre-order stack of (node, its sorted seed list): the parent of the next node is always on it
        std::vector<std::pair<uint32_t, std::vector<uint64_t>>> path;
        const std::vector<uint64_t> none;
        uint32_t nextNode = 0;
        std::vector<uint64_t> hashOut, countOut;
        std::vector<std::vector<uint64_t>> lists;
        const unsigned nThreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        auto flush = [&](Batch& b) {
            if (b.node.empty()) return;
            const uint64_t n = b.node.size();
            uint64_t win = 0;
            for (uint64_t i = 0; i < n; ++i) { const uint64_t L = b.off[i + 1] - b.off[i]; if (L >= (uint64_t)sp->k) win += L - (uint64_t)sp->k + 1; }
            hashOut.resize(win + 1); countOut.resize(n + 1);
            const int rc = seedListImpl(device, b.bases.data(), b.off.data(), n, sp, 0, 0, 2, hashOut.data(), nullptr, nullptr, countOut.data());
            if (rc != PM_OK) throw std::runtime_error(std::string("seeding the genomes failed: ") + pm_last_error());
            // sorted list per node: independent, one host thread per slice of the batch
            lists.assign(n, {});
            std::vector<uint64_t> wOff(n + 1, 0);
            for (uint64_t i = 0; i < n; ++i) { const uint64_t L = b.off[i + 1] - b.off[i]; wOff[i + 1] = wOff[i] + (L >= (uint64_t)sp->k ? L - (uint64_t)sp->k + 1 : 0); }
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < nThreads; ++t)
                pool.emplace_back([&, t]() {
                    for (uint64_t i = t; i < n; i += nThreads) {
                        lists[i].assign(hashOut.begin() + (size_t)wOff[i], hashOut.begin() + (size_t)(wOff[i] + countOut[i]));
                        std::sort(lists[i].begin(), lists[i].end());
                    }
                });
            for (std::thread& th : pool) th.join();
            for (uint64_t i = 0; i < n; ++i) {
                const uint32_t v = b.node[i];
                if (v != nextNode) throw std::runtime_error("builder: nodes left the pre-order");
                const uint32_t p = T.nodes[v].parent;
                while (!path.empty() && path.back().first != p) path.pop_back();
                if (p != kNoNode && path.empty()) throw std::runtime_error("builder: parent of a node is not on the current root path");
                H.nodeOffsets[v] = H.hash.size();
                diffSorted(p == kNoNode ? none : path.back().second, lists[i], H);
                H.identicalToParent[v] = H.hash.size() == H.nodeOffsets[v] ? 1 : 0;
                path.emplace_back(v, std::move(lists[i]));
                ++nextNode;
            }
            b.clear();
        };
        Batch batch;
        constexpr uint64_t kBatchBases = 96ull << 20;   // ~0.8 GB of seed-list output per batch
        walkPanmanGenomes(T, [&](uint32_t v, const std::string& g) {
            if (!batch.node.empty() && batch.bases.size() + g.size() > kBatchBases) flush(batch);
            batch.node.push_back(v);
            if (sp->hpc) { std::string c = g; hpcCollapse(c); batch.bases += c; } else batch.bases += g;
            batch.off.push_back(batch.bases.size());
        });
        flush(batch);
        if (nextNode != N) throw std::runtime_error("builder: not every node was visited");
        H.nodeOffsets[N] = H.hash.size();
        *out = hi.release();
        return PM_OK;
    });
}

}  // extern "C"
