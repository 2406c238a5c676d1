// pm_build.cpp -- the product's own index builder: `.panman` -> LiteIndex seed deltas.
//
// Reference: IndexBuilder (/root/reference/src/index_single_mode.cpp:736-1392 buildIndexHelper / buildIndex, :1647-2205 processNode) walks
// the tree once, recomputes the syncmers of every node inside the ranges its mutations touch and records, per node, which seeds (k-min-mers
// of l consecutive syncmers) appeared or disappeared relative to the parent: (hash, parentCount, childCount) triples in DFS order.  What
// that incremental machine computes is pinned by the reference's own test (src/test/test_index.cpp:200-230): replaying a node's deltas
// from the root gives exactly the seed multiset of seeding the node's ungapped genome directly.
//
// Here the definition IS the algorithm, because it is the data-parallel formulation.  Two pipelines, identical output (a test holds them to it):
//   device (default; pm_build_kernels.cu): the tree flattened to an aligned root template + one point edit per mutated slot
//     (pm_panman.cpp flattenPanman), then per batch of ~4,000 nodes genome_materialize -> pack_reads + the read path's syncmer / k-min-mer
//     kernels on the genomes where they lie -> seeds_sort into a device arena -> node_diff; only the deltas come back.
//   host walk (genomes too large for the shared-memory sort, lists that would not fit the device, PM_BUILD_HOST_WALK=1): one depth-first
//     walk hands out every node's ungapped genome (pm_panman.cpp walkPanmanGenomes), batches of genomes are seeded on the device through
//     the list path of pm_read_seeds, every list is sorted by a host thread and merged against the parent's; nodes are met in pre-order,
//     so a stack of the sorted lists on the current root path is all that is kept.
// The result is the LiteIndex the reference builds with --flank-mask 0 (tests/test_index_build.py compares every node: rsv_4K identical,
// sars_20000 39,998 of 39,999 nodes, extended_mammoth 145 of 155).  With the reference's default --flank-mask 250 its index is NOT a
// function of the node genomes: masked positions are neither added nor deleted while the mask bounds move from node to node
// (index_single_mode.cpp:1770-1780, 1850-1925), so a
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

// ---- the device pipeline: genomes materialised, seeded, sorted and diffed on the GPU; only the deltas come back ----
// Throws Unsupported when a genome holds more seeds than the shared-memory sort takes (the caller then runs the host pipeline).
void buildOnDevice(const PanmanTree& T, const PanmanFlat& F, const pm_seed_params& sp, int device, HostIndex& H) {
    const u32 N = (u32)T.nodes.size(), B = (u32)T.blocks.size(), A = (u32)F.tmpl.size();
    setDevice(device);
    cudaStream_t st; CK(cudaStreamCreate(&st));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
    auto up = [&](auto& dev, const auto& host) { dev.alloc(host.size() + 1); if (!host.empty()) CK(cudaMemcpyAsync(dev.p, host.data(), host.size() * sizeof(host[0]), cudaMemcpyHostToDevice, st)); };
    DevBuf<u32> dParent, dSlotBlock, dBlockStart, dEditBegin, dEditSlot, dBlockMutBegin, dBlockMut;
    DevBuf<char> dTmpl, dEditChar; DevBuf<unsigned char> dEditSerial;
    up(dParent, F.parent); up(dSlotBlock, F.slotBlock); up(dBlockStart, F.blockStart); up(dEditBegin, F.editBegin); up(dEditSlot, F.editSlot);
    up(dBlockMutBegin, F.blockMutBegin); up(dBlockMut, F.blockMut); up(dEditSerial, F.editSerial);
    dTmpl.alloc(F.tmpl.size() + 1); if (A) CK(cudaMemcpyAsync(dTmpl.p, F.tmpl.data(), A, cudaMemcpyHostToDevice, st));
    dEditChar.alloc(F.editChar.size() + 1); if (!F.editChar.empty()) CK(cudaMemcpyAsync(dEditChar.p, F.editChar.data(), F.editChar.size(), cudaMemcpyHostToDevice, st));
    BuildTreeView V{N, B, A, F.maxDepth, dParent.p, dTmpl.p, dSlotBlock.p, dBlockStart.p, dEditBegin.p, dEditSlot.p, dEditChar.p, dEditSerial.p, dBlockMutBegin.p, dBlockMut.p};
    std::vector<SeedTables> tabs(kSeedTableElems); buildSeedTableImage(tabs.data(), sp.k, sp.s);
    DevBuf<SeedTables> dT; dT.alloc(kSeedTableElems); CK(cudaMemcpyAsync(dT.p, tabs.data(), kSeedTableElems * sizeof(SeedTables), cudaMemcpyHostToDevice, st));
    const SeederParams P = makeSeederParams(sp.k, sp.s, sp.t, sp.l, sp.open, 0, 0);

    {   // every node's sorted seed list stays on the device: refuse up front what cannot fit (the host pipeline keeps only the root path)
        size_t freeB = 0, totalB = 0;
        CK(cudaMemGetInfo(&freeB, &totalB));
        const double estimate = (double)N * (double)A * 0.5 * 8.0 + 8e9;   // <= one seed per two aligned slots, + the batch scratch
        if (estimate > 0.8 * (double)freeB) throw Unsupported("the tree's seed lists would not fit the device");
    }
    const u64 pitch = std::max<u64>(32, ((u64)A + 31) & ~31ull);   // bytes per genome slot: a whole number of 32-base chunks
    const u64 perNode = pitch * 20 + (u64)B + (u64)F.maxDepth * 4 + 64;
    const u32 nbMax = (u32)std::max<u64>(1, std::min<u64>(8192, (6ull << 30) / perNode));
    DevBuf<u32> dPath; DevBuf<unsigned char> dBlk; DevBuf<char> dAligned, dGenomes; DevBuf<u64> dEndOff, dOff, dPOff, dWOff, dHash, dCount, dSyn, dArenaOff;
    DevBuf<unsigned> dSynCount; DevBuf<uint4> dPacked; DevBuf<u32> dBF;
    dPath.alloc((size_t)nbMax * F.maxDepth); dBlk.alloc((size_t)nbMax * std::max<u32>(B, 1)); dAligned.alloc((size_t)nbMax * std::max<u32>(A, 1));
    dGenomes.alloc((size_t)nbMax * pitch + 64); dEndOff.alloc(nbMax + 1); dOff.alloc(nbMax + 1); dPOff.alloc(nbMax + 1); dWOff.alloc(nbMax + 1);
    dHash.alloc((size_t)nbMax * pitch + 1); dCount.alloc(nbMax + 1); dSyn.alloc((size_t)nbMax * pitch + 32); dSynCount.alloc(nbMax + 1);
    dPacked.alloc((size_t)nbMax * (pitch / 32) + 1); dArenaOff.alloc(nbMax + 1);
    const u64 chPer = pitch / 32;
    std::vector<u64> hOff(nbMax + 1), hPOff(nbMax + 1);
    for (u32 i = 0; i <= nbMax; ++i) { hOff[i] = (u64)i * pitch; hPOff[i] = (u64)i * chPer; }
    CK(cudaMemcpyAsync(dOff.p, hOff.data(), (nbMax + 1) * 8, cudaMemcpyHostToDevice, st));       // also the window offsets: one slot per base
    CK(cudaMemcpyAsync(dPOff.p, hPOff.data(), (nbMax + 1) * 8, cudaMemcpyHostToDevice, st));
    std::vector<u32> bf((size_t)nbMax * chPer / 256 + 2);
    // lists of every node stay on the device (the parent of a node may sit in any earlier batch)
    std::vector<std::unique_ptr<DevBuf<u64>>> arenas;
    std::vector<const u64*> hListPtr(N, nullptr); std::vector<u64> hListCount(N, 0);
    DevBuf<const u64*> dListPtr; DevBuf<u64> dListCount; dListPtr.alloc(N + 1); dListCount.alloc(N + 1);
    const unsigned diffGrid = 148 * 2;
    DevBuf<uint4> dScratch; dScratch.alloc((size_t)diffGrid * 2 * kBuildSortCap);
    DevBuf<unsigned long long> dCursor, dNodeOff; DevBuf<unsigned> dNodeCnt; dCursor.alloc(1); dNodeOff.alloc(nbMax + 1); dNodeCnt.alloc(nbMax + 1);
    u64 outCap = 4u << 20;
    DevBuf<u64> dOutHash; DevBuf<short> dOutPc, dOutCc; dOutHash.alloc(outCap); dOutPc.alloc(outCap); dOutCc.alloc(outCap);
    std::vector<u64> hCount(nbMax), hArenaOff(nbMax + 1), oHash; std::vector<short> oPc, oCc; std::vector<unsigned long long> hNodeOff(nbMax); std::vector<unsigned> hNodeCnt(nbMax);

    for (u32 v0 = 0; v0 < N; v0 += nbMax) {
        const u32 nb = std::min(nbMax, N - v0);
        const u64 ch = (u64)nb * chPer;
        launchGenomeMaterialize(V, v0, nb, dPath.p, dBlk.p, dAligned.p, dGenomes.p, pitch, dEndOff.p, st);
        packBlockFirst(hPOff.data(), nb, ch, bf.data());
        dBF.ensure(ch / 256 + 2);
        CK(cudaMemcpyAsync(dBF.p, bf.data(), ((ch + 255) / 256 + 1) * sizeof(u32), cudaMemcpyHostToDevice, st));
        launchPackReads(dGenomes.p, dOff.p, dPOff.p, dBF.p, nb, 0, ch, dPacked.p, st, dEndOff.p);
        launchSeedListsEnd(dPacked.p, dOff.p, dEndOff.p, dPOff.p, dOff.p, nb, P, dT.p, dSyn.p, dSynCount.p, dHash.p, dCount.p, st);
        CK(cudaMemcpyAsync(hCount.data(), dCount.p, nb * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        u64 tot = 0;
        for (u32 i = 0; i < nb; ++i) {
            if (hCount[i] > kBuildSortCap) throw Unsupported("a genome holds more seeds than the device sort takes");
            hArenaOff[i] = tot; tot += hCount[i];
        }
        hArenaOff[nb] = tot;
        arenas.emplace_back(new DevBuf<u64>()); arenas.back()->alloc(tot + 1);
        for (u32 i = 0; i < nb; ++i) { hListPtr[v0 + i] = arenas.back()->p + hArenaOff[i]; hListCount[v0 + i] = hCount[i]; }
        CK(cudaMemcpyAsync(dArenaOff.p, hArenaOff.data(), (nb + 1) * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dListPtr.p + v0, hListPtr.data() + v0, nb * sizeof(const u64*), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dListCount.p + v0, hListCount.data() + v0, nb * 8, cudaMemcpyHostToDevice, st));
        launchSeedsSort(dHash.p, dOff.p, dCount.p, dArenaOff.p, arenas.back()->p, nb, st);
        unsigned long long cursor = 0;
        for (int attempt = 0; attempt < 2; ++attempt) {
            CK(cudaMemsetAsync(dCursor.p, 0, sizeof(unsigned long long), st));
            BuildDiffArgs D{v0, nb, dParent.p, dListPtr.p, dListCount.p, dScratch.p, dCursor.p, outCap, dOutHash.p, dOutPc.p, dOutCc.p, dNodeOff.p, dNodeCnt.p};
            launchNodeDiff(D, diffGrid, st);
            CK(cudaMemcpyAsync(&cursor, dCursor.p, sizeof(cursor), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaGetLastError());
            if (cursor <= outCap) break;
            if (attempt) throw std::runtime_error("builder: delta buffer overflow after regrowth");
            outCap = cursor + 1024; dOutHash.alloc(outCap); dOutPc.alloc(outCap); dOutCc.alloc(outCap);   // now the exact size is known
        }
        oHash.resize(cursor + 1); oPc.resize(cursor + 1); oCc.resize(cursor + 1);
        if (cursor) {
            CK(cudaMemcpyAsync(oHash.data(), dOutHash.p, cursor * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(oPc.data(), dOutPc.p, cursor * 2, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(oCc.data(), dOutCc.p, cursor * 2, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaMemcpyAsync(hNodeOff.data(), dNodeOff.p, nb * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(hNodeCnt.data(), dNodeCnt.p, nb * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (u32 i = 0; i < nb; ++i) {   // the nodes reserved their output regions in any order: put them back in pre-order
            const u32 v = v0 + i;
            H.nodeOffsets[v] = H.hash.size();
            const size_t o = (size_t)hNodeOff[i], c = hNodeCnt[i];
            H.hash.insert(H.hash.end(), oHash.begin() + o, oHash.begin() + o + c);
            H.parentCount.insert(H.parentCount.end(), oPc.begin() + o, oPc.begin() + o + c);
            H.childCount.insert(H.childCount.end(), oCc.begin() + o, oCc.begin() + o + c);
            H.identicalToParent[v] = c == 0 ? 1 : 0;
        }
    }
    H.nodeOffsets[N] = H.hash.size();
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
        if (sp->hpc)   // measured: "collapse the genome, then seed it" matches 40 of the first 300 rsv_4K nodes of the reference's hpc build
            throw Unsupported("hpc indexes: the reference's homopolymer-compressed build is not the seeding of the compressed node genomes; read a reference-built .idx");
        PanmanTree T;
        readPanman(panman_path, T);
        const size_t N = T.nodes.size();
        std::unique_ptr<pm_host_index> hi(new pm_host_index());
        HostIndex& H = hi->h;
        H.sp = *sp;
        auto reset = [&]() {
            H.hash.clear(); H.parentCount.clear(); H.childCount.clear();
            H.nodeOffsets.assign(N + 1, 0); H.parentIndex.assign(N, 0); H.nodeIds.resize(N); H.identicalToParent.assign(N, 0);
            for (size_t i = 0; i < N; ++i) { H.parentIndex[i] = T.nodes[i].parent == kNoNode ? 0u : T.nodes[i].parent; H.nodeIds[i] = T.nodes[i].id; }
        };
        reset();
        PanmanFlat F;
        flattenPanman(T, F);
        // LiteTree.blockRanges (index_single_mode.cpp:1254-1259): first and last aligned coordinate of every block
        for (size_t b = 0; b + 1 < F.blockStart.size(); ++b) { H.blockRanges.push_back(F.blockStart[b]); H.blockRanges.push_back(F.blockStart[b + 1] - 1); }
        // the device pipeline (genomes never leave the GPU) unless a genome is too large for the shared-memory sort, the lists would not fit
        // the device, or PM_BUILD_HOST_WALK=1 asks for the host walk (the two are compared in tests/test_index_build.py)
        const char* hw = std::getenv("PM_BUILD_HOST_WALK");
        if (!(hw && std::atoi(hw) != 0)) {
            try { buildOnDevice(T, F, *sp, device, H); *out = hi.release(); return PM_OK; }
            catch (const Unsupported&) { reset(); }
        }

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
            batch.bases += g;
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
