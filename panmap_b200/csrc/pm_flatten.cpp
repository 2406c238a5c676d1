// pm_flatten.cpp -- one-time, sample-independent preparation of the seed-delta index for the GPU (host side).
//
// Input: the reference's flat view of a LiteIndex (pm_index_desc; placement.cpp:1021-1092).  Output (FlatIndex):
//   * dense seed ids in first-appearance order along the DFS + per-node deltas re-sorted by seed id
//     (12 B/delta -> 4 B/delta, and the per-delta hash probe becomes an array gather with locality: a node's
//     lost seeds are mostly ancestral low ids, its new seeds are a contiguous fresh id range)
//   * NodeMetrics::genomeMagnitudeSquared / genomeUniqueSeedCount per node, accumulated in exactly the reference's
//     order (parent's value, then the node's deltas in stored order; placement.cpp:289-302,772-774) -> bit-identical
//   * DFS subtree ends, depth, reference BFS ranks, leaf flags
//   * ancestor chains + carry slots of the K2 tiles (the in-tile Euler-tour difference uses the subtree ends)
//   * delta streams: one packed 32-bit word per 0<->1 delta in 512-word chunks with in-band segment ends, a side list for
//     the rare deltas with a genome count >= 2 and the DFS-interval events of their tree prefix
#include "pm_host.h"
#include "pm_logic.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <unordered_map>

namespace pm {

static constexpr uint32_t kTileNodesK2H = kTileNodesK2;
static constexpr uint32_t NONE = 0xFFFFFFFFu;

// for every block of 256 packed chunks counted from chunk gBase (= packedOff[0] of the slice): the slice-local read r with
// packedOff[r] <= firstChunk < packedOff[r+1]
void packBlockFirst(const uint64_t* packedOff, uint64_t nReads, uint64_t nChunks, uint32_t* out) {
    uint64_t r = 0;
    const uint64_t gBase = nReads ? packedOff[0] : 0;
    const uint64_t nBlocks = (nChunks + 255) / 256;
    for (uint64_t b = 0; b < nBlocks; ++b) {
        const uint64_t g = gBase + b * 256;
        while (r + 1 < nReads && packedOff[r + 1] <= g) ++r;
        out[b] = static_cast<uint32_t>(r);
    }
    out[nBlocks] = static_cast<uint32_t>(nReads ? nReads - 1 : 0);
}

void bfsRanks(const uint32_t* parent, uint64_t N, std::vector<uint32_t>& rank) {
    rank.assign(N, 0);
    if (N == 0) return;
    std::vector<uint32_t> cOff(N + 1, 0), kids(N), order(N);
    for (uint64_t v = 1; v < N; ++v) cOff[parent[v] + 1]++;
    for (uint64_t v = 0; v < N; ++v) cOff[v + 1] += cOff[v];
    std::vector<uint32_t> fill(cOff.begin(), cOff.end() - 1);
    for (uint64_t v = 1; v < N; ++v) kids[fill[parent[v]]++] = static_cast<uint32_t>(v);
    uint64_t head = 0, tail = 0;
    order[tail++] = 0;
    while (head < tail) {
        const uint32_t v = order[head++];
        for (uint32_t e = cOff[v]; e < cOff[v + 1]; ++e) order[tail++] = kids[e];
    }
    if (tail != N) throw std::runtime_error("index tree is not connected to node 0");
    for (uint64_t r = 0; r < N; ++r) rank[order[r]] = static_cast<uint32_t>(r);
}

void flattenIndex(const pm_index_desc& d, uint32_t shard, uint32_t nShards, FlatIndex& F) {
    const uint64_t N = d.n_nodes, D = d.n_deltas;
    if (N == 0) throw std::runtime_error("index has no nodes");
    if (N >= 0xFFFFFFFFull) throw std::runtime_error("index has too many nodes");
    if (!d.node_offsets || !d.parent_index || (D && (!d.delta_hash || !d.delta_parent || !d.delta_child)))
        throw std::runtime_error("index descriptor has null arrays");
    if (d.node_offsets[N] != D) throw std::runtime_error("nodeChangeOffsets[N] does not equal the number of seed changes");
    if (d.seed.k < 1 || d.seed.k > kMaxK || d.seed.s < 1 || d.seed.s > d.seed.k || d.seed.t < 0 || d.seed.t > d.seed.k - d.seed.s)
        throw std::runtime_error("unsupported seeding parameters (need 1 <= s <= k <= 32, 0 <= t <= k-s)");
    if (nShards == 0 || shard >= nShards) throw std::runtime_error("bad shard index");
    F.N = N; F.D = D; F.sp = d.seed;

    // ---- tree ----
    F.parent.assign(d.parent_index, d.parent_index + N);
    F.parent[0] = NONE;
    F.depth.assign(N, 0); F.subEnd.assign(N, 0); F.isLeaf.assign(N, 1);
    for (uint64_t v = 1; v < N; ++v) {
        if (F.parent[v] >= v) throw std::runtime_error("nodes are not in DFS pre-order (parentIndex >= index)");
        if (d.node_offsets[v] < d.node_offsets[v - 1]) throw std::runtime_error("nodeChangeOffsets not monotone");
        F.depth[v] = F.depth[F.parent[v]] + 1;
        F.isLeaf[F.parent[v]] = 0;
    }
    if (d.node_offsets[N] < d.node_offsets[N - 1]) throw std::runtime_error("nodeChangeOffsets not monotone");
    {   // pre-order validity + subtree ends in one sweep: keep the current root->v path on a stack
        std::vector<uint32_t> path;
        path.push_back(0);
        for (uint64_t v = 1; v < N; ++v) {
            const uint32_t p = F.parent[v];
            while (!path.empty() && path.back() != p) { F.subEnd[path.back()] = static_cast<uint32_t>(v); path.pop_back(); }
            if (path.empty()) throw std::runtime_error("tree is not in DFS pre-order");
            path.push_back(static_cast<uint32_t>(v));
        }
        for (uint32_t u : path) F.subEnd[u] = static_cast<uint32_t>(N);
    }
    bfsRanks(F.parent.data(), N, F.bfsRank);

    // ---- genome-only accumulators in the reference's order ----
    F.gMagSq.assign(N, 0.0); F.gMag.assign(N, 0.0); F.gUnique.assign(N, 0);
    {
        std::vector<double> l1p(32768);
        for (int c = 0; c < 32768; ++c) l1p[c] = std::log1p(static_cast<double>(c));
        // parents precede children in DFS order, so a forward sweep sees the parent's final value first
        for (uint64_t v = 0; v < N; ++v) {
            double mag = v ? F.gMagSq[F.parent[v]] : 0.0;
            int64_t uq = v ? F.gUnique[F.parent[v]] : 0;
            for (uint64_t i = d.node_offsets[v]; i < d.node_offsets[v + 1]; ++i) {
                const int p = d.delta_parent[i], c = d.delta_child[i];
                const double lc = c > 0 ? l1p[c] : 0.0, lp = p > 0 ? l1p[p] : 0.0;
                const double a = lc * lc, b = lp * lp;
                mag += a - b;
                uq += (c > 0) - (p > 0);
            }
            F.gMagSq[v] = mag; F.gUnique[v] = uq; F.gMag[v] = std::sqrt(mag);
        }
    }

    // ---- dictionary: dense ids in first-appearance order ----
    std::vector<uint32_t> idAll(D);
    {
        std::unordered_map<uint64_t, uint32_t> dict;
        dict.reserve(static_cast<size_t>(D / 2 + 16));
        for (uint64_t i = 0; i < D; ++i) {
            auto it = dict.find(d.delta_hash[i]);
            if (it == dict.end()) {
                const uint32_t id = static_cast<uint32_t>(F.dictHash.size());
                dict.emplace(d.delta_hash[i], id);
                F.dictHash.push_back(d.delta_hash[i]);
                idAll[i] = id;
            } else idAll[i] = it->second;
        }
    }
    F.S = F.dictHash.size();
    {
        uint64_t cap = 16;
        while (cap < 2 * F.S + 2) cap <<= 1;
        F.dictMask = cap - 1;
        F.dictKeys.assign(cap, kEmptyKey); F.dictVals.assign(cap, NONE);
        for (uint64_t id = 0; id < F.S; ++id) {
            const uint64_t h = F.dictHash[id];
            if (h == kEmptyKey) continue;  // cannot be stored; such a seed can never match (2^-64 event)
            uint64_t s = mixKey(h) & F.dictMask;
            while (F.dictKeys[s] != kEmptyKey) s = (s + 1) & F.dictMask;
            F.dictKeys[s] = h; F.dictVals[s] = static_cast<uint32_t>(id);
        }
    }

    // ---- shard extent: contiguous DFS ranges balanced by delta count ----
    auto cut = [&](uint32_t g) -> uint32_t {
        if (g == 0) return 0;
        if (g >= nShards) return static_cast<uint32_t>(N);
        const uint64_t target = (D / nShards) * g + std::min<uint64_t>(g, D % nShards);
        const uint64_t* lo = std::lower_bound(d.node_offsets, d.node_offsets + N, target);
        uint64_t v = static_cast<uint64_t>(lo - d.node_offsets);
        const uint64_t even = (N * g) / nShards;  // degenerate trees with very few deltas: fall back to node balance
        if (D < 16ull * nShards) v = even;
        return static_cast<uint32_t>(std::min<uint64_t>(v, N));
    };
    F.nodeBegin = cut(shard); F.nodeEnd = cut(shard + 1);
    if (F.nodeEnd < F.nodeBegin) F.nodeEnd = F.nodeBegin;

    // ---- local nodes: ancestors of nodeBegin (root first) ++ shard nodes ----
    std::vector<uint32_t> anc;
    if (F.nodeBegin < F.nodeEnd) {
        for (uint32_t a = F.parent[F.nodeBegin]; a != NONE; a = F.parent[a]) anc.push_back(a);
        std::reverse(anc.begin(), anc.end());
    } else {
        anc.push_back(0);  // an empty shard still needs the root for the weighted-containment denominator
    }
    F.nAnc = static_cast<uint32_t>(anc.size());
    F.nLocal = F.nAnc + (F.nodeEnd - F.nodeBegin);
    F.lNode.resize(F.nLocal);
    for (uint32_t i = 0; i < F.nAnc; ++i) F.lNode[i] = anc[i];
    for (uint32_t v = F.nodeBegin; v < F.nodeEnd; ++v) F.lNode[F.nAnc + (v - F.nodeBegin)] = v;
    if (F.S >= (1ull << 31) - 1) throw std::runtime_error("index has too many distinct seeds (2^31 limit of the packed delta word)");
    {
        uint64_t nLocalDeltas = 0;
        for (uint32_t i = 0; i < F.nLocal; ++i) nLocalDeltas += d.node_offsets[F.lNode[i] + 1] - d.node_offsets[F.lNode[i]];
        F.nLocalDeltas = nLocalDeltas;
    }
    // ---- delta streams: one packed word per fast delta (sorted by seed id inside a node), side list for the rest ----
    F.nodeSeg.assign(N, NONE);
    F.dw.clear(); F.dw.reserve(static_cast<size_t>(F.nLocalDeltas + 512));
    F.nSeg = 0; F.nGenNodes = 0;
    std::vector<uint32_t> genNodes;   // global ids of the nodes that own general deltas, ascending
    std::vector<uint64_t> segEnd;     // position (in dw) of the last word of every segment
    {
        std::vector<uint32_t> ids;
        for (uint32_t i = 0; i < F.nLocal; ++i) {
            const uint32_t v = F.lNode[i];
            const uint64_t b = d.node_offsets[v], e = d.node_offsets[v + 1];
            ids.clear();
            bool hasGen = false;
            for (uint64_t j = b; j < e; ++j) {
                const int p = d.delta_parent[j], c = d.delta_child[j];
                if (p == 0 && c == 1) ids.push_back(2 * idAll[j]);
                else if (p == 1 && c == 0) ids.push_back(2 * idAll[j] + 1);
                else if (p != c) {
                    if (!hasGen) { hasGen = true; genNodes.push_back(v); }
                    F.genSlot.push_back(static_cast<uint32_t>(genNodes.size() - 1)); F.genId.push_back(idAll[j]);
                    F.genPc.push_back(static_cast<uint32_t>(static_cast<uint16_t>(d.delta_parent[j])) |
                                      (static_cast<uint32_t>(static_cast<uint16_t>(d.delta_child[j])) << 16));
                }
            }
            if (ids.empty()) continue;
            std::sort(ids.begin(), ids.end());
            F.dw.insert(F.dw.end(), ids.begin(), ids.end());
            segEnd.push_back(F.dw.size() - 1);
            F.nodeSeg[v] = F.nSeg++;
        }
    }
    F.nFast = F.dw.size();
    F.nGenNodes = static_cast<uint32_t>(genNodes.size());
    {
        const uint64_t CH = 512;
        F.nDeltaChunks = (F.nFast + CH - 1) / CH;
        F.dw.resize(F.nDeltaChunks * CH, static_cast<uint32_t>(2 * F.S));   // padding gathers the always-zero slot, never ends a segment
        {   // lane-interleaved storage: logical word 16 l + 4 q + r of a chunk -> stored word 4 (32 q + l) + r
            std::vector<uint32_t> st(F.dw.size());
            for (uint64_t c = 0; c < F.nDeltaChunks; ++c)
                for (uint32_t l = 0; l < 32; ++l)
                    for (uint32_t q = 0; q < 4; ++q)
                        for (uint32_t r = 0; r < 4; ++r) st[c * CH + 4 * (32 * q + l) + r] = F.dw[c * CH + 16 * l + 4 * q + r];
            F.dw.swap(st);
        }
        F.endMask.assign(F.nDeltaChunks * 32, 0);
        for (uint64_t e : segEnd) F.endMask[e >> 4] |= 1u << (e & 15);
        F.chunkSeg.assign(F.nDeltaChunks + 1, 0);
        uint32_t seg = 0;
        for (uint64_t c = 0; c < F.nDeltaChunks; ++c) {
            const bool inside = c > 0 && !(F.endMask[c * 32 - 1] >> 15);   // the previous word did not end its segment
            F.chunkSeg[c] = seg | (inside ? 0x80000000u : 0u);
            if (inside) F.boundarySegs.push_back(seg);
            for (uint64_t j = c * 32; j < (c + 1) * 32; ++j) seg += static_cast<uint32_t>(__builtin_popcount(F.endMask[j]));
        }
        F.chunkSeg[F.nDeltaChunks] = seg;
        if (seg != F.nSeg) throw std::runtime_error("internal: segment count mismatch");
        F.boundarySegs.erase(std::unique(F.boundarySegs.begin(), F.boundarySegs.end()), F.boundarySegs.end());
    }
    // general deltas: A_gen[w] = sum over nodes u with general deltas and u <= w < subEnd[u]  ->  +u at position u, -u at subEnd[u]
    if (!genNodes.empty()) {
        std::vector<std::pair<uint32_t, uint32_t>> ev;   // (position, slot | sign)
        for (uint32_t g = 0; g < genNodes.size(); ++g) {
            ev.push_back({genNodes[g], g});
            if (F.subEnd[genNodes[g]] < N) ev.push_back({F.subEnd[genNodes[g]], g | 0x80000000u});
        }
        std::sort(ev.begin(), ev.end());
        F.evSlot.resize(ev.size());
        F.evIdx.assign(N, 0);
        size_t e = 0;
        for (uint64_t w = 0; w < N; ++w) {
            while (e < ev.size() && ev[e].first <= w) { F.evSlot[e] = ev[e].second; ++e; }
            F.evIdx[w] = static_cast<uint32_t>(e);
        }
        for (; e < ev.size(); ++e) F.evSlot[e] = ev[e].second;
    }
    // root's deltas (the root is always local)
    for (uint64_t j = d.node_offsets[0]; j < d.node_offsets[1]; ++j) {
        F.rootId.push_back(idAll[j]);
        F.rootChild.push_back(static_cast<uint32_t>(static_cast<int32_t>(d.delta_child[j])));
    }

    // ---- K2 tiles: ancestor chains and carry slots ----
    const uint32_t nShardNodes = F.nodeEnd - F.nodeBegin;
    F.nK2Tiles = (nShardNodes + kTileNodesK2H - 1) / kTileNodesK2H;
    F.chainOff.assign(F.nK2Tiles + 1, 0);
    F.carrySlot.assign(N, NONE);
    for (uint32_t t = 0; t < F.nK2Tiles; ++t) {
        const uint32_t a0 = F.nodeBegin + t * kTileNodesK2H;
        const uint32_t a1 = std::min(a0 + kTileNodesK2H, F.nodeEnd);
        std::vector<uint32_t> ch;
        for (uint32_t a = F.parent[a0]; a != NONE; a = F.parent[a]) ch.push_back(a);
        std::reverse(ch.begin(), ch.end());
        F.chainNodes.insert(F.chainNodes.end(), ch.begin(), ch.end());
        F.chainOff[t + 1] = static_cast<uint32_t>(F.chainNodes.size());
        for (uint32_t w = a0; w < a1; ++w) {
            const uint32_t p = F.parent[w];
            if (p != NONE && p < a0) F.carrySlot[w] = F.depth[p];
        }
    }

    // ---- selection: shard nodes in global BFS order ----
    F.bfsNodes.resize(nShardNodes); F.bfsRanks.resize(nShardNodes);
    {
        std::vector<uint32_t> idx(nShardNodes);
        std::iota(idx.begin(), idx.end(), F.nodeBegin);
        std::sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return F.bfsRank[a] < F.bfsRank[b]; });
        for (uint32_t i = 0; i < nShardNodes; ++i) { F.bfsNodes[i] = idx[i]; F.bfsRanks[i] = F.bfsRank[idx[i]]; }
    }

    // ---- canonical hashes of the four homopolymer k-mers (placement.cpp:41-76) ----
    for (int b = 0; b < 4; ++b) {
        const uint64_t bv = codeHash(static_cast<unsigned>(b)), cv = codeHash(static_cast<unsigned>(3 - b));
        uint64_t f = 0, r = 0;
        for (int i = 0; i < d.seed.k; ++i) { f ^= rol64(bv, static_cast<unsigned>(d.seed.k - i - 1)); r ^= rol64(cv, static_cast<unsigned>(d.seed.k - i - 1)); }
        F.homo[b] = f < r ? f : r;
    }
}

}  // namespace pm
