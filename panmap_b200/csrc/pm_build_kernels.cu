// pm_build_kernels.cu -- device side of the index builder (pm_build.cpp).
//
//   genome_materialize   one block per node: the aligned root template (every block's gap slots and main positions in coordinate order)
//                        + the point edits of the nodes on the root -> node path, applied level by level, + the block presence / strand
//                        state of that path -> the node's ungapped genome (absent blocks and '-' dropped, inverted blocks reversed and
//                        complemented): what the reference rebuilds from scratch for one node at a time (panmap_utils.cpp:7-180)
//   (seeding: the read path's kernels, pm_kernels.cu launchSeedListsEnd)
//   seeds_sort           one block per node: its seed list sorted in shared memory (bitonic), written to the compact arena
//   node_diff            persistent blocks over the nodes: sorted child list against the parent's sorted list -> the node's
//                        (hash, parentCount, childCount) deltas in ascending hash order -- two ordered partial lists (hashes whose count
//                        changed seen from the child, hashes only the parent has) merged by rank
// Nothing here is on the placement path; the kernels are sized for correctness and streams of a few GB, not tuned to a roofline.
#include "pm_internal.h"

namespace pm {

__device__ __forceinline__ char complementBase(char c) {
    switch (c) {
        case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
        case 'R': return 'Y'; case 'Y': return 'R'; case 'K': return 'M'; case 'M': return 'K';
        case 'B': return 'V'; case 'V': return 'B'; case 'D': return 'H'; case 'H': return 'D';
        default: return c;
    }
}

// block-wide exclusive scan of one unsigned per thread (blockDim.x <= 1024); returns the exclusive prefix, *total = the sum
__device__ __forceinline__ unsigned blockExclusiveScan(unsigned v, unsigned* sWarp /* [33] */, unsigned* total) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (unsigned)d) incl += o; }
    __syncthreads();
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const unsigned nW = (blockDim.x + 31) >> 5;
        unsigned w = lane < nW ? sWarp[lane] : 0u, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= (unsigned)d) wi += o; }
        if (lane < nW) sWarp[lane] = wi - w;
        if (lane == 31) sWarp[32] = wi;
    }
    __syncthreads();
    *total = sWarp[32];
    return sWarp[warp] + incl - v;
}

__global__ void __launch_bounds__(256) genome_materialize(BuildTreeView T, u32 nodeBegin, u32 nNodes, u32* __restrict__ pathScratch, unsigned char* __restrict__ blkScratch,
                                                          char* __restrict__ aligned, char* __restrict__ genomes, u64 pitch, u64* __restrict__ endOff) {
    __shared__ unsigned sWarp[33];
    __shared__ unsigned sDepth;
    const u32 local = blockIdx.x;
    if (local >= nNodes) return;
    const u32 v = nodeBegin + local;
    u32* path = pathScratch + (size_t)local * T.maxDepth;
    unsigned char* bst = blkScratch + (size_t)local * T.nBlocks;   // bit 0 present, bit 1 forward
    char* al = aligned + (size_t)local * T.nSlots;
    // template copy (all threads) while thread 0 lists the path and replays its block mutations (panmap_utils.cpp:93-111)
    for (u32 i = threadIdx.x; i < T.nSlots; i += blockDim.x) al[i] = T.tmpl[i];
    for (u32 b = threadIdx.x; b < T.nBlocks; b += blockDim.x) bst[b] = 2u;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned d = 0;
        for (u32 u = v;; u = T.parent[u]) { path[d++] = u; if (T.parent[u] == kBuildNoNode || d >= T.maxDepth) break; }
        sDepth = d;
        for (int lv = (int)d - 1; lv >= 0; --lv) {
            const u32 u = path[lv];
            for (u32 i = T.blockMutBegin[u]; i < T.blockMutBegin[u + 1]; ++i) {
                const u32 m = T.blockMut[i], b = m >> 2;
                const bool insertion = m & 1u, inversion = m & 2u;
                if (insertion) bst[b] = (unsigned char)(1u | (inversion ? 0u : 2u));
                else if (inversion) bst[b] ^= 2u;
                else bst[b] = 2u;
            }
        }
    }
    __syncthreads();
    const unsigned depth = sDepth;
    for (int lv = (int)depth - 1; lv >= 0; --lv) {   // root first: a later node overrides an earlier one
        const u32 u = path[lv];
        const u32 e0 = T.editBegin[u], e1 = T.editBegin[u + 1];
        if (T.editSerial[u]) { if (threadIdx.x == 0) for (u32 i = e0; i < e1; ++i) al[T.editSlot[i]] = T.editChar[i]; }   // a slot written twice: in order
        else for (u32 i = e0 + threadIdx.x; i < e1; i += blockDim.x) al[T.editSlot[i]] = T.editChar[i];
        __syncthreads();
    }
    // compaction in output order: position q of the aligned order reads slot q of a forward block, the mirrored slot of an inverted one
    const u32 per = (T.nSlots + blockDim.x - 1) / blockDim.x;
    const u32 q0 = min(T.nSlots, threadIdx.x * per), q1 = min(T.nSlots, q0 + per);
    auto charAt = [&](u32 q) -> char {
        const u32 b = T.slotBlock[q];
        const unsigned char st = bst[b];
        if (!(st & 1u)) return '-';
        if (st & 2u) return al[q];
        const u32 bs = T.blockStart[b], be = T.blockStart[b + 1];
        return complementBase(al[bs + (be - 1u - q)]);
    };
    unsigned cnt = 0;
    for (u32 q = q0; q < q1; ++q) cnt += charAt(q) != '-';
    unsigned total;
    unsigned o = blockExclusiveScan(cnt, sWarp, &total);
    char* g = genomes + (size_t)local * pitch;
    for (u32 q = q0; q < q1; ++q) { const char c = charAt(q); if (c != '-') g[o++] = c; }
    if (threadIdx.x == 0) endOff[local] = (u64)local * pitch + total;
}
void launchGenomeMaterialize(const BuildTreeView& T, u32 nodeBegin, u32 nNodes, u32* pathScratch, unsigned char* blkScratch, char* aligned, char* genomes,
                             u64 pitch, u64* endOff, cudaStream_t st) {
    if (!nNodes) return;
    noteLaunch(), genome_materialize<<<nNodes, 256, 0, st>>>(T, nodeBegin, nNodes, pathScratch, blkScratch, aligned, genomes, pitch, endOff);
}

// ---- seeds_sort: list r (count[r] hashes at in[winOff[r] ...]) sorted into arena[arenaOff[r] ...]; count[r] <= kSortCap ----
__global__ void __launch_bounds__(1024) seeds_sort(const u64* __restrict__ in, const u64* __restrict__ winOff, const u64* __restrict__ count,
                                                   const u64* __restrict__ arenaOff, u64* __restrict__ arena, u32 nLists) {
    extern __shared__ u64 sKeys[];
    const u32 r = blockIdx.x;
    if (r >= nLists) return;
    const unsigned n = (unsigned)count[r];
    unsigned m = 1;
    while (m < n) m <<= 1;
    const u64* src = in + winOff[r];
    for (unsigned i = threadIdx.x; i < m; i += blockDim.x) sKeys[i] = i < n ? src[i] : ~0ull;
    __syncthreads();
    for (unsigned k = 2; k <= m; k <<= 1)
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned i = threadIdx.x; i < m; i += blockDim.x) {
                const unsigned x = i ^ j;
                if (x > i) {
                    const u64 a = sKeys[i], b = sKeys[x];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { sKeys[i] = b; sKeys[x] = a; }
                }
            }
            __syncthreads();
        }
    u64* dst = arena + arenaOff[r];
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) dst[i] = sKeys[i];   // a real hash equal to the padding value sorts among the padding: still the n smallest
}
void launchSeedsSort(const u64* in, const u64* winOff, const u64* count, const u64* arenaOff, u64* arena, u32 nLists, cudaStream_t st) {
    if (!nLists) return;
    const size_t sm = (size_t)kBuildSortCap * sizeof(u64);
    cudaFuncSetAttribute(seeds_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    noteLaunch(), seeds_sort<<<nLists, 1024, sm, st>>>(in, winOff, count, arenaOff, arena, nLists);
}

// ---- node_diff ----
__device__ __forceinline__ unsigned lowerBound(const u64* __restrict__ a, unsigned n, u64 x) {
    unsigned lo = 0, hi = n;
    while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}
struct DiffEntry { u64 h; u32 pc, cc; };
__global__ void __launch_bounds__(1024) node_diff(BuildDiffArgs A) {
    __shared__ unsigned sWarp[33];
    __shared__ unsigned long long sBase;
    DiffEntry* listA = reinterpret_cast<DiffEntry*>(A.scratch) + (size_t)blockIdx.x * 2 * kBuildSortCap;
    DiffEntry* listB = listA + kBuildSortCap;
    for (u32 local = blockIdx.x; local < A.nNodes; local += gridDim.x) {
        const u32 v = A.nodeBegin + local;
        const u64* c = A.listPtr[v]; const unsigned nc = (unsigned)A.listCount[v];
        const u32 par = A.parent[v];
        const u64* p = par == kBuildNoNode ? nullptr : A.listPtr[par];
        const unsigned np = par == kBuildNoNode ? 0u : (unsigned)A.listCount[par];
        // pass A: hashes of the child (first of each run) whose multiplicity differs from the parent's
        unsigned perA = (nc + blockDim.x - 1) / blockDim.x, i0 = min(nc, threadIdx.x * perA), i1 = min(nc, i0 + perA), cntA = 0;
        for (unsigned i = i0; i < i1; ++i) {
            if (i && c[i] == c[i - 1]) continue;
            unsigned cc = 1; while (i + cc < nc && c[i + cc] == c[i]) ++cc;
            const unsigned lb = lowerBound(p, np, c[i]); unsigned pc = 0; while (lb + pc < np && p[lb + pc] == c[i]) ++pc;
            cntA += pc != cc;
        }
        unsigned totA; unsigned oA = blockExclusiveScan(cntA, sWarp, &totA);
        for (unsigned i = i0; i < i1; ++i) {
            if (i && c[i] == c[i - 1]) continue;
            unsigned cc = 1; while (i + cc < nc && c[i + cc] == c[i]) ++cc;
            const unsigned lb = lowerBound(p, np, c[i]); unsigned pc = 0; while (lb + pc < np && p[lb + pc] == c[i]) ++pc;
            if (pc != cc) listA[oA++] = DiffEntry{c[i], pc, cc};
        }
        // pass B: hashes only the parent has
        unsigned perB = (np + blockDim.x - 1) / blockDim.x, j0 = min(np, threadIdx.x * perB), j1 = min(np, j0 + perB), cntB = 0;
        for (unsigned j = j0; j < j1; ++j) {
            if (j && p[j] == p[j - 1]) continue;
            const unsigned lb = lowerBound(c, nc, p[j]);
            cntB += !(lb < nc && c[lb] == p[j]);
        }
        unsigned totB; unsigned oB = blockExclusiveScan(cntB, sWarp, &totB);
        for (unsigned j = j0; j < j1; ++j) {
            if (j && p[j] == p[j - 1]) continue;
            const unsigned lb = lowerBound(c, nc, p[j]);
            if (!(lb < nc && c[lb] == p[j])) { unsigned pc = 1; while (j + pc < np && p[j + pc] == p[j]) ++pc; listB[oB++] = DiffEntry{p[j], pc, 0u}; }
        }
        __syncthreads();
        const unsigned tot = totA + totB;
        if (threadIdx.x == 0) {
            sBase = atomicAdd(A.cursor, (unsigned long long)tot);
            A.nodeOff[local] = sBase; A.nodeCnt[local] = tot;
        }
        __syncthreads();
        const unsigned long long base = sBase;
        if (base + tot <= A.outCap) {   // merge by rank: the two lists are sorted and share no hash
            for (unsigned i = threadIdx.x; i < totA; i += blockDim.x) {
                const DiffEntry e = listA[i];
                unsigned lo = 0, hi = totB; while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (listB[mid].h < e.h) lo = mid + 1; else hi = mid; }
                const unsigned long long o = base + i + lo;
                A.outHash[o] = e.h; A.outPc[o] = (short)min(e.pc, 32767u); A.outCc[o] = (short)min(e.cc, 32767u);
            }
            for (unsigned j = threadIdx.x; j < totB; j += blockDim.x) {
                const DiffEntry e = listB[j];
                unsigned lo = 0, hi = totA; while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (listA[mid].h < e.h) lo = mid + 1; else hi = mid; }
                const unsigned long long o = base + j + lo;
                A.outHash[o] = e.h; A.outPc[o] = (short)min(e.pc, 32767u); A.outCc[o] = 0;
            }
        }
        __syncthreads();
    }
}
void launchNodeDiff(const BuildDiffArgs& A, unsigned grid, cudaStream_t st) {
    if (!A.nNodes) return;
    noteLaunch(), node_diff<<<grid, 1024, 0, st>>>(A);
}

}  // namespace pm
