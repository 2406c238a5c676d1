// pm_device.cuh -- device-side helpers shared by the kernel translation units (pm_kernels.cu, pm_shard_kernels.cu).
#pragma once
#include "pm_kernels.cuh"

namespace pm {

// ------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 shflU64(u64 v, int srcLane) {
    return __shfl_sync(0xffffffffu, v, srcLane);
}
__device__ __forceinline__ u64 shflUpU64(u64 v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ u64 shflXorU64(u64 v, int d) { return __shfl_xor_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shflXorF64(double v, int d) { return __shfl_xor_sync(0xffffffffu, v, d); }

__device__ __forceinline__ void fxAtomicAdd(u64* acc /* lo, hi */, fx128 v) {
    // exact 128-bit accumulation with two 64-bit atomics: the number of carries out of the low word does not
    // depend on the order of the additions, so the result is deterministic.
    const u64 old = atomicAdd(reinterpret_cast<unsigned long long*>(&acc[0]), v.lo);
    const u64 carry = (old + v.lo < old) ? 1ULL : 0ULL;
    const u64 hiAdd = (u64)v.hi + carry;
    if (hiAdd) atomicAdd(reinterpret_cast<unsigned long long*>(&acc[1]), hiAdd);
}
__device__ __forceinline__ fx128 fxWarpSum(fx128 v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        fx128 o; o.lo = shflXorU64(v.lo, d); o.hi = (i64)shflXorU64((u64)v.hi, d);
        v = fxAdd(v, o);
    }
    return v;
}
__device__ __forceinline__ long long warpSumLL(long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += (long long)shflXorU64((u64)v, d);
    return v;
}

// ------------------------------------------------------------------------------------------------------
// count table insert (open addressing, linear probing; keys are 64-bit seed hashes)
// ------------------------------------------------------------------------------------------------------
// SampleAcc::overflow is a set of kOvf* bits raised from many threads
__device__ __forceinline__ void raiseFlag(SampleAcc* acc, long long bits) { atomicOr(reinterpret_cast<unsigned long long*>(&acc->overflow), (unsigned long long)bits); }
__device__ __forceinline__ void tableInsert(TableSlot* __restrict__ table, u64 mask, u64 h, u32 add, SampleAcc* acc) {
    if (h == kEmptyKey) { atomicAdd((unsigned long long*)&acc->emptyKeyCount, (unsigned long long)add); return; }
    u64 slot = mixKey(h) & mask;
    for (int probe = 0; probe < 8192; ++probe) {
        // keys are write-once (EMPTY -> key), so a possibly stale L1 copy is safe: a stale EMPTY is resolved by the CAS, a cached
        // non-EMPTY key is final.  Hot seeds therefore hit L1 and only the count update travels to L2.
        u64 cur = __ldca(&table[slot].key);
        if (cur == kEmptyKey) {
            cur = atomicCAS((unsigned long long*)&table[slot].key, (unsigned long long)kEmptyKey, (unsigned long long)h);
            if (cur == kEmptyKey) cur = h;
        }
        if (cur == h) { atomicAdd(&table[slot].count, add); return; }
        slot = (slot + 1) & mask;
    }
    raiseFlag(acc, kOvfTable);
}

__device__ __forceinline__ uint4 ldSlot(const TableSlot* t, u64 i) { return __ldcs(reinterpret_cast<const uint4*>(t) + i); }
__device__ __forceinline__ u64 slotKey(const uint4& v) { return (u64)v.x | ((u64)v.y << 32); }

__device__ __forceinline__ long long resolveMinSupport(long long multiSum, long long multiCount, int configured) {
    if (configured >= 0) return configured;
    const double est = multiCount > 0 ? (double)(u64)multiSum / (double)(u64)multiCount : 0.0;
    return est > 3.0 ? 2 : 1;
}

static inline unsigned streamGrid(u64 n, unsigned perThread) {
    u64 g = (n + 256ull * perThread - 1) / (256ull * perThread);
    if (g > 148 * 16) g = 148 * 16;
    return (unsigned)(g ? g : 1);
}

// true in exactly one block of the grid: the one whose threads get here last (everything the other blocks wrote before is visible to it).
// *ctr must be zero before the launch and is zero again afterwards; sFlag: one word of shared memory.
__device__ __forceinline__ bool lastBlockDone(unsigned* ctr, unsigned* sFlag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const bool last = atomicAdd(ctr, 1u) == gridDim.x * gridDim.y - 1u;
        if (last) *ctr = 0u;
        *sFlag = last ? 1u : 0u;
    }
    __syncthreads();
    const bool r = *sFlag != 0u;
    if (r) __threadfence();
    return r;
}

// ---- tolerance chain over a set of prefix-maximum records (placement.cpp:355-371), one 256-thread block per metric ----------------
// Records are unordered: every step picks the lowest-rank record after the last event that beats best + tol.  The valid records are
// first staged in shared memory (a handful per metric in practice), so a step costs shared-memory latency instead of an L2 round trip
// per record; more than kChainCap records are walked where they lie.
constexpr int kChainCap = 1536;
struct ChainShared {
    unsigned long long sMin[8]; unsigned long long sPick; unsigned sN;
    u32 rank[kChainCap]; u32 node[kChainCap]; double score[kChainCap];
};
template <class RecAt>   // recAt(idx, rank, score, node) -> record idx exists
__device__ __forceinline__ Selection chainReplay(ChainShared& S, unsigned total, RecAt recAt) {
    const unsigned tid = threadIdx.x;
    if (tid == 0) S.sN = 0;
    __syncthreads();
    for (unsigned idx = tid; idx < total; idx += 256) {
        u32 rk, nd; double sc;
        if (!recAt(idx, rk, sc, nd)) continue;
        const unsigned o = atomicAdd(&S.sN, 1u);
        if (o < (unsigned)kChainCap) { S.rank[o] = rk; S.node[o] = nd; S.score[o] = sc; }
    }
    __syncthreads();
    const unsigned n = S.sN;
    const bool staged = n <= (unsigned)kChainCap;
    const unsigned span = staged ? n : total;
    double best = 0.0; u32 bestNode = kNone; long long lastRank = -1;
    while (true) {
        const double thr = best + fmax(best * 0.0001, 1e-9);
        unsigned long long pick = ~0ULL;
        for (unsigned i = tid; i < span; i += 256) {
            u32 rk, nd; double sc;
            if (staged) { rk = S.rank[i]; sc = S.score[i]; }
            else if (!recAt(i, rk, sc, nd)) continue;
            if ((long long)rk > lastRank && sc > thr) {
                const unsigned long long key = ((unsigned long long)rk << 32) | i;
                pick = key < pick ? key : pick;
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { const unsigned long long o = shflXorU64(pick, d); pick = o < pick ? o : pick; }
        if ((tid & 31) == 0) S.sMin[tid >> 5] = pick;
        __syncthreads();
        if (tid == 0) { unsigned long long v = S.sMin[0]; for (int w = 1; w < 8; ++w) v = S.sMin[w] < v ? S.sMin[w] : v; S.sPick = v; }
        __syncthreads();
        const unsigned long long p = S.sPick;
        __syncthreads();
        if (p == ~0ULL) break;
        const unsigned i = (unsigned)(p & 0xFFFFFFFFu);
        if (staged) { best = S.score[i]; bestNode = S.node[i]; }
        else { u32 rk; recAt(i, rk, best, bestNode); }
        lastRank = (long long)(p >> 32);
    }
    Selection s; s.best = best; s.bestNode = bestNode; s.lastRank = lastRank < 0 ? kNone : (u32)lastRank;
    return s;
}

// ---- computeReadSeedMagnitudes pieces shared by entries_finalize (one GPU) and gathered_finalize (one sample over several GPUs) ----
constexpr int kHistSmem = 2048;
struct FinalizeAcc { fx128 mag, lsum; long long kept; u32 maxc; };
struct FinalizeShared { u64 red[8][4]; long long kept[8]; unsigned mx[8]; };
// a kept seed with read count c: log1p from the host-computed table, exact sums, count histogram; returns log1p(c)
__device__ __forceinline__ double finalizeSums(const DevIndexView& I, const WorkspaceView& W, u32 c, FinalizeAcc& A, unsigned* sHist) {
    const double l = c < (u32)kLog1pLut ? __ldg(&I.log1pLut[c]) : log1p((double)c);
    ++A.kept; A.maxc = max(A.maxc, c);
    if (c < (u32)kLog1pLut) { if (c < (u32)kHistSmem) atomicAdd(&sHist[c], 1u); else atomicAdd(&W.countHist[c], 1u); }
    A.mag = fxAdd(A.mag, fxFromDouble(l * l));
    A.lsum = fxAdd(A.lsum, fxFromDouble(l));
    return l;
}
// seed hash -> dense seed id of the index, kNone when the index does not hold it
__device__ __forceinline__ u32 dictLookup(const DevIndexView& I, u64 k) {
    if (k == kEmptyKey) return kNone;
    u64 s = mixKey(k) & I.dictMask;
    while (true) {
        const uint4 d = __ldg(reinterpret_cast<const uint4*>(I.dict) + s);
        const u64 dk = (u64)d.x | ((u64)d.y << 32);
        if (dk == k) return d.z;
        if (dk == kEmptyKey) return kNone;
        s = (s + 1) & I.dictMask;
    }
}
// 256-thread blocks: histogram flush + this block's FinPartial
__device__ __forceinline__ void finalizeBlockEpilogue(const WorkspaceView& W, const FinalizeAcc& A, unsigned* sHist, FinalizeShared* sh) {
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    __syncthreads();
    for (int i = tid; i < kHistSmem; i += blockDim.x) if (sHist[i]) atomicAdd(&W.countHist[i], sHist[i]);
    const fx128 mag = fxWarpSum(A.mag), lsum = fxWarpSum(A.lsum);
    const long long kept = warpSumLL(A.kept);
    unsigned mx = A.maxc;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if (lane == 0) { sh->red[warp][0] = mag.lo; sh->red[warp][1] = (u64)mag.hi; sh->red[warp][2] = lsum.lo; sh->red[warp][3] = (u64)lsum.hi; sh->kept[warp] = kept; sh->mx[warp] = mx; }
    __syncthreads();
    if (tid == 0) {
        fx128 m = fxZero(), l = fxZero(); long long kp = 0; unsigned mm = 0;
        for (int q = 0; q < 8; ++q) {
            fx128 t; t.lo = sh->red[q][0]; t.hi = (i64)sh->red[q][1]; m = fxAdd(m, t);
            t.lo = sh->red[q][2]; t.hi = (i64)sh->red[q][3]; l = fxAdd(l, t);
            kp += sh->kept[q]; mm = max(mm, sh->mx[q]);
        }
        FinPartial P; P.mag[0] = m.lo; P.mag[1] = (u64)m.hi; P.lsum[0] = l.lo; P.lsum[1] = (u64)l.hi; P.kept = kp; P.maxc = mm;
        W.finPart[blockIdx.x] = P;
    }
}

}  // namespace pm
