// pm_kernels.cu -- hand-written sm_100a kernels of the placement hot path.
//
//   pack_reads      ASCII reads -> 4-bit base codes, 32 per 16-byte chunk, every read 16-byte aligned
//   seed_reads      lane-per-read rolling hashes (k-mer + s-mer, both strands), sliding s-mer minimum,
//                   closed/open syncmer test, k-min-mers, insertion into the open-addressing count table
//   table_*         homopolymer removal, auto min-support statistics, log1p + exact magnitude sums,
//                   scatter of log1p(count) into the dense per-seed-id array the delta kernel gathers from
//   node_deltas     K1: streams the (seedId, parent|child) delta arrays once, per-node parent-relative sums
//   prefix_scores   K2: exact (128-bit fixed point) tree prefix over the DFS order, the five scores
//   bfs_* / chain / ties   selection with the reference's order-dependent tolerance chain (placement.cpp:355-401)
//
// This is integer/byte streaming work bounded by HBM and issue rate; no tensor cores are involved.
#include "pm_device.cuh"
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <vector>

namespace pm {

// ------------------------------------------------------------------------------------------------------
// pack_reads
// ------------------------------------------------------------------------------------------------------
// One thread per 32-base chunk.  The block gets the read of its first chunk from a small host-built table, stages the
// chunk/byte offsets of the <= 257 reads it can touch in shared memory, and every thread then locates its read there.
// Source bytes are fetched as nine aligned 32-bit words per thread (neighbouring threads read neighbouring 32-byte
// segments, so the lines are shared through L1) and realigned with funnel shifts; base -> code via a shared 256-byte table.
template <bool HAS_END>   // HAS_END: reads were homopolymer-compressed in place, endOff[r] is where read r now ends
__global__ void __launch_bounds__(256) pack_reads(const char* __restrict__ reads, const u64* __restrict__ off,
                                                  const u64* __restrict__ packedOff, const u32* __restrict__ blockFirst, u64 nReads,
                                                  u64 gBase, u64 nChunks, uint4* __restrict__ packed, const u64* __restrict__ endOff) {
    __shared__ u64 sPO[258];
    __shared__ u64 sOff[258];
    __shared__ u64 sEnd[HAS_END ? 258 : 1];
    __shared__ unsigned char sLut[256];
    __shared__ u64 sFirst;
    const u64 g0 = gBase + (u64)blockIdx.x * 256;   // gBase: first chunk of the read slice this launch covers
    sLut[threadIdx.x] = (unsigned char)baseCode((unsigned char)threadIdx.x);
    if (threadIdx.x == 0) sFirst = blockFirst[blockIdx.x];   // read owning this block's first chunk (computed with the chunk offsets on the host)
    __syncthreads();
    const u64 rFirst = sFirst;
    for (int i = threadIdx.x; i < 258; i += 256) {
        const u64 r = rFirst + i;
        sPO[i] = r <= nReads ? __ldg(&packedOff[r]) : ~0ULL;
        sOff[i] = r <= nReads ? __ldg(&off[r]) : 0;
        if (HAS_END) sEnd[i] = r < nReads ? __ldg(&endOff[r]) : 0;   // homopolymer-compressed reads end before the next one begins
    }
    __syncthreads();
    const u64 g = g0 + threadIdx.x;
    if (g >= gBase + nChunks) return;
    int lo = 0, hi = 257;   // largest i with sPO[i] <= g
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (sPO[mid] <= g) lo = mid; else hi = mid;
    }
    u64 c, src, e;
    if (sPO[hi] <= g) {
        // more than 256 reads behind this block's first chunk (runs of empty reads): locate the read in global memory
        u64 glo = rFirst + 257, ghi = nReads;
        while (ghi - glo > 1) {
            const u64 mid = (glo + ghi) >> 1;
            if (__ldg(&packedOff[mid]) <= g) glo = mid; else ghi = mid;
        }
        c = g - __ldg(&packedOff[glo]); src = __ldg(&off[glo]) + 32 * c; e = HAS_END ? __ldg(&endOff[glo]) : __ldg(&off[glo + 1]);
    } else {
        c = g - sPO[lo]; src = sOff[lo] + 32 * c; e = HAS_END ? sEnd[lo] : sOff[lo + 1];
    }
    const int n = (HAS_END && e <= src) ? 0 : (int)((e - src) < 32 ? (e - src) : 32);
    const unsigned* __restrict__ wsrc = reinterpret_cast<const unsigned*>(reads + (src & ~3ULL));
    const unsigned sh = (unsigned)(src & 3ULL) * 8u;
    unsigned x[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) x[q] = __ldg(wsrc + q);
    unsigned w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        unsigned out = 0;
#pragma unroll
        for (int hw = 0; hw < 2; ++hw) {
            const unsigned bytes = __funnelshift_r(x[2 * q + hw], x[2 * q + hw + 1], sh);
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
                const int j = 8 * q + 4 * hw + bb;
                const unsigned code = j < n ? sLut[(bytes >> (8 * bb)) & 0xFFu] : 4u;
                out |= code << (4 * (4 * hw + bb));
            }
        }
        w[q] = out;
    }
    packed[g] = make_uint4(w[0], w[1], w[2], w[3]);
}

void launchPackReads(const char* reads, const u64* off, const u64* packedOff, const u32* blockFirst, u64 nReads, u64 gBase, u64 nChunks,
                     uint4* packed, cudaStream_t st, const u64* endOff) {
    if (nChunks == 0) return;
    const unsigned grid = (unsigned)((nChunks + 255) / 256);
    if (endOff) noteLaunch(), pack_reads<true><<<grid, 256, 0, st>>>(reads, off, packedOff, blockFirst, nReads, gBase, nChunks, packed, endOff);
    else noteLaunch(), pack_reads<false><<<grid, 256, 0, st>>>(reads, off, packedOff, blockFirst, nReads, gBase, nChunks, packed, nullptr);
}

// Chunk offsets of a slice of reads on the device: packedOff[i] = gBase + sum_{j<i} ceil(len_j / 32), i = 0..n (n+1 entries), so
// that they need not cross PCIe (8 bytes per read).  Two small launches: per-tile sums (4096 reads per block), then every block
// adds the sums of the tiles before it to an in-tile exclusive scan.
constexpr int kOffTile = 4096;
__device__ __forceinline__ u64 blockSumU64(u64 v, u64* sRed) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += shflXorU64(v, d);
    if ((threadIdx.x & 31) == 0) sRed[threadIdx.x >> 5] = v;
    __syncthreads();
    u64 t = 0;
    for (int w = 0; w < 32; ++w) t += sRed[w];
    __syncthreads();
    return t;
}
__global__ void __launch_bounds__(1024) chunk_tile_sums(const u64* __restrict__ off, u64 n, u64* __restrict__ tileSum) {
    __shared__ u64 sRed[32];
    const u64 base = (u64)blockIdx.x * kOffTile + 4ull * threadIdx.x;
    u64 v = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) if (base + q < n) v += (off[base + q + 1] - off[base + q] + 31) >> 5;
    v = blockSumU64(v, sRed);
    if (threadIdx.x == 0) tileSum[blockIdx.x] = v;
}
__global__ void __launch_bounds__(1024) chunk_offsets_tiled(const u64* __restrict__ off, u64 n, u64 gBase, const u64* __restrict__ tileSum,
                                                            u64* __restrict__ packedOff) {
    __shared__ u64 sRed[32];
    __shared__ u64 sWarp[32];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    u64 pre = 0;
    for (unsigned b = tid; b < blockIdx.x; b += 1024) pre += tileSum[b];
    pre = blockSumU64(pre, sRed) + gBase;
    const u64 base = (u64)blockIdx.x * kOffTile + 4ull * tid;
    u64 c[4], tot = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { c[q] = base + q < n ? (off[base + q + 1] - off[base + q] + 31) >> 5 : 0; tot += c[q]; }
    u64 incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u64 o = shflUpU64(incl, d); if (lane >= (unsigned)d) incl += o; }
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    for (unsigned w = 0; w < warp; ++w) pre += sWarp[w];
    u64 run = pre + incl - tot;
#pragma unroll
    for (int q = 0; q < 4; ++q) { if (base + q <= n) packedOff[base + q] = run; run += c[q]; }   // entry n = the slice total
}
void launchChunkOffsets(const u64* off, u64 n, u64 gBase, u64* tileSum, u64* packedOff, cudaStream_t st) {
    const unsigned tiles = (unsigned)(n / kOffTile + 1);   // +1: the tile that holds entry n
    noteLaunch(), chunk_tile_sums<<<tiles, 1024, 0, st>>>(off, n, tileSum);
    noteLaunch(), chunk_offsets_tiled<<<tiles, 1024, 0, st>>>(off, n, gBase, tileSum, packedOff);
}

// ------------------------------------------------------------------------------------------------------
// Homopolymer compression of the reads when the index was built with --hpc (placement.cpp:1145-1165, seeding::hpcCompress,
// seeding.cpp:286-306): a base is dropped when it equals its predecessor ignoring case.  One warp per read, in place inside the
// read's own byte range (the write position never overtakes the read position); endOff[r] = one past the last kept byte.
// ------------------------------------------------------------------------------------------------------
// quals (optional, --min-seed-quality): the quality of the first base of every run moves along with it (placement.cpp:1147-1159)
__global__ void __launch_bounds__(256) hpc_compress(char* __restrict__ reads, const u64* __restrict__ off, u64 nReads, u64* __restrict__ endOff,
                                                    char* __restrict__ quals) {
    const unsigned lane = threadIdx.x & 31u;
    const u64 warpsTotal = (u64)gridDim.x * (blockDim.x >> 5);
    for (u64 r = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nReads; r += warpsTotal) {
        const u64 b = off[r], L = off[r + 1] - b;
        u64 written = 0;
        int prevUp = -1;   // upper-cased last byte of the previous window
        for (u64 i0 = 0; i0 < L; i0 += 32) {
            const u64 i = i0 + lane;
            const int c = i < L ? (int)(unsigned char)reads[b + i] : -2;
            const char qc = (quals && i < L) ? quals[b + i] : 0;
            const int up = (c >= 'a' && c <= 'z') ? c - 32 : c;
            int before = __shfl_up_sync(0xffffffffu, up, 1);
            if (lane == 0) before = prevUp;
            const bool keep = i < L && up != before;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            __syncwarp();
            if (keep) { const u64 o = b + written + __popc(m & ((1u << lane) - 1u)); reads[o] = (char)c; if (quals) quals[o] = qc; }
            __syncwarp();
            written += __popc(m);
            prevUp = __shfl_sync(0xffffffffu, up, 31);
        }
        if (lane == 0) endOff[r] = b + written;
    }
}
void launchHpcCompress(char* reads, const u64* off, u64 nReads, u64* endOff, cudaStream_t st, char* quals) {
    if (nReads == 0) return;
    u64 g = (nReads + 7) / 8; if (g > 148ull * 8) g = 148ull * 8;
    noteLaunch(), hpc_compress<<<(unsigned)g, 256, 0, st>>>(reads, off, nReads, endOff, quals);
}

// ------------------------------------------------------------------------------------------------------
// --dedup (placement.cpp:1550-1620 with dedupReads): every distinct read STRING counts once.  One warp per read: position-tagged
// byte hash, open-addressing set of (31-bit tag, read index) claimed with one 64-bit CAS, and a byte-for-byte comparison against
// the read that holds a slot with the same tag -- the reference compares the raw strings, so case and the exact ambiguity
// letters matter and the packed codes cannot be used.  Which copy of a duplicated read survives is irrelevant (same seeds).
// Slices of a sample insert into the same set one after the other, so the decision is global.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dedup_mark(const char* __restrict__ reads, const u64* __restrict__ off, u64 rBegin, u64 rEnd,
                                                  unsigned long long* slots, u64 mask, unsigned char* __restrict__ dup,
                                                  const u64* __restrict__ endOff) {
    const unsigned lane = threadIdx.x & 31u;
    const u64 warpsTotal = (u64)gridDim.x * (blockDim.x >> 5);
    for (u64 r = rBegin + (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rEnd; r += warpsTotal) {
        const u64 b = off[r], L = (endOff ? endOff[r] : off[r + 1]) - b;
        u64 h = 0;
        for (u64 i = lane; i < L; i += 32) h += mixKey((u64)(unsigned char)reads[b + i] + (i << 8) + 0x51ED270B1ULL);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) h += shflXorU64(h, d);
        h = mixKey(h ^ (L * 0x9E3779B97F4A7C15ULL));
        const u64 tag = (h >> 33) & 0x7FFFFFFFULL;
        const unsigned long long mine = (tag << 32) | (r & 0xFFFFFFFFULL);
        u64 s = h & mask;
        bool isDup = false;
        for (u64 probe = 0; probe <= mask; ++probe) {
            unsigned long long cur = 0;
            if (lane == 0) cur = atomicCAS(&slots[s], ~0ULL, mine);
            cur = shflU64(cur, 0);
            if (cur == ~0ULL) break;   // claimed: first read with this string
            if ((cur >> 32) == tag) {
                const u64 o = cur & 0xFFFFFFFFULL;
                const u64 bo = off[o];
                bool eq = ((endOff ? endOff[o] : off[o + 1]) - bo) == L;
                if (eq) for (u64 i = lane; i < L; i += 32) eq = eq && reads[b + i] == reads[bo + i];
                if (__all_sync(0xffffffffu, eq)) { isDup = true; break; }
            }
            s = (s + 1) & mask;
        }
        if (lane == 0) dup[r] = isDup ? 1 : 0;
    }
}
void launchDedup(const char* reads, const u64* off, u64 rBegin, u64 rEnd, unsigned long long* slots, u64 mask, unsigned char* dup, cudaStream_t st,
                 const u64* endOff) {
    if (rEnd <= rBegin) return;
    u64 g = (rEnd - rBegin + 7) / 8; if (g > 148ull * 8) g = 148ull * 8;
    noteLaunch(), dedup_mark<<<(unsigned)g, 256, 0, st>>>(reads, off, rBegin, rEnd, slots, mask, dup, endOff);
}

// ------------------------------------------------------------------------------------------------------
// Seeding = two kernels, mirroring the reference's own split:
//   syncmers_*          seeding::rollingSyncmers (seeding.cpp:47-229): one lane per read, all lanes of a warp at the same
//                       read position (lock-step, so the block-end pass of the sliding minimum is convergent); each lane
//                       appends the canonical hashes of its read's syncmers (trim filter applied) to the read's own
//                       region of synBuf.  No table traffic, no warp collectives in the hot loop.
//   seeds_from_syncmers placement.cpp:1598-1686: one warp per read, lanes over consecutive syncmers; k-min-mers in closed
//                       form, then the open-addressing count table (latency-bound probes, hidden by full occupancy).
// A read's syncmer region starts at slot 32 * packedOff[r] (its packed chunks cover >= len slots).
// ------------------------------------------------------------------------------------------------------
constexpr int kSeedThreads = 128;

// generic (any k <= 32, s, t, open/closed).  MODE 0: hashes -> synBuf (+ count); MODE 1: (hash, isReverse, pos) lists.
// MODE 3 (--min-seed-quality, placement.cpp:1179-1240 / 1388-1533): EVERY syncmer goes to synBuf, and outRev[same index] says whether
// it "passes": start inside the trimmed range and average Phred quality over its k bases >= minQ.  (int)(signed char)q - 33 summed
// over k bases and compared with minQ * k is the reference's `double(sum) / k < minQ` exactly (|sum / k - minQ| >= 1/k when unequal).
template <int MODE>
__global__ void __launch_bounds__(kSeedThreads) syncmers_generic(const uint4* __restrict__ packed, const u64* __restrict__ off,
                                                                 const u64* __restrict__ packedOff, const u64* __restrict__ winOff,
                                                                 u64 nReads, SeederParams P, const SeedTables* __restrict__ gT,
                                                                 u64* __restrict__ synBuf, unsigned* __restrict__ synCount,
                                                                 u64* outHash, unsigned char* outRev, long long* outPos, u64* outCount,
                                                                 const unsigned char* __restrict__ dup, const u64* __restrict__ endOff,
                                                                 const char* __restrict__ quals, int minQualSum) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    SeedTables* sT = reinterpret_cast<SeedTables*>(smemRaw);
    u64* rings = reinterpret_cast<u64*>(smemRaw + sizeof(SeedTables));
    for (int i = threadIdx.x; i < (int)(sizeof(SeedTables) / 8); i += blockDim.x)
        reinterpret_cast<u64*>(sT)[i] = reinterpret_cast<const u64*>(gT)[i];
    __syncthreads();
    const SeedTables& T = *sT;
    const unsigned lane = threadIdx.x & 31u;
    const u64 warpsTotal = (u64)gridDim.x * (blockDim.x >> 5);
    const u64 warpId = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (u64 r0 = warpId * 32; r0 < nReads; r0 += warpsTotal * 32) {
        const u64 r = r0 + lane;
        const bool valid = r < nReads;
        const u64 b = valid ? off[r] : 0;
        int L = valid ? (int)((endOff ? endOff[r] : off[r + 1]) - b) : 0;
        if (L < P.k) L = 0;  // shorter than k: no windows (seeding.cpp:50)
        if (dup && valid && dup[r]) L = 0;   // --dedup: a byte-identical read was seen before (placement.cpp:1550-1620)
        const u64 pOff = valid ? packedOff[r] : 0;
        const int nCh = (L + 31) >> 5;
        int maxL = L;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) maxL = max(maxL, __shfl_xor_sync(0xffffffffu, maxL, d));
        const int nChMax = (maxL + 31) >> 5;
        const int validEnd = L - P.trimEnd - P.k;
        ReadSeederT<kSeedThreads> sd;
        sd.reset(rings + threadIdx.x, P.w);
        unsigned cnt = 0;
        const u64 obase = (MODE == 0 || MODE == 3) ? pOff * 32 : (valid ? winOff[r] : 0);
#pragma unroll 1
        for (int c = 0; c < nChMax; ++c) {
            uint4 v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
            if (c < nCh) v = packed[pOff + c];
#pragma unroll 1
            for (int wi = 0; wi < 4; ++wi) {
                const unsigned word = wi == 0 ? v.x : wi == 1 ? v.y : wi == 2 ? v.z : v.w;
                if (c * 32 + wi * 8 >= maxL) break;
#pragma unroll 2
                for (int bi = 0; bi < 8; ++bi) {
                    const int i = c * 32 + wi * 8 + bi;
                    if (i < L) {
                        const unsigned code = (word >> (4 * bi)) & 0xFu;
                        u64 h; bool rev;
                        if (sd.pushBase(i, code, T, P, h, rev)) {
                            const int pos = i - P.k + 1;
                            if (MODE == 1) { outHash[obase + cnt] = h; outRev[obase + cnt] = rev ? 1 : 0; outPos[obase + cnt] = pos; ++cnt; }
                            else if (MODE == 3) {
                                int qs = 0;
                                for (int q = 0; q < P.k; ++q) qs += (int)(signed char)quals[b + pos + q] - 33;
                                synBuf[obase + cnt] = h;
                                outRev[obase + cnt] = (pos >= P.trimStart && pos <= validEnd && qs >= minQualSum) ? 1 : 0;
                                ++cnt;
                            }
                            else if (pos >= P.trimStart && pos <= validEnd) { synBuf[obase + cnt] = h; ++cnt; }
                        }
                    }
                }
            }
        }
        if (valid) { if (MODE == 0 || MODE == 3) synCount[r] = cnt; else outCount[r] = cnt; }
    }
}

// Closed syncmers with t == 0 (panmap's default) and compile-time (k, s): the main loop is unrolled over one block of
// W = k-s+1 s-mers so that the ring slot is a compile-time constant -- ring addresses are immediates and the block-start /
// block-end cases and the window geometry are resolved statically.  With t == 0 the closed-syncmer test only asks whether
// the OLDEST or the NEWEST s-mer of the window attains the window minimum:
//   newest: fs == min(window)                    (fs is in a register)
//   oldest: F[p] == min(window)  <=>  F[p] is the suffix minimum of its block tail (one bit per slot, produced by the
//           block-end pass) and that suffix minimum is <= the running minimum of the current block
// so the ring keeps only the in-place suffix minima (2*W words per lane instead of 4*W) plus two W-bit masks.
// ASCII: the kernel reads the bases themselves (`reads`, one byte per base) and turns 8 of them at a time into the 4-bit codes
// through a 256-byte table in shared memory -- no pack_reads pass and no packed copy of the sample; !ASCII: 4-bit codes from `packed`.
// NOTRIM: trimEnd == 0 (the default), so "window inside the trimmed range" is implied by i < L and costs nothing
// (mask & bit) != 0 through an opaque AND: the front end otherwise rewrites it as (mask >> p) & 1 == 1, three ALU instructions
// where ptxas has one LOP3 with a predicate result
__device__ __forceinline__ bool maskBit(unsigned mask, unsigned bit) {
    unsigned t;
    asm("and.b32 %0, %1, %2;" : "=r"(t) : "r"(mask), "r"(bit));
    return t != 0u;
}
constexpr int kPairStride = 96;   // >= 12 * 7 + 8 entries, a multiple of 16 so that every table starts on the same bank
template <int K, int S, bool ASCII, bool NOTRIM>
__global__ void __launch_bounds__(kSeedThreads) syncmers_fast(const uint4* __restrict__ packed, const u64* __restrict__ off,
                                                              const u64* __restrict__ packedOff, u64 nReads, SeederParams P,
                                                              const SeedTables* __restrict__ gT, u64* __restrict__ synBuf,
                                                              unsigned* __restrict__ synCount, const unsigned char* __restrict__ dup,
                                                              const u64* __restrict__ endOff, const char* __restrict__ reads) {
    constexpr int W = K - S + 1;
    static_assert(K >= 8 && K <= 20 && S >= 8 && S < K && (K - S + 1) % 4 == 0, "word history holds 20 bases; the block length must be a multiple of 4");
    extern __shared__ __align__(16) unsigned char smemRaw[];
    SeedTables* sT = reinterpret_cast<SeedTables*>(smemRaw);
    unsigned char* sLut = smemRaw + sizeof(SeedTables);
    // (outgoing base, incoming base) pair tables: one look-up and one XOR per rolling hash instead of two and two
    // [4][kPairStride]: fk, rk, fs, rs; entry (o, nw) at 12 o + nw: the sixteen ACGT x ACGT pairs land in sixteen different
    // bank pairs (12 o mod 16 = 0, 12, 8, 4), so a warp's 64-bit look-ups are conflict-free unless ambiguous bases are involved
    u64* sPair = reinterpret_cast<u64*>(smemRaw + sizeof(SeedTables) + 256);
    for (int i = threadIdx.x; i < (int)(sizeof(SeedTables) / 8); i += blockDim.x)
        reinterpret_cast<u64*>(sT)[i] = reinterpret_cast<const u64*>(gT)[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        sLut[i] = (unsigned char)baseCode((unsigned char)i);
        const int tb = i >> 6, o = (i >> 3) & 7, nw = i & 7;
        sPair[tb * kPairStride + o * 12 + nw] = tb == 0 ? gT->fwdOldK[o] ^ gT->fwdNew[nw] : tb == 1 ? gT->revOld[o] ^ gT->revNewK[nw]
                                              : tb == 2 ? gT->fwdOldS[o] ^ gT->fwdNew[nw] : gT->revOld[o] ^ gT->revNewS[nw];
    }
    __syncthreads();
    const SeedTables& T = *sT;
    // F / suffix-min rings of the two strands: every index below is a compile-time constant (the block loop is fully unrolled), so
    // the rings live in registers and the only shared-memory traffic left is the 8-entry constant tables
    u64 rF[W], rR[W];
    const unsigned lane = threadIdx.x & 31u;
    const u64 warpsTotal = (u64)gridDim.x * (blockDim.x >> 5);
    const u64 warpId = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (u64 r0 = warpId * 32; r0 < nReads; r0 += warpsTotal * 32) {
        const u64 r = r0 + lane;
        const bool valid = r < nReads;
        const u64 b = valid ? off[r] : 0;
        int L = valid ? (int)((endOff ? endOff[r] : off[r + 1]) - b) : 0;
        if (L < K) L = 0;
        if (dup && valid && dup[r]) L = 0;   // --dedup: a byte-identical read was seen before
        const u64 pOff = valid ? packedOff[r] : 0;
        const uint4* __restrict__ src = packed + pOff;
        const char* rbase = ASCII ? reads + (b & ~3ULL) : nullptr;   // 4-byte aligned base of the read, rshift = its misalignment
        const unsigned rshift = (unsigned)(b & 3ULL);
        u64* __restrict__ dst = synBuf + pOff * 32;
        int maxL = L, minL = L;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { maxL = max(maxL, __shfl_xor_sync(0xffffffffu, maxL, d)); minL = min(minL, __shfl_xor_sync(0xffffffffu, minL, d)); }

        // The kernel is bound by the integer ALU pipe, so the loop below is written to need as few integer instructions per base
        // as possible: the bases that leave the k-mer / s-mer windows come from lagged copies of the packed words (one AND + one
        // shift each, no base history to maintain), every 64-bit minimum doubles as the comparison the syncmer test needs, and
        // the trim / ambiguity / window-complete conditions are one interval test on the read position.
        u64 fk = 0, rk = 0, fs = 0, rs = 0, preF = kEmptyKey, preR = kEmptyKey, firstF = 0, firstR = 0;
        unsigned maskF = 0, maskR = 0, cnt = 0;
        // the bases travel as one code byte each, four to a word: cur = bases 4n .. 4n+3, hN = the word N words back ("ambiguous"
        // before the read); combK / combS hold 12 * outgoing + incoming for the four positions of the current word
        unsigned cur = 0x04040404u, h1 = 0x04040404u, h2 = 0x04040404u, h3 = 0x04040404u, h4 = 0x04040404u, h5 = 0x04040404u;
        unsigned combK = 0, combS = 0;
        uint4 v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        int iLo = P.trimStart + K - 1;           // a window ending at base i is reported iff iLo <= i <= iHi:
        const int iHi = L - P.trimEnd - 1;       //   complete, inside the trimmed range, no ambiguous base in it (iLo moves past those)
#pragma unroll
        for (int q = 0; q < W; ++q) { rF[q] = 0; rR[q] = 0; }

        // base i enters: rolling hashes of the k-mer and s-mer ending at i, both strands (seeding.cpp:147-195).  ph = i & 3 is a
        // compile-time constant inside the unrolled block (W is a multiple of 4), so the word bookkeeping below is resolved statically
        // and every per-base index is ONE byte permute: with 4-bit codes the three shift+mask pairs per base were 6 of ~55 ALU-pipe
        // instructions.  FULL (compile time): base i lies inside every lane's read, so the `i < L` / `i < maxL` tests are dropped
        auto fetchRoll = [&](int i, int ph, auto full) {
            constexpr bool FULL = decltype(full)::value;
            if (ph == 0) {   // next four bases
                h5 = h4; h4 = h3; h3 = h2; h2 = h1; h1 = cur;
                if (ASCII) {
                    if (FULL || i < L) {   // 4 bytes from an arbitrary byte address: two aligned words, one funnel shift, four table look-ups
                        const unsigned a = rshift + (unsigned)i;
                        const unsigned* wp = reinterpret_cast<const unsigned*>(rbase + (a & ~3u));
                        const unsigned x0 = __ldg(wp), x1 = __ldg(wp + 1);
                        const unsigned bts = __funnelshift_r(x0, x1, (a & 3u) * 8u);
                        unsigned cw = sLut[bts >> 24];
                        cw = cw * 256u + sLut[(bts >> 16) & 0xFFu];
                        cw = cw * 256u + sLut[(bts >> 8) & 0xFFu];
                        cw = cw * 256u + sLut[bts & 0xFFu];
                        cur = cw;
                    }
                } else {   // 4-bit codes from `packed` (list utilities): spread four nibbles to bytes
                    if ((i & 31) == 0) { if (FULL || i < L) v = src[i >> 5]; }
                    const int wsel = (i >> 3) & 3;
                    const unsigned w32 = wsel == 0 ? v.x : wsel == 1 ? v.y : wsel == 2 ? v.z : v.w;
                    const unsigned n16 = (w32 >> (16 * ((i >> 2) & 1))) & 0xFFFFu;
                    cur = (((n16 & 0xF000u) << 12) | ((n16 & 0x0F00u) << 8) | ((n16 & 0x00F0u) << 4) | (n16 & 0x000Fu)) & 0x07070707u;
                }
                constexpr int KA = K / 4, KB = K % 4, SA = S / 4, SB = S % 4;
                const unsigned hist[7] = {cur, h1, h2, h3, h4, h5, 0x04040404u};
                // byte j of the lag word = base 4n + j - K: four consecutive bytes of (hist[KA+1], hist[KA]) starting at byte 4 - KB
                constexpr unsigned selK = (4 - KB) | ((5 - KB) << 4) | ((6 - KB) << 8) | ((7 - KB) << 12);
                constexpr unsigned selS = (4 - SB) | ((5 - SB) << 4) | ((6 - SB) << 8) | ((7 - SB) << 12);
                const unsigned lagK = KB ? __byte_perm(hist[KA + 1], hist[KA], selK) : hist[KA];
                const unsigned lagS = SB ? __byte_perm(hist[SA + 1], hist[SA], selS) : hist[SA];
                combK = lagK * 12u + cur;   // bytewise: codes are <= 7, 12 * 7 + 7 < 256
                combS = lagS * 12u + cur;
            }
            if (FULL || i < L) {
                const unsigned pk = __byte_perm(combK, 0u, 0x4440u + (unsigned)ph), ps = __byte_perm(combS, 0u, 0x4440u + (unsigned)ph);
                fk = rol1(fk) ^ sPair[pk];
                rk = ror1(rk) ^ sPair[kPairStride + pk];
                fs = rol1(fs) ^ sPair[2 * kPairStride + ps];
                rs = ror1(rs) ^ sPair[3 * kPairStride + ps];
                if (cur & (0x04u << (8 * ph))) iLo = max(iLo, i + K);   // codes >= 4 are ambiguous
            }
        };
        // prologue: the first S-1 bases only feed the rolling hashes
#pragma unroll 1
        for (int i = 0; i < S - 1 && i < maxL; ++i) fetchRoll(i, i & 3, std::false_type{});
        // one block of W s-mers; s-mer index q = i - (S-1), slot j = q mod W
        auto block = [&](int q0, auto full) {
            constexpr bool FULL = decltype(full)::value;
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const int i = q0 + j + S - 1;
                if (FULL || i < maxL) {   // warp-uniform
                    fetchRoll(i, (j + S - 1) & 3, full);   // q0 is a multiple of W, W of 4
                    if (FULL || i < L) {
                        // running minimum of the current block; leF: the newest s-mer attains it
                        bool leF = true, leR = true;
                        if (j == 0) { preF = fs; preR = rs; firstF = fs; firstR = rs; }
                        else { leF = fs <= preF; leR = rs <= preR; preF = leF ? fs : preF; preR = leR ? rs : preR; }
                        // closed syncmer, t == 0: the OLDEST or the NEWEST s-mer of the window attains the window minimum
                        const int pslot = (j + 1 == W) ? 0 : j + 1;
                        bool fsyn, rsyn;
                        if (pslot == 0) { fsyn = leF || firstF == preF; rsyn = leR || firstR == preR; }
                        else {
                            const u64 sf = rF[pslot], sr = rR[pslot];   // suffix minima of the previous block from the oldest s-mer on
                            // (mask & constant) != 0 is one LOP3 with a predicate result; (mask >> pslot) & 1 compiled to three instructions
                            fsyn = (leF && fs <= sf) || (maskBit(maskF, 1u << pslot) && sf <= preF);
                            rsyn = (leR && rs <= sr) || (maskBit(maskR, 1u << pslot) && sr <= preR);
                        }
                        if ((fsyn || rsyn) && i >= iLo && (NOTRIM || i <= iHi) && fk != rk) {
                            dst[cnt] = umin64(fk, rk);
                            ++cnt;
                        }
                        rF[j] = fs; rR[j] = rs;
                        if (j == W - 1) {   // block complete: in-place suffix minima + "is its own suffix minimum" masks
                            // the newest s-mer (slot W-1, = fs / rs) is its own suffix minimum; slot 0 is never looked up (a window that
                            // starts there is the block itself)
                            u64 a = fs, bb = rs; unsigned mF = 1u << (W - 1), mR = 1u << (W - 1);
#pragma unroll
                            for (int q = W - 2; q >= 1; --q) {
                                const u64 x = rF[q], y = rR[q];
                                const bool ia = x <= a, ib = y <= bb;
                                a = ia ? x : a; bb = ib ? y : bb;
                                if (ia) mF |= 1u << q;
                                if (ib) mR |= 1u << q;
                                rF[q] = a; rR[q] = bb;
                            }
                            maskF = mF; maskR = mR;
                        }
                    }
                }
            }
        };
        // main loop: blocks that lie inside every lane's read take the copy without the per-base length tests
#pragma unroll 1
        for (int q0 = 0; q0 + S - 1 < maxL; q0 += W) {
            if (q0 + W + S - 1 <= minL) block(q0, std::true_type{});
            else block(q0, std::false_type{});
        }
        if (valid) synCount[r] = cnt;
    }
}

// ---- syncmers_rank: the s = 8 closed-syncmer kernel on hash RANKS (pm_logic.cuh "closed syncmers with s = 8 by RANK") -------------
// Same output as syncmers_fast<K, 8>, about half the integer work per base: the two rolling s-mer hashes and their 64-bit sliding
// minima are replaced by two look-ups in a 128 KB rank table in shared memory (brought in by one TMA bulk copy per block while the
// warps set up their reads) and a sliding minimum over both strands' ranks at once (one VIMNMX.U16x2 per comparison, which also
// returns the two per-strand "<=" predicates).  The rolling k-mer hashes -- the values that are reported -- are unchanged.
// One persistent block per SM (the table takes 128 of its 227 KB of shared memory).
__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// STAGE (ASCII input, reads of at most kStageMaxLen bases -- the host guarantees it): the bytes of a warp's 32 consecutive reads are one
// contiguous span of the sample, brought into the warp's own staging buffer by ONE TMA bulk copy per group (coalesced, no L1 wavefronts, no
// sector fetched twice) instead of 32 lanes each pulling 4 bytes at a time from its own line (18 L1 wavefronts per request, a third of the
// kernel's L1 data-pipe load, and 245 MB of DRAM reads for 150 MB of bases); the lanes then read their words from shared memory.
constexpr int kStageMaxLen = 151;
constexpr unsigned kStageBytes = 32u * kStageMaxLen + 32u;   // span of 32 reads + alignment of both ends to 16 bytes
constexpr unsigned kStagePitch = kStageBytes + 16u;          // + the word past the last one (funnel shift)
template <int K, bool ASCII, bool NOTRIM, int THREADS, bool STAGE = false>
__global__ void __launch_bounds__(THREADS, 1) syncmers_rank(const uint4* __restrict__ packed, const u64* __restrict__ off,
                                                                 const u64* __restrict__ packedOff, u64 nReads, SeederParams P,
                                                                 const SeedTables* __restrict__ gT, u64* __restrict__ synBuf,
                                                                 unsigned* __restrict__ synCount, const unsigned char* __restrict__ dup,
                                                                 const u64* __restrict__ endOff, const char* __restrict__ reads) {
    constexpr int S = kRankS, W = K - S + 1;
    static_assert(K >= 12 && K <= 20 && W % 4 == 0 && W <= 12, "word history holds 20 bases; the block length must be a multiple of 4");
    extern __shared__ __align__(128) unsigned char smemRaw[];
    unsigned char* sRank = smemRaw;                                              // u16[65536]
    u64* sPair = reinterpret_cast<u64*>(smemRaw + kRankEntries * 2);             // [2][kPairStride]: fk, rk (outgoing, incoming) pair tables
    unsigned char* sLut = smemRaw + kRankEntries * 2 + 2 * kPairStride * 8;      // ASCII -> base code
    static_assert(!STAGE || ASCII, "staging is for ASCII input");
    unsigned char* sStage = smemRaw + kRankEntries * 2 + 2 * kPairStride * 8 + 256 + (size_t)(threadIdx.x >> 5) * kStagePitch;   // this warp's
    __shared__ __align__(8) unsigned long long sBar;
    __shared__ __align__(8) unsigned long long sWarpBar[THREADS / 32];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(&sBar)));
        if (STAGE)
            for (int w = 0; w < THREADS / 32; ++w) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(&sWarpBar[w])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        sLut[i] = (unsigned char)baseCode((unsigned char)i);
        if (i < 128) {
            const int tb = i >> 6, o = (i >> 3) & 7, nw = i & 7;
            sPair[tb * kPairStride + o * 12 + nw] = tb == 0 ? gT->fwdOldK[o] ^ gT->fwdNew[nw] : gT->revOld[o] ^ gT->revNewK[nw];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {   // the rank table sits right behind the SeedTables (pm_api.cu: seedTableBlob): four 32 KB bulk copies, one barrier
        const unsigned bar = smemAddr(&sBar);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kRankEntries * 2u) : "memory");
        const unsigned char* src = reinterpret_cast<const unsigned char*>(gT + 1);
#pragma unroll
        for (unsigned c = 0; c < 4; ++c)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(sRank + c * 32768u)),
                         "l"(src + c * 32768u), "r"(32768u), "r"(bar)
                         : "memory");
    }
    bool tableReady = false;
    unsigned stagePhase = 0;   // parity of this warp's staging barrier
    RankWindow<W> win;
    const unsigned lane = threadIdx.x & 31u;
    const u64 warpsTotal = (u64)gridDim.x * (blockDim.x >> 5);
    const u64 warpId = (u64)(threadIdx.x >> 5) * gridDim.x + blockIdx.x;   // consecutive groups of 32 reads go to different SMs
    for (u64 r0 = warpId * 32; r0 < nReads; r0 += warpsTotal * 32) {
        const u64 r = r0 + lane;
        const bool valid = r < nReads;
        const u64 b = valid ? off[r] : 0;
        int L = valid ? (int)((endOff ? endOff[r] : off[r + 1]) - b) : 0;
        if (L < K) L = 0;
        if (dup && valid && dup[r]) L = 0;   // --dedup: a byte-identical read was seen before
        const u64 pOff = valid ? packedOff[r] : 0;
        const uint4* __restrict__ src = packed + pOff;
        const char* rbase = ASCII ? reads + (b & ~3ULL) : nullptr;   // 4-byte aligned base of the read, rshift = its misalignment
        const unsigned rshift = (unsigned)(b & 3ULL);
        u64* __restrict__ dst = synBuf + pOff * 32;
        int maxL = L, minL = L;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { maxL = max(maxL, __shfl_xor_sync(0xffffffffu, maxL, d)); minL = min(minL, __shfl_xor_sync(0xffffffffu, minL, d)); }
        unsigned sOff = 0;   // STAGE: byte offset of the lane's read inside the staging buffer
        if (STAGE) {
            const u64 gB = __shfl_sync(0xffffffffu, b, 0);                      // lane 0 is always a valid read
            const u64 rEnd = r0 + 32 < nReads ? r0 + 32 : nReads;
            const u64 gE = off[rEnd];
            const u64 g0 = gB & ~15ULL;
            const unsigned bytes = (unsigned)(((gE + 15ULL) & ~15ULL) - g0);  // the reads buffer ends with 64 spare bytes
            if (bytes > kStageBytes) __trap();                                  // the launcher checked the longest read
            sOff = valid ? (unsigned)(b - g0) : 0u;
            __syncwarp();   // every lane has consumed its words of the previous group
            if (bytes) {
                const unsigned wbar = smemAddr(&sWarpBar[threadIdx.x >> 5]);
                if (lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wbar), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(sStage)),
                                 "l"(reads + g0), "r"(bytes), "r"(wbar)
                                 : "memory");
                }
                unsigned ok;
                do {
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(wbar), "r"(stagePhase) : "memory");
                } while (!ok);
                stagePhase ^= 1u;
            }
        }
        if (!tableReady) {   // first group of the warp: everything above overlapped the bulk copy
            unsigned ok;
            do {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smemAddr(&sBar)) : "memory");
            } while (!ok);
            tableReady = true;
        }

        u64 fk = 0, rk = 0;
        unsigned HF = 0, HR = 0, cnt = 0;
        // the bases travel as one code byte each, four to a word: cur = bases 4n .. 4n+3, hN = the word N words back ("ambiguous"
        // before the read); combK holds 12 * outgoing + incoming for the four positions of the current word
        unsigned cur = 0x04040404u, h1 = 0x04040404u, h2 = 0x04040404u, h3 = 0x04040404u, h4 = 0x04040404u, h5 = 0x04040404u;
        unsigned combK = 0;
        bool ambWord = false;
        uint4 v4 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        int iLo = P.trimStart + K - 1;           // a window ending at base i is reported iff iLo <= i <= iHi:
        const int iHi = L - P.trimEnd - 1;       //   complete, inside the trimmed range, no ambiguous base in it (iLo moves past those)
        win.reset();

        // the next four bases (i .. i+3, i a multiple of 4) as one code byte each
        auto loadWord = [&](int i) -> unsigned {
            if (ASCII) {   // 4 bytes from an arbitrary byte address: two aligned words, one funnel shift, four table look-ups
                const unsigned a = (STAGE ? sOff : rshift) + (unsigned)i;
                unsigned x0, x1;
                if (STAGE) {
                    const unsigned* wp = reinterpret_cast<const unsigned*>(sStage + (a & ~3u));
                    x0 = wp[0]; x1 = wp[1];
                } else {
                    const unsigned* wp = reinterpret_cast<const unsigned*>(rbase + (a & ~3u));
                    x0 = __ldg(wp); x1 = __ldg(wp + 1);
                }
                const unsigned bts = __funnelshift_r(x0, x1, (a & 3u) * 8u);
                unsigned cw = sLut[bts >> 24];
                cw = cw * 256u + sLut[(bts >> 16) & 0xFFu];
                cw = cw * 256u + sLut[(bts >> 8) & 0xFFu];
                cw = cw * 256u + sLut[bts & 0xFFu];
                return cw;
            }
            // 4-bit codes from `packed`: spread four nibbles to bytes
            if ((i & 31) == 0) v4 = src[i >> 5];
            const int wsel = (i >> 3) & 3;
            const unsigned w32 = wsel == 0 ? v4.x : wsel == 1 ? v4.y : wsel == 2 ? v4.z : v4.w;
            const unsigned n16 = (w32 >> (16 * ((i >> 2) & 1))) & 0xFFFFu;
            return (((n16 & 0xF000u) << 12) | ((n16 & 0x0F00u) << 8) | ((n16 & 0x00F0u) << 4) | (n16 & 0x000Fu)) & 0x07070707u;
        };
        // word cw becomes the current one: word history for the k-mer's outgoing bases, 2-bit histories for the s-mer ranks
        auto pushWord = [&](unsigned cw, auto clean) {
            constexpr bool CLEAN = decltype(clean)::value;
            h5 = h4; h4 = h3; h3 = h2; h2 = h1; h1 = cur; cur = cw;
            constexpr int KA = K / 4, KB = K % 4;
            const unsigned hist[7] = {cur, h1, h2, h3, h4, h5, 0x04040404u};
            // byte j of the lag word = base 4n + j - K: four consecutive bytes of (hist[KA+1], hist[KA]) starting at byte 4 - KB
            constexpr unsigned selK = (4 - KB) | ((5 - KB) << 4) | ((6 - KB) << 8) | ((7 - KB) << 12);
            const unsigned lagK = KB ? __byte_perm(hist[KA + 1], hist[KA], selK) : hist[KA];
            combK = lagK * 12u + cur;   // bytewise: codes are <= 7, 12 * 7 + 7 < 256
            rankPushWord(cur, HF, HR);
            if (!CLEAN) ambWord = (cur & 0x04040404u) != 0u;
        };
        // base i (phase ph of the current word) enters the k-mer's rolling hashes on both strands (seeding.cpp:147-195)
        auto roll = [&](int i, int ph, auto clean) {
            constexpr bool CLEAN = decltype(clean)::value;
            const unsigned pk = __byte_perm(combK, 0u, 0x4440u + (unsigned)ph);
            fk = rol1(fk) ^ sPair[pk];
            rk = ror1(rk) ^ sPair[kPairStride + pk];
            if (!CLEAN) { if (ambWord) { if (cur & (0x04u << (8 * ph))) iLo = max(iLo, i + K); } }   // codes >= 4 are ambiguous
        };
        // prologue: the first S-1 bases only feed the rolling hashes and the histories
#pragma unroll 1
        for (int i = 0; i < S - 1 && i < maxL; ++i) {
            if ((i & 3) == 0) { unsigned cw = cur; if (i < L) cw = loadWord(i); pushWord(cw, std::false_type{}); }
            if (i < L) roll(i, i & 3, std::false_type{});
        }
        // one block of W s-mers; s-mer index q = i - (S-1), slot j = q mod W.  FULL: the block lies inside every lane's read (no length
        // tests) and its words were loaded up front (pre); CLEAN: none of them -- nor the word the block starts in -- has an ambiguous base
        auto block = [&](int q0, auto full, auto clean, const unsigned* pre) {
            constexpr bool FULL = decltype(full)::value;
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const int i = q0 + j + S - 1;
                if (FULL || i < maxL) {   // warp-uniform
                    const int ph = (j + S - 1) & 3;   // q0 is a multiple of W, W of 4
                    if (ph == 0) {
                        unsigned cw = cur;
                        if (FULL) cw = pre[(j - 1) / 4];
                        else if (i < L) cw = loadWord(i);
                        pushWord(cw, clean);
                    }
                    if (FULL || i < L) {
                        roll(i, ph, clean);
                        const unsigned rf = *reinterpret_cast<const unsigned short*>(sRank + rankAddrF(HF, ph));
                        const unsigned rr = *reinterpret_cast<const unsigned short*>(sRank + rankAddrR(HR, ph));
                        const bool syn = win.step(j, __byte_perm(rf, rr, 0x5410u));
                        if (syn && i >= iLo && (NOTRIM || i <= iHi) && fk != rk) {
                            dst[cnt] = umin64(fk, rk);
                            ++cnt;
                        }
                    }
                }
            }
        };
        // main loop: blocks that lie inside every lane's read take the copies without the per-base length tests
#pragma unroll 1
        for (int q0 = 0; q0 + S - 1 < maxL; q0 += W) {
            if (q0 + W + S - 1 <= minL) {
                unsigned pre[W / 4];
                bool amb = ambWord;
#pragma unroll
                for (int t = 0; t < W / 4; ++t) { pre[t] = loadWord(q0 + S + 4 * t); amb = amb || (pre[t] & 0x04040404u) != 0u; }
                if (!__any_sync(0xffffffffu, amb)) block(q0, std::true_type{}, std::true_type{}, pre);
                else block(q0, std::true_type{}, std::false_type{}, pre);
            } else block(q0, std::false_type{}, std::false_type{}, nullptr);
        }
        if (valid) synCount[r] = cnt;
    }
}

// placement.cpp:1625-1682: one warp per read; lane j builds the k-min-mer of syncmers j..j+l-1 in closed form
//   Fw = XOR_w rol(h[j+w], k*(l-1-w)),  Rw = XOR_w rol(h[j+w], k*w),  seed = min(Fw,Rw) unless Fw == Rw     (l > 1)
//   seed = h[j]                                                                                                 (l <= 1)
// MODE 0: count in the open-addressing table; MODE 2: ordered per-read list at outHash[winOff[r] ...]
// KT/LT > 0: k and l are compile-time constants (rotations become immediates); KT == 0: runtime k, l.
template <int MODE, int KT, int LT>
__global__ void __launch_bounds__(256) seeds_from_syncmers(const u64* __restrict__ synBuf, const unsigned* __restrict__ synCount,
                                                           const u64* __restrict__ packedOff, const u64* __restrict__ winOff, u64 nReads,
                                                           int kRt, int lRt, TableSlot* table, u64 mask, SampleAcc* acc,
                                                           u64* outHash, u64* outCount, cudaTextureObject_t tableTex,
                                                           const unsigned char* __restrict__ pass = nullptr) {
    // pass (--min-seed-quality only, same indexing as synBuf): a window counts only when all of its syncmers pass
    const int k = KT > 0 ? KT : kRt, l = KT > 0 ? LT : lRt;
    const unsigned lane = threadIdx.x & 31u;
    const u64 warpsTotal = (u64)gridDim.x * (blockDim.x >> 5);
    if (MODE == 0) {
        // counting: a warp works on TWO reads at a time, two seeds per lane and read, so that four independent chains (list loads ->
        // seed -> first probe -> update) are in flight per lane: the kernel is a sequence of L2 round trips per read otherwise
        auto seedOf = [&](const u64* __restrict__ h, const unsigned char* __restrict__ pp, int nSeeds, int j, u64& seed) -> bool {
            if (j >= nSeeds) return false;
            if (pp) { for (int w = 0; w < (l <= 1 ? 1 : l); ++w) if (!pp[j + w]) return false; }
            if (l <= 1) { seed = h[j]; return true; }
            u64 fw = 0, rw = 0;
            if (KT > 0) {
#pragma unroll
                for (int w = 0; w < (LT > 0 ? LT : 1); ++w) {
                    const u64 x = h[j + w];
                    fw ^= rol64(x, (unsigned)((KT * (LT - 1 - w)) & 63));
                    rw ^= rol64(x, (unsigned)((KT * w) & 63));
                }
            } else {
                for (int w = 0; w < l; ++w) {
                    const u64 x = h[j + w];
                    fw ^= rol64(x, (unsigned)(k * (l - 1 - w)));
                    rw ^= rol64(x, (unsigned)(k * w));
                }
            }
            seed = umin64(fw, rw);
            return fw != rw;
        };
        for (u64 rA = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rA < nReads; rA += 2 * warpsTotal) {
            const u64 rB = rA + warpsTotal;
            const bool isB = rB < nReads;
            const int nA = (int)__ldg(&synCount[rA]), nB = isB ? (int)__ldg(&synCount[rB]) : 0;
            const u64 oA = __ldg(&packedOff[rA]) * 32, oB = isB ? __ldg(&packedOff[rB]) * 32 : 0;
            const u64* __restrict__ hA = synBuf + oA; const u64* __restrict__ hB = synBuf + oB;
            const unsigned char* __restrict__ pA = pass ? pass + oA : nullptr; const unsigned char* __restrict__ pB = pass ? pass + oB : nullptr;
            const int sA = l <= 1 ? nA : (nA >= l ? nA - l + 1 : 0), sB = l <= 1 ? nB : (nB >= l ? nB - l + 1 : 0);
            for (int j0 = 0; j0 < max(sA, sB); j0 += 64) {
                u64 sd[4] = {0, 0, 0, 0}, ps[4], ky[4]; bool has[4];
                has[0] = seedOf(hA, pA, sA, j0 + (int)lane, sd[0]); has[1] = seedOf(hB, pB, sB, j0 + (int)lane, sd[1]);
                has[2] = seedOf(hA, pA, sA, j0 + 32 + (int)lane, sd[2]); has[3] = seedOf(hB, pB, sB, j0 + 32 + (int)lane, sd[3]);
                // first probe through the texture path (keys are write-once: a stale EMPTY only sends the seed to the CAS path)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    ps[q] = mixKey(sd[q]) & mask; ky[q] = 0;
                    if (has[q]) { const uint4 t = tex1Dfetch<uint4>(tableTex, (int)ps[q]); ky[q] = (u64)t.x | ((u64)t.y << 32); }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (has[q]) { if (ky[q] == sd[q] && sd[q] != kEmptyKey) atomicAdd(&table[ps[q]].count, 1u); else tableInsert(table, mask, sd[q], 1u, acc); }
            }
        }
        return;
    }
    for (u64 r = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nReads; r += warpsTotal) {
        const int n = (int)synCount[r];
        const u64* __restrict__ h = synBuf + packedOff[r] * 32;
        const int nSeeds = l <= 1 ? n : (n >= l ? n - l + 1 : 0);
        u64 written = 0;
        const unsigned char* __restrict__ pp = pass ? pass + packedOff[r] * 32 : nullptr;
        auto seedAt = [&](int j, u64& seed) -> bool {
            if (j >= nSeeds) return false;
            if (pp) { for (int w = 0; w < (l <= 1 ? 1 : l); ++w) if (!pp[j + w]) return false; }
            if (l <= 1) { seed = h[j]; return true; }
            u64 fw = 0, rw = 0;
            if (KT > 0) {
#pragma unroll
                for (int w = 0; w < (LT > 0 ? LT : 1); ++w) {
                    const u64 x = h[j + w];
                    fw ^= rol64(x, (unsigned)((KT * (LT - 1 - w)) & 63));
                    rw ^= rol64(x, (unsigned)((KT * w) & 63));
                }
            } else {
                for (int w = 0; w < l; ++w) {
                    const u64 x = h[j + w];
                    fw ^= rol64(x, (unsigned)(k * (l - 1 - w)));
                    rw ^= rol64(x, (unsigned)(k * w));
                }
            }
            seed = umin64(fw, rw);
            return fw != rw;
        };
        for (int j0 = 0; j0 < nSeeds; j0 += 64) {
            // two seeds per lane so that both table probes are in flight together
            u64 s0 = 0, s1 = 0;
            const bool h0 = seedAt(j0 + (int)lane, s0), h1 = seedAt(j0 + 32 + (int)lane, s1);
            if (MODE == 0) {
                u64 k0 = 0, k1 = 0, p0 = 0, p1 = 0;
                // first probe through the texture path (keys are write-once: a stale EMPTY only sends the seed to the CAS path)
                if (h0) { p0 = mixKey(s0) & mask; const uint4 t = tex1Dfetch<uint4>(tableTex, (int)p0); k0 = (u64)t.x | ((u64)t.y << 32); }
                if (h1) { p1 = mixKey(s1) & mask; const uint4 t = tex1Dfetch<uint4>(tableTex, (int)p1); k1 = (u64)t.x | ((u64)t.y << 32); }
                if (h0) { if (k0 == s0) atomicAdd(&table[p0].count, 1u); else tableInsert(table, mask, s0, 1u, acc); }
                if (h1) { if (k1 == s1) atomicAdd(&table[p1].count, 1u); else tableInsert(table, mask, s1, 1u, acc); }
            } else {
                unsigned m = __ballot_sync(0xffffffffu, h0);
                if (h0) outHash[winOff[r] + written + __popc(m & ((1u << lane) - 1u))] = s0;
                written += __popc(m);
                m = __ballot_sync(0xffffffffu, h1);
                if (h1) outHash[winOff[r] + written + __popc(m & ((1u << lane) - 1u))] = s1;
                written += __popc(m);
            }
        }
        if (MODE != 0 && lane == 0) outCount[r] = written;
    }
}
// count_seeds: the table-insertion half of seeding for whole samples.  A warp takes 32 reads at a time (lane q holds the list of
// read r0 + q), numbers their seeds consecutively and works through them 128 at a time, so that every lane has four independent
// chains (syncmer loads -> seed -> first probe) in flight: the kernel is bound by memory latency, not by bandwidth or issue.
// AGG (whole samples in one launch): one 1024-thread block per SM keeps an open-addressing table of (seed, count) in shared memory
// (kAggSlots entries, 64-bit CAS to claim a key, 32-bit atomic add to count) and flushes it into the global table at the end.
// At sequencing depth most seed instances repeat a few thousand genome seeds, and the kernel is bound by L2 request rate (probe +
// red.add per instance, the hot keys concentrated on few slices): those instances now stay on the SM.  Seeds that find their
// probe window taken go to the global table directly, so the result is the same multiset of (seed, count) either way.
constexpr int kAggSlots = 16384;
// below this many reads per launch the flush of 148 shared-memory tables costs more than the pre-aggregation saves (tuning override: PM_AGG_MIN_READS)
static u64 aggMinReads() {
    static const u64 v = [] { const char* e = std::getenv("PM_AGG_MIN_READS"); return e ? (u64)std::strtoull(e, nullptr, 10) : (u64)(1u << 18); }();
    return v;
}
__device__ __forceinline__ bool aggAdd(u64* __restrict__ sKey, u32* __restrict__ sCnt, u64 s, u64 m) {
    u32 i = (u32)(m >> 40) & (kAggSlots - 1);
#pragma unroll
    for (int p = 0; p < 2; ++p) {   // two slots: junk (error) seeds fill the table early and every further probe is issued for a few lanes only (4: 442 us, 2: 410 us, 1: 401 us but a quarter of the hot keys would lose their slot)
        u64 kk = sKey[i];
        if (kk == kEmptyKey) {
            kk = atomicCAS(reinterpret_cast<unsigned long long*>(&sKey[i]), (unsigned long long)kEmptyKey, (unsigned long long)s);
            if (kk == kEmptyKey) kk = s;
        }
        if (kk == s) { atomicAdd(&sCnt[i], 1u); return true; }
        i = (i + 1) & (kAggSlots - 1);
    }
    return false;
}
template <int KT, int LT, bool AGG>
__global__ void __launch_bounds__(AGG ? 1024 : 256) count_seeds(const u64* __restrict__ synBuf, const unsigned* __restrict__ synCount,
                                                   const u64* __restrict__ packedOff, u64 nReads, int kRt, int lRt, TableSlot* table,
                                                   u64 mask, SampleAcc* acc, cudaTextureObject_t tableTex) {
    extern __shared__ __align__(16) unsigned char aggRaw[];
    u64* sKey = reinterpret_cast<u64*>(aggRaw);
    u32* sCnt = reinterpret_cast<u32*>(aggRaw + (size_t)kAggSlots * sizeof(u64));
    if (AGG) {
        for (int i = threadIdx.x; i < kAggSlots; i += blockDim.x) { sKey[i] = kEmptyKey; sCnt[i] = 0; }
        __syncthreads();
    }
    const int k = KT > 0 ? KT : kRt, l = KT > 0 ? LT : lRt;
    const unsigned lane = threadIdx.x & 31u;
    const u64 warpsTotal = (u64)gridDim.x * (blockDim.x >> 5);
    for (u64 r0 = ((u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; r0 < nReads; r0 += warpsTotal * 32) {
        const u64 r = r0 + lane;
        const int n = r < nReads ? (int)__ldg(&synCount[r]) : 0;
        const u64 listBits = reinterpret_cast<u64>(synBuf + (r < nReads ? __ldg(&packedOff[r]) : 0) * 32);
        const int nS = l <= 1 ? n : (n >= l ? n - l + 1 : 0);
        int incl = nS;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += o; }
        const int E = incl - nS;                                   // seeds of the lanes before this one
        const int T = __shfl_sync(0xffffffffu, incl, 31);          // seeds of the 32 reads
        for (int i0 = 0; i0 < T; i0 += 128) {
            u64 sd[4], slot[4]; bool has[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + 32 * q + (int)lane;
                int lo = 0;                                        // last lane whose first seed number is <= i
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int Em = __shfl_sync(0xffffffffu, E, lo + step);
                    if (Em <= i) lo += step;
                }
                const int j = i - __shfl_sync(0xffffffffu, E, lo);
                const u64* __restrict__ h = reinterpret_cast<const u64*>(shflU64(listBits, lo));
                has[q] = i < T; sd[q] = 0;
                if (has[q]) {
                    if (l <= 1) sd[q] = __ldcs(h + j);
                    else {
                        u64 fw = 0, rw = 0;
                        if (KT > 0) {
#pragma unroll
                            for (int w = 0; w < (LT > 0 ? LT : 1); ++w) {
                                const u64 x = __ldcs(h + j + w);   // streaming: the lists must not push the table out of L2
                                fw ^= rol64(x, (unsigned)((KT * (LT - 1 - w)) & 63));
                                rw ^= rol64(x, (unsigned)((KT * w) & 63));
                            }
                        } else {
                            for (int w = 0; w < l; ++w) {
                                const u64 x = __ldg(h + j + w);
                                fw ^= rol64(x, (unsigned)(k * (l - 1 - w)));
                                rw ^= rol64(x, (unsigned)(k * w));
                            }
                        }
                        sd[q] = umin64(fw, rw);
                        has[q] = fw != rw;
                    }
                }
                slot[q] = mixKey(sd[q]);
            }
            if (AGG) {   // after ALL list loads of the round were issued: atomics are ordering points for the compiler
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (has[q] && sd[q] != kEmptyKey && aggAdd(sKey, sCnt, sd[q], slot[q])) has[q] = false;   // counted on the SM
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) slot[q] &= mask;
            u64 key[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {   // first probe through the texture path (keys are write-once: a stale EMPTY only sends the seed to the CAS path)
                key[q] = kEmptyKey;
                if (has[q]) { const uint4 t = tex1Dfetch<uint4>(tableTex, (int)slot[q]); key[q] = (u64)t.x | ((u64)t.y << 32); }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (has[q]) { if (key[q] == sd[q] && sd[q] != kEmptyKey) atomicAdd(&table[slot[q]].count, 1u); else tableInsert(table, mask, sd[q], 1u, acc); }
        }
    }
    if (AGG) {
        __syncthreads();
        for (int i = threadIdx.x; i < kAggSlots; i += blockDim.x) {
            const u32 c = sCnt[i];
            if (c) tableInsert(table, mask, sKey[i], c, acc);
        }
    }
}
// count_seeds_lane: the same counting with one LANE per read (as in the syncmer kernel): a lane walks its own list, keeps the last
// LT-1 syncmers in registers and loads one new syncmer per seed -- no search for "which read does seed i belong to", a third of the
// loads and a quarter of the instructions of the flat numbering above.  Four seeds per lane are formed before anything is probed.
// QUEUE (whole samples): seeds that miss the block's shared-memory table are not walked through the global table here -- a few lanes
// at a time, every one a dependent chain of L2 round trips that the whole warp waits for -- but appended to a miss queue in global
// memory (per-warp staging in shared memory, one reservation per ~100 entries); count_misses then inserts them with every lane busy.
constexpr int kWarpQueue = 128;   // staged misses per warp
template <int KT, int LT, bool AGG, bool QUEUE>
__global__ void __launch_bounds__(AGG ? 1024 : 256, 1) count_seeds_lane(const u64* __restrict__ synBuf, const unsigned* __restrict__ synCount,
                                                   const u64* __restrict__ packedOff, u64 nReads, TableSlot* table,
                                                   u64 mask, SampleAcc* acc, cudaTextureObject_t tableTex, u64* __restrict__ missQ, u64 missCap, int listPrefetch) {
    static_assert(LT == 1 || LT == 3, "lane-per-read counting is specialised for l = 1 and l = 3");
    static_assert(AGG || !QUEUE, "the miss queue belongs to the pre-aggregating variant");
    extern __shared__ __align__(16) unsigned char aggRaw[];
    u64* sKey = reinterpret_cast<u64*>(aggRaw);
    u32* sCnt = reinterpret_cast<u32*>(aggRaw + (size_t)kAggSlots * sizeof(u64));
    u64* wq = reinterpret_cast<u64*>(aggRaw + (size_t)kAggSlots * (sizeof(u64) + sizeof(u32))) + (size_t)(threadIdx.x >> 5) * kWarpQueue;
    unsigned wcnt = 0;   // warp-uniform
    if (AGG) {
        for (int i = threadIdx.x; i < kAggSlots; i += blockDim.x) { sKey[i] = kEmptyKey; sCnt[i] = 0; }
        __syncthreads();
    }
    const unsigned lane = threadIdx.x & 31u;
    auto flushQueue = [&]() {   // whole warp
        __syncwarp();
        unsigned base32 = 0;
        if (lane == 0) base32 = atomicAdd(&acc->missCount, wcnt);
        const u64 base = __shfl_sync(0xffffffffu, base32, 0);
        for (unsigned i = lane; i < wcnt; i += 32) {
            if (base + i < missCap) missQ[base + i] = wq[i];
            else tableInsert(table, mask, wq[i], 1u, acc);   // queue full: the direct way (count_misses stops at missCap)
        }
        __syncwarp();
        wcnt = 0;
    };
    const u64 warpsTotal = (u64)gridDim.x * (blockDim.x >> 5);
    for (u64 r0 = ((u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; r0 < nReads; r0 += warpsTotal * 32) {
        const u64 r = r0 + lane;
        const int n = r < nReads ? (int)__ldg(&synCount[r]) : 0;
        const u64* __restrict__ h = synBuf + (r < nReads ? __ldg(&packedOff[r]) : 0) * 32;
        const int nS = LT <= 1 ? n : (n >= LT ? n - LT + 1 : 0);
        int maxS = nS;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) maxS = max(maxS, __shfl_xor_sync(0xffffffffu, maxS, d));
        // the whole list of the lane's read is asked for at once (L2 prefetch, one request per 128-byte line): without it every round of
        // four seeds starts with a fresh 32-byte sector from DRAM, one exposed DRAM latency per round and lane
        if (listPrefetch)
            for (int e = 16; e < n; e += 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(h + e));
        u64 h0 = 0, h1 = 0;   // the two syncmers before the next one to load (l = 3)
        if (LT == 3 && nS > 0) { h0 = __ldg(h); h1 = __ldg(h + 1); }
        constexpr int R = QUEUE ? 8 : 4;   // seeds per lane and round: without the global walk there are registers for eight loads in flight
        for (int j0 = 0; j0 < maxS; j0 += R) {
            u64 x[R], sd[R], slot[R]; bool has[R];
#pragma unroll
            for (int q = 0; q < R; q += 2) {   // 16 bytes at a time (lists start on 256-byte boundaries, j0 + LT - 1 is even): every sector is fetched once
                has[q] = j0 + q < nS; has[q + 1] = j0 + q + 1 < nS;
                uint4 t = make_uint4(0u, 0u, 0u, 0u);
                if (has[q]) t = __ldg(reinterpret_cast<const uint4*>(h + j0 + q + (LT - 1)));   // the second half may lie past the list's end: unused then
                x[q] = (u64)t.x | ((u64)t.y << 32); x[q + 1] = (u64)t.z | ((u64)t.w << 32);
            }
#pragma unroll
            for (int q = 0; q < R; ++q) {
                if (LT == 1) sd[q] = x[q];
                else {
                    const u64 fw = rol64(h0, (unsigned)((KT * 2) & 63)) ^ rol64(h1, (unsigned)(KT & 63)) ^ x[q];
                    const u64 rw = h0 ^ rol64(h1, (unsigned)(KT & 63)) ^ rol64(x[q], (unsigned)((KT * 2) & 63));
                    sd[q] = umin64(fw, rw);
                    has[q] = has[q] && fw != rw;
                    h0 = h1; h1 = x[q];
                }
                slot[q] = mixKey(sd[q]);
            }
            if (AGG) {
#pragma unroll
                for (int q = 0; q < R; ++q)
                    if (has[q] && sd[q] != kEmptyKey && aggAdd(sKey, sCnt, sd[q], slot[q])) has[q] = false;   // counted on the SM
            }
            if (QUEUE) {
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const unsigned m = __ballot_sync(0xffffffffu, has[q]);
                    if (m) {
                        if (wcnt + (unsigned)__popc(m) > (unsigned)kWarpQueue) flushQueue();
                        if (has[q]) wq[wcnt + (unsigned)__popc(m & ((1u << lane) - 1u))] = sd[q];
                        wcnt += (unsigned)__popc(m);
                    }
                }
                continue;
            }
            u64 key[R];
#pragma unroll
            for (int q = 0; q < R; ++q) {
                slot[q] &= mask; key[q] = kEmptyKey;
                if (has[q]) { const uint4 t = tex1Dfetch<uint4>(tableTex, (int)slot[q]); key[q] = (u64)t.x | ((u64)t.y << 32); }
            }
            // new keys claim their slot with a CAS whose result takes a round trip to L2: issue the CASes of all four seeds before
            // looking at any result (one exposed latency per round instead of up to four); true collisions go the general way
            bool claim[R], slow[R]; unsigned long long was[R];
#pragma unroll
            for (int q = 0; q < R; ++q) {
                claim[q] = slow[q] = false;
                if (has[q]) {
                    if (sd[q] == kEmptyKey) slow[q] = true;
                    else if (key[q] == sd[q]) atomicAdd(&table[slot[q]].count, 1u);
                    else if (key[q] == kEmptyKey) claim[q] = true;
                    else slow[q] = true;
                }
            }
#pragma unroll
            for (int q = 0; q < R; ++q)
                if (claim[q]) was[q] = atomicCAS(reinterpret_cast<unsigned long long*>(&table[slot[q]].key), (unsigned long long)kEmptyKey, (unsigned long long)sd[q]);
#pragma unroll
            for (int q = 0; q < R; ++q)
                if (claim[q]) { if (was[q] == kEmptyKey || was[q] == sd[q]) atomicAdd(&table[slot[q]].count, 1u); else slow[q] = true; }
#pragma unroll
            for (int q = 0; q < R; ++q)
                if (slow[q]) tableInsert(table, mask, sd[q], 1u, acc);
        }
    }
    if (QUEUE) { if (wcnt) flushQueue(); }
    if (AGG) {   // flush: the block's (seed, count) pairs into the global table, four claims in flight per thread
        __syncthreads();
        for (int i0 = threadIdx.x * 4; i0 < kAggSlots; i0 += blockDim.x * 4) {
            u64 k4[4], sl[4], seen[4]; u32 c4[4]; unsigned long long was[4]; bool claim[4], slow[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { k4[q] = sKey[i0 + q]; c4[q] = sCnt[i0 + q]; sl[q] = mixKey(k4[q]) & mask; seen[q] = kEmptyKey; }
#pragma unroll
            for (int q = 0; q < 4; ++q) if (c4[q]) seen[q] = __ldcg(&table[sl[q]].key);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                claim[q] = slow[q] = false;
                if (c4[q]) {
                    if (seen[q] == k4[q]) atomicAdd(&table[sl[q]].count, c4[q]);
                    else if (seen[q] == kEmptyKey) claim[q] = true;
                    else slow[q] = true;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (claim[q]) was[q] = atomicCAS(reinterpret_cast<unsigned long long*>(&table[sl[q]].key), (unsigned long long)kEmptyKey, (unsigned long long)k4[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (claim[q]) { if (was[q] == kEmptyKey || was[q] == k4[q]) atomicAdd(&table[sl[q]].count, c4[q]); else slow[q] = true; }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (slow[q]) tableInsert(table, mask, k4[q], c4[q], acc);
        }
    }
}
// the miss queue of count_seeds_lane<.., QUEUE> into the global table: one seed per thread and turn, four turns in flight
// four seeds of one thread into the global table: first probes (texture path) of all four in flight, then the claims of the new keys, then
// the counts; true collisions go the general way (shared by count_misses and count_buckets)
__device__ __forceinline__ void insertFour(const u64 (&sd)[4], const bool (&has)[4], TableSlot* table, u64 mask, SampleAcc* acc, cudaTextureObject_t tableTex) {
    u64 slot[4], key[4]; bool claim[4], slow[4]; unsigned long long was[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        slot[q] = mixKey(sd[q]) & mask; key[q] = kEmptyKey;
        if (has[q]) { const uint4 t = tex1Dfetch<uint4>(tableTex, (int)slot[q]); key[q] = (u64)t.x | ((u64)t.y << 32); }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        claim[q] = slow[q] = false;
        if (has[q]) {
            if (sd[q] == kEmptyKey) slow[q] = true;
            else if (key[q] == sd[q]) atomicAdd(&table[slot[q]].count, 1u);
            else if (key[q] == kEmptyKey) claim[q] = true;
            else slow[q] = true;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (claim[q]) was[q] = atomicCAS(reinterpret_cast<unsigned long long*>(&table[slot[q]].key), (unsigned long long)kEmptyKey, (unsigned long long)sd[q]);
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (claim[q]) { if (was[q] == kEmptyKey || was[q] == sd[q]) atomicAdd(&table[slot[q]].count, 1u); else slow[q] = true; }
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (slow[q]) tableInsert(table, mask, sd[q], 1u, acc);
}
__global__ void __launch_bounds__(256) count_misses(const u64* __restrict__ missQ, u64 missCap, TableSlot* table, u64 mask, SampleAcc* acc,
                                                    cudaTextureObject_t tableTex) {
    const u64 n = min((u64)acc->missCount, missCap);
    const u64 T = (u64)gridDim.x * blockDim.x;
    for (u64 i0 = (u64)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * T) {
        u64 sd[4]; bool has[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const u64 i = i0 + (u64)q * T; has[q] = i < n; sd[q] = has[q] ? __ldcs(missQ + i) : 0; }
        insertFour(sd, has, table, mask, acc, tableTex);
    }
}

// ---- partitioned counting: read tables larger than L2 (bacterial-scale samples: 4e8 seed instances, 4e7 distinct, a 2 GB table) ------------
// Counting straight into such a table is one random DRAM sector per seed instance -- configs[3] spends 20.8 of its 27 ms there.  Instead
// the instances are first SCATTERED by table region (bucket = high bits of the home slot, regions of ~32 MB so that one fits L2 next to the
// streams passing through), then counted bucket after bucket by every warp at the same time: all probes and atomics of a phase land in
// the one table region that is L2-resident, and DRAM sees only streams (the lists, the scattered seeds written and read once, the
// table itself once).  No staging, no flush protocol: every (resident warp, bucket) pair owns a fixed region, positions come from the warp's
// own cursors in shared memory (no cross-warp contention), the 8-byte stores of a region are consecutive and merge in L2.  A full
// region (capacity = 1.25x the expected share + slack; hash-uniform shares vary by a few per cent) sends its seed the direct way.
template <int KT, int LT>
__global__ void __launch_bounds__(1024, 1) scatter_seeds_lane(const u64* __restrict__ synBuf, const unsigned* __restrict__ synCount,
                                                              const u64* __restrict__ packedOff, u64 nReads, WorkspaceView W) {
    static_assert(LT == 1 || LT == 3, "lane-per-read seed formation is specialised for l = 1 and l = 3");
    extern __shared__ unsigned sCurAll[];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, B = W.bktCount;
    unsigned* sCur = sCurAll + warp * B;
    const u64 gw = (u64)blockIdx.x * 32 + warp;
    u32* fill = W.bktFill + gw * B;
    for (unsigned b = lane; b < B; b += 32) sCur[b] = fill[b];   // cursors continue where the previous slice of the sample stopped
    __syncwarp();
    u64* __restrict__ region = W.bktBuf + gw * B * (u64)W.bktRegionCap;
    const u64 mask = W.tableMask;
    const u64 warpsTotal = (u64)gridDim.x * 32;
    for (u64 r0 = gw * 32; r0 < nReads; r0 += warpsTotal * 32) {
        const u64 r = r0 + lane;
        const int n = r < nReads ? (int)__ldg(&synCount[r]) : 0;
        const u64* __restrict__ h = synBuf + (r < nReads ? __ldg(&packedOff[r]) : 0) * 32;
        const int nS = LT <= 1 ? n : (n >= LT ? n - LT + 1 : 0);
        int maxS = nS;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) maxS = max(maxS, __shfl_xor_sync(0xffffffffu, maxS, d));
        u64 h0 = 0, h1 = 0;
        if (LT == 3 && nS > 0) { h0 = __ldg(h); h1 = __ldg(h + 1); }
        for (int j0 = 0; j0 < maxS; j0 += 4) {
            u64 x[4], sd[4]; bool has[4];
#pragma unroll
            for (int q = 0; q < 4; q += 2) {   // 16 bytes at a time, as in count_seeds_lane
                has[q] = j0 + q < nS; has[q + 1] = j0 + q + 1 < nS;
                uint4 t = make_uint4(0u, 0u, 0u, 0u);
                if (has[q]) t = __ldcs(reinterpret_cast<const uint4*>(h + j0 + q + (LT - 1)));
                x[q] = (u64)t.x | ((u64)t.y << 32); x[q + 1] = (u64)t.z | ((u64)t.w << 32);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (LT == 1) sd[q] = x[q];
                else {
                    const u64 fw = rol64(h0, (unsigned)((KT * 2) & 63)) ^ rol64(h1, (unsigned)(KT & 63)) ^ x[q];
                    const u64 rw = h0 ^ rol64(h1, (unsigned)(KT & 63)) ^ rol64(x[q], (unsigned)((KT * 2) & 63));
                    sd[q] = umin64(fw, rw);
                    has[q] = has[q] && fw != rw;
                    h0 = h1; h1 = x[q];
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (!has[q]) continue;
                const unsigned b = (unsigned)((mixKey(sd[q]) & mask) >> W.bktShift);
                const unsigned pos = atomicAdd(&sCur[b], 1u);
                if (pos < W.bktRegionCap) __stcs(region + (u64)b * W.bktRegionCap + pos, sd[q]);
                else tableInsert(W.table, mask, sd[q], 1u, W.acc);   // region full: counted right away
            }
        }
    }
    __syncwarp();
    for (unsigned b = lane; b < B; b += 32) fill[b] = min(sCur[b], W.bktRegionCap);
}
// second half: bucket after bucket, every warp its own region of the bucket, four seeds per lane in flight
__global__ void __launch_bounds__(1024, 1) count_buckets(WorkspaceView W) {
    const unsigned lane = threadIdx.x & 31u, B = W.bktCount;
    const u64 gw = (u64)blockIdx.x * 32 + (threadIdx.x >> 5);
    const u64* __restrict__ region = W.bktBuf + gw * B * (u64)W.bktRegionCap;
    for (unsigned b = 0; b < B; ++b) {
        const unsigned n = W.bktFill[gw * B + b];
        const u64* __restrict__ src = region + (u64)b * W.bktRegionCap;
        for (unsigned i0 = lane; i0 < n; i0 += 128) {
            u64 sd[4]; bool has[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const unsigned i = i0 + 32u * q; has[q] = i < n; sd[q] = has[q] ? __ldcs(src + i) : 0; }
            insertFour(sd, has, W.table, W.tableMask, W.acc, W.tableTex);
        }
    }
}
bool bucketCountingSupports(int k, int l) { return (l == 3 && (k == 19 || k == 15)) || l <= 1; }
void launchCountBuckets(WorkspaceView W, cudaStream_t st) {
    if (!W.bktCount) return;
    noteLaunch(), count_buckets<<<kBktBlocks, 1024, 0, st>>>(W);
}
template <int KT, int LT>
static void launchScatterT(const u64* synBuf, const unsigned* synCount, const u64* packedOff, u64 nReads, const WorkspaceView& W, cudaStream_t st) {
    const size_t sm = (size_t)32 * W.bktCount * sizeof(unsigned);
    noteLaunch(), scatter_seeds_lane<KT, LT><<<kBktBlocks, 1024, sm, st>>>(synBuf, synCount, packedOff, nReads, W);
}
static void launchScatter(const u64* synBuf, const unsigned* synCount, const u64* packedOff, u64 nReads, int k, int l, const WorkspaceView& W, cudaStream_t st) {
    if (k == 19 && l == 3) return launchScatterT<19, 3>(synBuf, synCount, packedOff, nReads, W, st);
    if (k == 15 && l == 3) return launchScatterT<15, 3>(synBuf, synCount, packedOff, nReads, W, st);
    return launchScatterT<0, 1>(synBuf, synCount, packedOff, nReads, W, st);
}
// Measured on the 1M x 150 bp sample (r02f): count_seeds_lane without the global walk 328 us + count_misses 136-158 us = 475 us against
// 393 us with the walk inside -- the walk was not the bottleneck (list loads + shared-memory probes are), and on its own the queue pass
// pays the 11.5 M L2 round trips that the fused kernel hides.  Kept as an experiment switch: PM_MISS_QUEUE=1
static bool useMissQueue() {
    static const bool v = [] { const char* e = std::getenv("PM_MISS_QUEUE"); return e ? std::atoi(e) != 0 : false; }();
    return v;
}
static int listPrefetch() {   // experiment switch PM_LIST_PREFETCH (0 / 1)
    static const int v = [] { const char* e = std::getenv("PM_LIST_PREFETCH"); return e ? std::atoi(e) : 0; }();
    return v;
}
template <int KT, int LT>
static void launchCountLane(const u64* synBuf, const unsigned* synCount, const u64* packedOff, u64 nReads, TableSlot* table, u64 mask, SampleAcc* acc,
                            cudaTextureObject_t tableTex, u64* missQ, u64 missCap, cudaStream_t st) {
    if (nReads >= aggMinReads()) {   // whole samples and the per-rank slices of sharded ones: per-SM pre-aggregation (a block must see enough reads for its table to pay off)
        const size_t sm = (size_t)kAggSlots * (sizeof(u64) + sizeof(u32));
        if (missQ && missCap && useMissQueue()) {
            const size_t smq = sm + (size_t)32 * kWarpQueue * sizeof(u64);
            cudaMemsetAsync(&acc->missCount, 0, sizeof(acc->missCount), st);
            cudaFuncSetAttribute(count_seeds_lane<KT, LT, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smq);
            noteLaunch(), count_seeds_lane<KT, LT, true, true><<<148, 1024, smq, st>>>(synBuf, synCount, packedOff, nReads, table, mask, acc, tableTex, missQ, missCap, listPrefetch());
            noteLaunch(), count_misses<<<148 * 8, 256, 0, st>>>(missQ, missCap, table, mask, acc, tableTex);
            return;
        }
        cudaFuncSetAttribute(count_seeds_lane<KT, LT, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        noteLaunch(), count_seeds_lane<KT, LT, true, false><<<148, 1024, sm, st>>>(synBuf, synCount, packedOff, nReads, table, mask, acc, tableTex, nullptr, 0, listPrefetch());
    } else {
        u64 g = (nReads + 255) / 256; if (g > 148ull * 8) g = 148ull * 8;
        noteLaunch(), count_seeds_lane<KT, LT, false, false><<<(unsigned)(g ? g : 1), 256, 0, st>>>(synBuf, synCount, packedOff, nReads, table, mask, acc, tableTex, nullptr, 0, listPrefetch());
    }
}
template <int MODE>
static void launchSeedsFromSyncmers(const u64* synBuf, const unsigned* synCount, const u64* packedOff, const u64* winOff, u64 nReads, int k, int l,
                                    TableSlot* table, u64 mask, SampleAcc* acc, u64* outHash, u64* outCount, cudaTextureObject_t tableTex, cudaStream_t st);
// launches below this many reads (slices of the host-buffer pipeline, per-rank slices of sharded samples) give the lane-per-read kernels at most
// one group of 32 reads per resident warp: a single wave of long dependent chains.  There the warp-per-read kernel (lanes over the seeds of
// one read, every chain one probe long) is faster.  Tuning override: PM_COUNT_WARP_BELOW
static u64 countWarpBelow() {
    static const u64 v = [] { const char* e = std::getenv("PM_COUNT_WARP_BELOW"); return e ? (u64)std::strtoull(e, nullptr, 10) : (u64)200000; }();   // measured: 125 k reads 90 vs 107 us, 250 k equal, 500 k+ slower
    return v;
}
static void launchCountSeeds(const u64* synBuf, const unsigned* synCount, const u64* packedOff, u64 nReads, int k, int l, TableSlot* table, u64 mask,
                             SampleAcc* acc, cudaTextureObject_t tableTex, u64* missQ, u64 missCap, cudaStream_t st) {
    if (nReads < countWarpBelow()) return launchSeedsFromSyncmers<0>(synBuf, synCount, packedOff, nullptr, nReads, k, l, table, mask, acc, nullptr, nullptr, tableTex, st);
    // panmap's parameter sets (l = 3 with k = 19 / 15, and l <= 1) use the lane-per-read kernel; any other (k, l) the flat one below
    if (k == 19 && l == 3) return launchCountLane<19, 3>(synBuf, synCount, packedOff, nReads, table, mask, acc, tableTex, missQ, missCap, st);
    if (k == 15 && l == 3) return launchCountLane<15, 3>(synBuf, synCount, packedOff, nReads, table, mask, acc, tableTex, missQ, missCap, st);
    if (l <= 1) return launchCountLane<0, 1>(synBuf, synCount, packedOff, nReads, table, mask, acc, tableTex, missQ, missCap, st);
    if (nReads >= aggMinReads()) {   // whole samples and the per-rank slices of sharded ones: per-SM pre-aggregation (a block must see enough reads for its table to pay off)
        const size_t sm = (size_t)kAggSlots * (sizeof(u64) + sizeof(u32));
        cudaFuncSetAttribute(count_seeds<0, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        noteLaunch(), count_seeds<0, 0, true><<<148, 1024, sm, st>>>(synBuf, synCount, packedOff, nReads, k, l, table, mask, acc, tableTex);
        return;
    }
    u64 g = (nReads + 255) / 256; if (g > 148ull * 8) g = 148ull * 8;
    noteLaunch(), count_seeds<0, 0, false><<<(unsigned)(g ? g : 1), 256, 0, st>>>(synBuf, synCount, packedOff, nReads, k, l, table, mask, acc, tableTex);
}

template <int MODE>
static void launchSeedsFromSyncmers(const u64* synBuf, const unsigned* synCount, const u64* packedOff, const u64* winOff, u64 nReads, int k, int l,
                                    TableSlot* table, u64 mask, SampleAcc* acc, u64* outHash, u64* outCount, cudaTextureObject_t tableTex, cudaStream_t st) {
    u64 g = (nReads + 7) / 8; if (g > 148ull * 8) g = 148ull * 8;
    const unsigned grid = (unsigned)(g ? g : 1);
    if (k == 19 && l == 3)
        noteLaunch(), seeds_from_syncmers<MODE, 19, 3><<<grid, 256, 0, st>>>(synBuf, synCount, packedOff, winOff, nReads, k, l, table, mask, acc, outHash, outCount, tableTex);
    else if (k == 15 && l == 3)
        noteLaunch(), seeds_from_syncmers<MODE, 15, 3><<<grid, 256, 0, st>>>(synBuf, synCount, packedOff, winOff, nReads, k, l, table, mask, acc, outHash, outCount, tableTex);
    else
        noteLaunch(), seeds_from_syncmers<MODE, 0, 0><<<grid, 256, 0, st>>>(synBuf, synCount, packedOff, winOff, nReads, k, l, table, mask, acc, outHash, outCount, tableTex);
}

static size_t genericSmemBytes(const SeederParams& P) {
    return sizeof(SeedTables) + (size_t)seederRingWords(P.k, P.s, 1) * kSeedThreads * sizeof(u64);
}
static unsigned seedGrid(u64 nReads) {
    u64 g = (nReads + kSeedThreads - 1) / kSeedThreads;
    if (g > 148ull * 16) g = 148ull * 16;
    return (unsigned)(g ? g : 1);
}
// launches with at least this many reads take syncmers_rank (its blocks each bring the 128 KB rank table in: not worth it for small slices).
// Tuning override: PM_RANK_MIN_READS (0 = always, huge = never)
static u64 rankMinReads() {
    static const u64 v = [] { const char* e = std::getenv("PM_RANK_MIN_READS"); return e ? (u64)std::strtoull(e, nullptr, 10) : (u64)20000; }();
    return v;
}
static int rankThreads() {   // tuning override: PM_RANK_THREADS (512 or 768)
    static const int v = [] { const char* e = std::getenv("PM_RANK_THREADS"); return e ? std::atoi(e) : 512; }();
    return v;
}
static bool rankStage() {   // experiment switch: PM_RANK_STAGE=1 stages the reads through shared memory (TMA bulk copy per group of 32 reads)
    static const bool v = [] { const char* e = std::getenv("PM_RANK_STAGE"); return e ? std::atoi(e) != 0 : false; }();
    return v;
}
template <int K, int THREADS>
static void launchRankT(const uint4* packed, const u64* off, const u64* packedOff, u64 nReads, const SeederParams& P, const SeedTables* dT,
                        u64* synBuf, unsigned* synCount, const unsigned char* dup, const u64* endOff, const char* reads, cudaStream_t st) {
    const size_t sm = (size_t)kRankEntries * 2 + 2 * kPairStride * sizeof(u64) + 256;
    u64 g = (nReads + THREADS - 1) / THREADS; if (g > 148) g = 148;
    const unsigned grid = (unsigned)(g ? g : 1);
    if (THREADS == 512 && reads && P.maxLen > 0 && P.maxLen <= kStageMaxLen && rankStage()) {   // reads staged through shared memory by TMA
        const size_t sms = sm + (size_t)(THREADS / 32) * kStagePitch;
        if (P.trimEnd == 0) {
            cudaFuncSetAttribute(syncmers_rank<K, true, true, THREADS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sms);
            noteLaunch(), syncmers_rank<K, true, true, THREADS, true><<<grid, THREADS, sms, st>>>(packed, off, packedOff, nReads, P, dT, synBuf, synCount, dup, endOff, reads);
        } else {
            cudaFuncSetAttribute(syncmers_rank<K, true, false, THREADS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sms);
            noteLaunch(), syncmers_rank<K, true, false, THREADS, true><<<grid, THREADS, sms, st>>>(packed, off, packedOff, nReads, P, dT, synBuf, synCount, dup, endOff, reads);
        }
        return;
    }
#define PM_RANK_LAUNCH(A, N)                                                                                                        \
    do {                                                                                                                            \
        cudaFuncSetAttribute(syncmers_rank<K, A, N, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);                \
        noteLaunch(), syncmers_rank<K, A, N, THREADS><<<grid, THREADS, sm, st>>>(packed, off, packedOff, nReads, P, dT, synBuf, synCount, dup, endOff, reads); \
    } while (0)
    if (reads && P.trimEnd == 0) PM_RANK_LAUNCH(true, true);
    else if (reads) PM_RANK_LAUNCH(true, false);
    else if (P.trimEnd == 0) PM_RANK_LAUNCH(false, true);
    else PM_RANK_LAUNCH(false, false);
#undef PM_RANK_LAUNCH
}
template <int K>
static void launchRank(const uint4* packed, const u64* off, const u64* packedOff, u64 nReads, const SeederParams& P, const SeedTables* dT,
                       u64* synBuf, unsigned* synCount, const unsigned char* dup, const u64* endOff, const char* reads, cudaStream_t st) {
    if (rankThreads() == 768) return launchRankT<K, 768>(packed, off, packedOff, nReads, P, dT, synBuf, synCount, dup, endOff, reads, st);
    return launchRankT<K, 512>(packed, off, packedOff, nReads, P, dT, synBuf, synCount, dup, endOff, reads, st);
}
template <int K, int S>
static void launchFast(const uint4* packed, const u64* off, const u64* packedOff, u64 nReads, const SeederParams& P, const SeedTables* dT,
                       u64* synBuf, unsigned* synCount, const unsigned char* dup, const u64* endOff, const char* reads, cudaStream_t st) {
    if (S == kRankS && nReads >= rankMinReads()) return launchRank<K>(packed, off, packedOff, nReads, P, dT, synBuf, synCount, dup, endOff, reads, st);
    const size_t sm = sizeof(SeedTables) + 256 + 4 * kPairStride * sizeof(u64);
    if (reads && P.trimEnd == 0)
        noteLaunch(), syncmers_fast<K, S, true, true><<<seedGrid(nReads), kSeedThreads, sm, st>>>(packed, off, packedOff, nReads, P, dT, synBuf, synCount, dup, endOff, reads);
    else if (reads)
        noteLaunch(), syncmers_fast<K, S, true, false><<<seedGrid(nReads), kSeedThreads, sm, st>>>(packed, off, packedOff, nReads, P, dT, synBuf, synCount, dup, endOff, reads);
    else if (P.trimEnd == 0)   // 4-bit codes (pm_place_packed and the list utilities)
        noteLaunch(), syncmers_fast<K, S, false, true><<<seedGrid(nReads), kSeedThreads, sm, st>>>(packed, off, packedOff, nReads, P, dT, synBuf, synCount, dup, endOff, nullptr);
    else
        noteLaunch(), syncmers_fast<K, S, false, false><<<seedGrid(nReads), kSeedThreads, sm, st>>>(packed, off, packedOff, nReads, P, dT, synBuf, synCount, dup, endOff, nullptr);
}
// true when launchSeedTable hashes these parameters straight from the ASCII reads (no pack_reads needed beforehand)
bool seedTableReadsAscii(const SeederParams& P) { return !P.open && P.t == 0 && P.s == 8 && (P.k == 19 || P.k == 15); }
static void launchSyncmers(const uint4* packed, const u64* off, const u64* packedOff, u64 nReads, const SeederParams& P, const SeedTables* dT,
                           u64* synBuf, unsigned* synCount, const unsigned char* dup, const u64* endOff, cudaStream_t st, const char* reads = nullptr) {
    if (!P.open && P.t == 0 && P.k == 19 && P.s == 8) return launchFast<19, 8>(packed, off, packedOff, nReads, P, dT, synBuf, synCount, dup, endOff, reads, st);
    if (!P.open && P.t == 0 && P.k == 15 && P.s == 8) return launchFast<15, 8>(packed, off, packedOff, nReads, P, dT, synBuf, synCount, dup, endOff, reads, st);
    const size_t sm = genericSmemBytes(P);
    cudaFuncSetAttribute(syncmers_generic<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    noteLaunch(), syncmers_generic<0><<<seedGrid(nReads), kSeedThreads, sm, st>>>(packed, off, packedOff, nullptr, nReads, P, dT, synBuf, synCount, nullptr,
                                                                     nullptr, nullptr, nullptr, dup, endOff, nullptr, 0);
}
// reads -> count table: syncmer lists per read, then their seeds into the table.  (Running the two as one kernel, or concurrently
// on two streams, was measured and is slower: both are limited by the same L1/LSU data pipe and the hashing needs every warp an
// SM can hold.)
void launchSeedTable(const uint4* packed, const u64* off, const u64* packedOff, u64 nReads, const SeederParams& P,
                     const SeedTables* dTables, WorkspaceView W, cudaStream_t st, cudaEvent_t between, const unsigned char* dup,
                     const u64* endOff, const char* reads, cudaEvent_t tableReady) {
    if (nReads == 0) { if (tableReady) cudaStreamWaitEvent(st, tableReady, 0); return; }
    launchSyncmers(packed, off, packedOff, nReads, P, dTables, W.synBuf, W.synCount, dup, endOff, st, reads);
    if (between) cudaEventRecord(between, st);
    if (tableReady) cudaStreamWaitEvent(st, tableReady, 0);   // the table was cleared on a side stream while the syncmer kernel ran
    if (W.bktCount) launchScatter(W.synBuf, W.synCount, packedOff, nReads, P.k, P.l, W, st);   // partitioned counting: launchCountBuckets follows the last slice
    else launchCountSeeds(W.synBuf, W.synCount, packedOff, nReads, P.k, P.l, W.table, W.tableMask, W.acc, W.tableTex, W.missQ, W.missCap, st);
}
// --min-seed-quality > 0 (off by default): the generic kernels with per-syncmer pass flags; quals has one byte per base at the reads'
// offsets (compressed in lockstep for hpc indexes), synPass one byte per synBuf entry
void launchSeedTableQuality(const uint4* packed, const u64* off, const u64* packedOff, u64 nReads, const SeederParams& P,
                            const SeedTables* dTables, WorkspaceView W, cudaStream_t st, const u64* endOff, const char* quals,
                            int minSeedQuality, unsigned char* synPass) {
    if (nReads == 0) return;
    const size_t sm = genericSmemBytes(P);
    cudaFuncSetAttribute(syncmers_generic<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    noteLaunch(), syncmers_generic<3><<<seedGrid(nReads), kSeedThreads, sm, st>>>(packed, off, packedOff, nullptr, nReads, P, dTables, W.synBuf, W.synCount,
                                                                     nullptr, synPass, nullptr, nullptr, nullptr, endOff, quals, minSeedQuality * P.k);
    u64 g = (nReads + 7) / 8; if (g > 148ull * 8) g = 148ull * 8;
    noteLaunch(), seeds_from_syncmers<0, 0, 0><<<(unsigned)(g ? g : 1), 256, 0, st>>>(W.synBuf, W.synCount, packedOff, nullptr, nReads, P.k, P.l, W.table,
                                                                         W.tableMask, W.acc, nullptr, nullptr, W.tableTex, synPass);
}
// mode 1: syncmer (hash, isReverse, pos) lists == seeding::rollingSyncmers(returnAll=false); mode 2: per-read seed lists
void launchSeedList(const uint4* packed, const u64* off, const u64* packedOff, const u64* winOff, u64 nReads,
                    const SeederParams& P, const SeedTables* dTables, int mode, u64* synBuf, unsigned* synCount, u64* outHash,
                    unsigned char* outRev, long long* outPos, u64* outCount, cudaStream_t st) {
    if (nReads == 0) return;
    if (mode == 1) {
        const size_t sm = genericSmemBytes(P);
        cudaFuncSetAttribute(syncmers_generic<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        noteLaunch(), syncmers_generic<1><<<seedGrid(nReads), kSeedThreads, sm, st>>>(packed, off, packedOff, winOff, nReads, P, dTables, nullptr, nullptr,
                                                                         outHash, outRev, outPos, outCount, nullptr, nullptr, nullptr, 0);
    } else {
        launchSyncmers(packed, off, packedOff, nReads, P, dTables, synBuf, synCount, nullptr, nullptr, st);
        launchSeedsFromSyncmers<2>(synBuf, synCount, packedOff, winOff, nReads, P.k, P.l, nullptr, 0, nullptr, outHash, outCount, 0, st);
    }
}

// index builder (pm_build_kernels.cu): per-sequence seed lists of sequences that sit at a fixed pitch (off[r] = r * pitch) and end at
// endOff[r]; sequence r's seeds land at outHash[winOff[r] ...], outCount[r] of them
void launchSeedListsEnd(const uint4* packed, const u64* off, const u64* endOff, const u64* packedOff, const u64* winOff, u64 nSeqs, const SeederParams& P,
                        const SeedTables* dTables, u64* synBuf, unsigned* synCount, u64* outHash, u64* outCount, cudaStream_t st) {
    if (nSeqs == 0) return;
    launchSyncmers(packed, off, packedOff, nSeqs, P, dTables, synBuf, synCount, nullptr, endOff, st);
    launchSeedsFromSyncmers<2>(synBuf, synCount, packedOff, winOff, nSeqs, P.k, P.l, nullptr, 0, nullptr, outHash, outCount, 0, st);
}

// ------------------------------------------------------------------------------------------------------
// table maintenance
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) table_clear(TableSlot* table, u64 cap) {
    const uint4 e = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
    uint4* t = reinterpret_cast<uint4*>(table);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (u64)gridDim.x * blockDim.x) t[i] = e;
}
void launchTableClear(WorkspaceView W, cudaStream_t st) { noteLaunch(), table_clear<<<streamGrid(W.tableCap, 4), 256, 0, st>>>(W.table, W.tableCap); }
// the same pass also zeroes the sample's accumulators (start of a sample whose first kernel, the syncmer kernel, touches neither)
__global__ void __launch_bounds__(256) sample_begin(TableSlot* table, u64 cap, SampleAcc* acc) {
    const uint4 e = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
    uint4* t = reinterpret_cast<uint4*>(table);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (u64)gridDim.x * blockDim.x) t[i] = e;
    if (blockIdx.x == 0) {
        static_assert(sizeof(SampleAcc) % 4 == 0, "SampleAcc is cleared word by word");
        u32* a = reinterpret_cast<u32*>(acc);
        for (unsigned i = threadIdx.x; i < sizeof(SampleAcc) / 4; i += blockDim.x) a[i] = 0u;
    }
}
void launchSampleBegin(WorkspaceView W, cudaStream_t st) { noteLaunch(), sample_begin<<<streamGrid(W.tableCap, 4), 256, 0, st>>>(W.table, W.tableCap, W.acc); }

__global__ void __launch_bounds__(256) table_import(TableSlot* table, u64 mask, SampleAcc* acc, const u64* __restrict__ hash,
                                                    const long long* __restrict__ count, u64 n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        if (count[i] > 0) tableInsert(table, mask, hash[i], (u32)count[i], acc);
}
void launchTableImport(WorkspaceView W, const u64* hash, const long long* count, u64 n, cudaStream_t st) {
    if (!n) return;
    noteLaunch(), table_import<<<streamGrid(n, 1), 256, 0, st>>>(W.table, W.tableMask, W.acc, hash, count, n);
}

__global__ void __launch_bounds__(256) table_export(const TableSlot* __restrict__ table, u64 cap, const SampleAcc* acc, u64* outHash,
                                                    long long* outCount, unsigned* counter, u64 outCap) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (u64)gridDim.x * blockDim.x) {
        const uint4 v = ldSlot(table, i);
        const u64 k = slotKey(v); const u32 c = v.z;
        if (k != kEmptyKey && c > 0) {
            const unsigned o = atomicAdd(counter, 1u);
            if (o < outCap) { outHash[o] = k; outCount[o] = (long long)c; }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && acc->emptyKeyCount > 0) {
        const unsigned o = atomicAdd(counter, 1u);
        if (o < outCap) { outHash[o] = kEmptyKey; outCount[o] = acc->emptyKeyCount; }
    }
}
void launchTableExport(WorkspaceView W, u64* hash, long long* count, unsigned* counter, u64 cap, cudaStream_t st) {
    noteLaunch(), table_export<<<streamGrid(W.tableCap, 2), 256, 0, st>>>(W.table, W.tableCap, W.acc, hash, count, counter, cap);
}

// Same-address global atomics are serviced one at a time by the L2 (a few ns each), so the passes below never let more than one
// thread per block touch a shared counter: blocks compact / reduce in shared memory and publish per-block partials.
//
// pass 1 (the only pass over the whole table): erase the four homopolymer k-mer hashes (placement.cpp:1708-1718), gather the
// statistics of the auto min-read-support rule (placement.cpp:931-955) and the pre-filter totals, and compact the occupied
// slots into a dense (key, count) list so that everything that follows works on U entries instead of the table's capacity.
// A block handles 2048 slots at a time: eight independent 16-byte loads per thread, compaction in shared memory, one global
// atomic for the block's share of the list, coalesced copy-out.
constexpr int kScanSlots = 2048;
__global__ void __launch_bounds__(256) table_scan(WorkspaceView W, const u64* __restrict__ homo) {
    __shared__ u64 sKey[kScanSlots];
    __shared__ u32 sCnt[kScanSlots];
    __shared__ unsigned sN, sBase;
    __shared__ long long sRed[8][4];
    const u64 h0 = homo[0], h1 = homo[1], h2 = homo[2], h3 = homo[3];
    TableSlot* table = W.table; const u64 cap = W.tableCap; SampleAcc* acc = W.acc;
    long long ms = 0, mc = 0, en = 0, total = 0;
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    if (tid == 0) sN = 0;
    __syncthreads();
    const u64 nBlocks = (cap + kScanSlots - 1) / kScanSlots;
    for (u64 blk = blockIdx.x; blk < nBlocks; blk += gridDim.x) {
        const u64 base = blk * kScanSlots;
        uint4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { const u64 i = base + q * 256 + tid; v[q] = i < cap ? ldSlot(table, i) : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0); }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const u64 k = slotKey(v[q]);
            bool occ = k != kEmptyKey && v[q].z > 0;
            if (occ && (k == h0 || k == h1 || k == h2 || k == h3)) { table[base + q * 256 + tid].count = 0; occ = false; }
            if (occ) { const long long c = v[q].z; ++en; total += c; if (c >= 2) { ms += c; ++mc; } }
            const unsigned m = __ballot_sync(0xffffffffu, occ);
            if (m) {
                unsigned o = 0;
                if (lane == 0) o = atomicAdd(&sN, (unsigned)__popc(m));
                o = __shfl_sync(0xffffffffu, o, 0) + __popc(m & ((1u << lane) - 1u));
                if (occ) { sKey[o] = k; sCnt[o] = v[q].z; }
            }
        }
        __syncthreads();
        const unsigned n = sN;
        if (tid == 0 && n) sBase = atomicAdd(&acc->entCount, n);
        __syncthreads();
        const unsigned ob = sBase;
        if (tid == 0) sN = 0;   // every thread has read n; the barrier below orders the reset before the next block's atomics
        for (unsigned i = tid; i < n; i += 256) { W.entKey[ob + i] = sKey[i]; W.entCnt[ob + i] = sCnt[i]; }
        __syncthreads();
    }
    if (blockIdx.x == 0 && tid == 0 && acc->emptyKeyCount > 0) {   // the one key that cannot live in the table
        const long long c = acc->emptyKeyCount; ++en; total += c; if (c >= 2) { ms += c; ++mc; }
    }
    ms = warpSumLL(ms); mc = warpSumLL(mc); en = warpSumLL(en); total = warpSumLL(total);
    if (lane == 0) { sRed[tid >> 5][0] = ms; sRed[tid >> 5][1] = mc; sRed[tid >> 5][2] = en; sRed[tid >> 5][3] = total; }
    __syncthreads();
    if (tid < 4) {
        long long t = 0;
        for (int q = 0; q < 8; ++q) t += sRed[q][tid];
        reinterpret_cast<long long*>(&W.scanPart[blockIdx.x])[tid] = t;
    }
}

// ------------------------------------------------------------------------------------------------------
// --seed-mask-fraction (placement.cpp:1748-1799; off by default): drop the floor(frac * U) most frequent seeds before the
// min-support rule and the magnitudes see the table.  The reference sorts by count only, so which seeds go at a tied cut is
// unspecified there; like the oracle this path takes the smaller hashes first.  The cut (count c*, hash h*) is found by
// bisection with one counting pass over the compacted entries per step, driven from the host -- a handful of tiny launches
// on an optional path, nothing on the default one.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mask_count(WorkspaceView W, u32 cThr, u64 hThr, unsigned long long* __restrict__ out) {
    unsigned long long above = 0, tied = 0;   // entries with count > cThr; with count == cThr and key <= hThr
    const unsigned n = W.acc->entCount;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u32 c = W.entCnt[i];
        if (c > cThr) ++above;
        else if (c == cThr && c > 0 && W.entKey[i] <= hThr) ++tied;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && W.acc->emptyKeyCount > 0) {   // the key that lives outside the table (largest hash)
        const unsigned long long c = (unsigned long long)W.acc->emptyKeyCount;
        if (c > cThr) ++above; else if (c == cThr && kEmptyKey <= hThr) ++tied;
    }
    above = (unsigned long long)warpSumLL((long long)above); tied = (unsigned long long)warpSumLL((long long)tied);
    if ((threadIdx.x & 31u) == 0) { if (above) atomicAdd(out, above); if (tied) atomicAdd(out + 1, tied); }
}
__global__ void __launch_bounds__(256) mask_apply(WorkspaceView W, u32 cThr, u64 hThr, unsigned partIdx) {
    long long ms = 0, mc = 0, en = 0, total = 0;   // what the masked seeds had contributed to pass 1's statistics
    const unsigned n = W.acc->entCount;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u32 c = W.entCnt[i];
        if (c == 0) continue;
        const u64 k = W.entKey[i];
        if (!(c > cThr || (c == cThr && k <= hThr))) continue;
        W.entCnt[i] = 0;   // pass 2 skips it
        u64 slot = mixKey(k) & W.tableMask;
        for (int probe = 0; probe < 8192 && W.table[slot].key != k; ++probe) slot = (slot + 1) & W.tableMask;   // present by construction
        if (W.table[slot].key == k) W.table[slot].count = 0;   // like the homopolymer seeds: gone for the exported table too
        --en; total -= c; if (c >= 2) { ms -= c; --mc; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && W.acc->emptyKeyCount > 0) {
        const long long c = W.acc->emptyKeyCount;
        if ((unsigned long long)c > cThr || ((unsigned long long)c == cThr && kEmptyKey <= hThr)) {
            W.acc->emptyKeyCount = 0; --en; total -= c; if (c >= 2) { ms -= c; --mc; }
        }
    }
    ms = warpSumLL(ms); mc = warpSumLL(mc); en = warpSumLL(en); total = warpSumLL(total);
    if ((threadIdx.x & 31u) == 0 && en) {
        unsigned long long* p = reinterpret_cast<unsigned long long*>(&W.scanPart[partIdx]);   // zeroed by the host; two's complement adds
        atomicAdd(p + 0, (unsigned long long)ms); atomicAdd(p + 1, (unsigned long long)mc);
        atomicAdd(p + 2, (unsigned long long)en); atomicAdd(p + 3, (unsigned long long)total);
    }
}
// synchronous; returns the number of masked seeds.  scratch: two device words
static u64 maskTopSeeds(WorkspaceView W, double frac, unsigned nScanParts, unsigned long long* scratch, cudaStream_t st) {
    std::vector<ScanPartial> parts(nScanParts);
    cudaMemcpyAsync(parts.data(), W.scanPart, nScanParts * sizeof(ScanPartial), cudaMemcpyDeviceToHost, st);
    cudaMemsetAsync(&W.scanPart[nScanParts], 0, sizeof(ScanPartial), st);
    cudaStreamSynchronize(st);
    u64 U = 0;
    for (const ScanPartial& sp : parts) U += (u64)sp.entries;
    const u64 numToMask = (u64)(size_t)(frac * (double)U);   // placement.cpp:1775
    const u64 drop = std::min(numToMask, U);
    if (drop == 0) return 0;
    const unsigned grid = streamGrid(U, 4);
    auto query = [&](u32 c, u64 h, u64& above, u64& tied) {
        unsigned long long r[2];
        cudaMemsetAsync(scratch, 0, sizeof(r), st);
        noteLaunch(), mask_count<<<grid, 256, 0, st>>>(W, c, h, scratch);
        cudaMemcpyAsync(r, scratch, sizeof(r), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        above = r[0]; tied = r[1];
    };
    // c* = count of the drop-th most frequent seed: the smallest c with #(count > c) < drop
    u64 lo = 1, hi = 0xFFFFFFFFull, above = 0, tied = 0;
    while (lo < hi) {
        const u64 mid = lo + (hi - lo) / 2;
        query((u32)mid, 0, above, tied);
        if (above < drop) hi = mid; else lo = mid + 1;
    }
    const u32 cStar = (u32)lo;
    query(cStar, ~0ull, above, tied);
    const u64 need = drop - above;   // of the seeds with count c*, the `need` smallest hashes go
    u64 hStar = ~0ull;
    if (tied > need) {
        u64 hl = 0, hh = ~0ull;
        while (hl < hh) {   // smallest h with #(count == c*, key <= h) >= need
            const u64 mid = hl + (hh - hl) / 2;
            u64 a2, t2; query(cStar, mid, a2, t2);
            if (t2 >= need) hh = mid; else hl = mid + 1;
        }
        hStar = hl;
    }
    noteLaunch(), mask_apply<<<grid, 256, 0, st>>>(W, cStar, hStar, nScanParts);
    return drop;
}

// pass 2: computeReadSeedMagnitudes (placement.cpp:957-984) + scatter of log1p(count) to the seed-id array, one thread per
// compacted entry: log1p table, exact sums, count histogram, dictionary probe, scatter -- all independent, so the random
// accesses of many entries are in flight at once.  The seed id found for an entry (or kNone) is written back next to it so
// that reset_sample can clear exactly the touched ell entries without a separate list.
__global__ void __launch_bounds__(256) entries_finalize(DevIndexView I, WorkspaceView W, int configuredMinSupport, unsigned nScanParts) {
    __shared__ unsigned sHist[kHistSmem];
    __shared__ FinalizeShared sFin;
    __shared__ long long sStat[4];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (int i = tid; i < kHistSmem; i += blockDim.x) sHist[i] = 0;
    if (warp < 4) {   // totals of pass 1: warp q sums field q of the per-block partials
        long long t = 0;
        for (unsigned b = lane; b < nScanParts; b += 32) t += reinterpret_cast<const long long*>(&W.scanPart[b])[warp];
        t = warpSumLL(t);
        if (lane == 0) sStat[warp] = t;
    }
    __syncthreads();
    SampleAcc* acc = W.acc;
    if (blockIdx.x == 0 && tid == 0) { acc->multiSum = sStat[0]; acc->multiCount = sStat[1]; acc->entries = sStat[2]; acc->unique = sStat[2]; acc->total = sStat[3]; }
    const u32 minSup = (u32)resolveMinSupport(sStat[0], sStat[1], configuredMinSupport);
    FinalizeAcc A; A.mag = fxZero(); A.lsum = fxZero(); A.kept = 0; A.maxc = 0;
    const unsigned n = acc->entCount;
    for (unsigned i = blockIdx.x * blockDim.x + tid; i < n; i += gridDim.x * blockDim.x) {
        const u32 c = __ldcs(&W.entCnt[i]);
        u32 id = kNone;
        if (c >= minSup && c != 0) {   // c == 0: masked
            const double l = finalizeSums(I, W, c, A, sHist);
            id = dictLookup(I, __ldcs(&W.entKey[i]));
            if (id != kNone) W.ell[id] = __double2ll_rn(l * kEllScale);   // l >= ln 2: an exact multiple of 2^-53
        }
        W.entId[i] = id;
    }
    if (blockIdx.x == 0 && tid == 0 && acc->emptyKeyCount > 0) {
        const u32 c = (u32)acc->emptyKeyCount;
        if (c >= minSup) finalizeSums(I, W, c, A, sHist);
    }
    finalizeBlockEpilogue(W, A, sHist, &sFin);
}

// The reference adds the U' values log1p(count) (and their squares) one by one into an f64 (placement.cpp:967-977).
// That sequential sum drifts from the exact sum by O(U' * 2^-53) -- 1.7e-12 relative on the sars_20000 sample, more than the
// 1e-12 parity tolerance, and nearly independent of the (unspecified, hash-map) order because the addends repeat.
// driftOf() (pm_logic.cuh) returns the expectation of that drift over random orders, in closed form from the histogram of
// read counts: while the running sum is in binade [2^e, 2^(e+1)) every addition of x is rounded to a multiple of
// u = 2^(e-52), i.e. contributes rint(x/u)*u - x, and a fraction (hi-lo)/T of the additions happens in that binade.
// Adding it to the exact fixed-point sum reproduces the reference's value to ~1e-14 relative.
__device__ __forceinline__ double blockSumF64(double e, double* sRed) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) e += shflXorF64(e, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sRed[threadIdx.x >> 5] = e;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += sRed[w];
    return tot;
}

// root_and_scalars: (i) the weighted-containment denominator over the ROOT's deltas (placement.cpp:1863-1876), all blocks;
// (ii) in the block that finishes last: totals of the finalize pass from its per-block partials, the drift terms, the sample's scalars
// (computeReadSeedMagnitudes, placement.cpp:957-984).  One launch instead of two, and nothing of (ii) goes through global memory twice.
__global__ void __launch_bounds__(256) root_and_scalars(DevIndexView I, WorkspaceView W, int configuredMinSupport, unsigned nFinParts) {
    __shared__ double sRed[8];
    __shared__ u64 sFx[8][4];
    __shared__ long long sK[8], sM[8];
    __shared__ u64 sTot[4];
    __shared__ long long sKept, sMaxc;
    __shared__ unsigned sLast;
    SampleAcc* a = W.acc;
    const unsigned tid = threadIdx.x;
    if (I.hasRoot && I.rootDCount) {
        fx128 s = fxZero();
        for (u64 i = (u64)blockIdx.x * blockDim.x + tid; i < I.rootDCount; i += (u64)gridDim.x * blockDim.x) {
            const int c = (int)__ldg(&I.rootChild[i]);
            if (c > 0 && W.ell[__ldg(&I.rootId[i])] != 0) s = fxAdd(s, fxFromDouble(1.0 / (double)c));
        }
        s = fxWarpSum(s);
        if ((tid & 31) == 0 && (s.lo | (u64)s.hi)) fxAtomicAdd(a->wcDen, s);
    }
    __syncthreads();
    if (tid == 0) { __threadfence(); sLast = (atomicAdd(&a->finDone, 1u) == gridDim.x - 1u) ? 1u : 0u; }
    __syncthreads();
    if (!sLast) return;
    __threadfence();
    {   // totals of the finalize pass from the per-block partials (exact, order independent)
        fx128 pm = fxZero(), pl = fxZero(); long long pk = 0, px = 0;
        for (unsigned b = tid; b < nFinParts; b += blockDim.x) {
            const FinPartial P = W.finPart[b];
            fx128 t; t.lo = P.mag[0]; t.hi = (i64)P.mag[1]; pm = fxAdd(pm, t);
            t.lo = P.lsum[0]; t.hi = (i64)P.lsum[1]; pl = fxAdd(pl, t);
            pk += P.kept; px = max(px, P.maxc);
        }
        pm = fxWarpSum(pm); pl = fxWarpSum(pl); pk = warpSumLL(pk);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) px = max(px, (long long)shflXorU64((u64)px, d));
        if ((tid & 31) == 0) { const int w = tid >> 5; sFx[w][0] = pm.lo; sFx[w][1] = (u64)pm.hi; sFx[w][2] = pl.lo; sFx[w][3] = (u64)pl.hi; sK[w] = pk; sM[w] = px; }
        __syncthreads();
        if (tid == 0) {
            fx128 m = fxZero(), l = fxZero(); long long k = 0, x = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
                fx128 t; t.lo = sFx[w][0]; t.hi = (i64)sFx[w][1]; m = fxAdd(m, t);
                t.lo = sFx[w][2]; t.hi = (i64)sFx[w][3]; l = fxAdd(l, t);
                k += sK[w]; x = max(x, sM[w]);
            }
            // sharded samples: seeds with read count 1 that the index does not hold arrive as a count (SampleAcc::n1NotIndex): each of them
            // adds exactly what finalizeSums adds for c == 1 (their histogram share was entered by gathered_finalize)
            const long long n1 = a->n1NotIndex;
            if (n1 > 0 && resolveMinSupport(a->multiSum, a->multiCount, configuredMinSupport) <= 1) {
                const double l1 = I.log1pLut[1];
                m = fxAdd(m, fxMulU64(fxFromDouble(l1 * l1), (u64)n1)); l = fxAdd(l, fxMulU64(fxFromDouble(l1), (u64)n1));
                k += n1; x = max(x, 1LL);
            }
            a->magSq[0] = m.lo; a->magSq[1] = (u64)m.hi; a->logSum[0] = l.lo; a->logSum[1] = (u64)l.hi; a->kept = k; a->maxKeptCount = x;
            sTot[0] = m.lo; sTot[1] = (u64)m.hi; sTot[2] = l.lo; sTot[3] = (u64)l.hi; sKept = k; sMaxc = x;
        }
        __syncthreads();
    }
    fx128 m; m.lo = sTot[0]; m.hi = (i64)sTot[1];
    fx128 l; l.lo = sTot[2]; l.hi = (i64)sTot[3];
    const double magSqExact = fxToDouble(m), logSumExact = fxToDouble(l);
    double eMag = 0.0, eLog = 0.0;
    if (logSumExact > 0.0 && magSqExact > 0.0) {
        Binades BM, BL;
        makeBinades(magSqExact, BM); makeBinades(logSumExact, BL);
        const int cMax = (int)min((long long)kLog1pLut - 1, sMaxc);
        for (int c0 = tid; c0 <= cMax; c0 += 4 * (int)blockDim.x) {   // four bins per turn: their loads are in flight together
            unsigned mult[4]; double x[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int c = c0 + q * (int)blockDim.x; mult[q] = c <= cMax ? __ldcg(&W.countHist[c]) : 0u; x[q] = c <= cMax ? __ldg(&I.log1pLut[c]) : 0.0; }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (mult[q]) { eMag += driftOf(BM, x[q] * x[q]) * (double)mult[q]; eLog += driftOf(BL, x[q]) * (double)mult[q]; }
        }
    }
    const double dMag = blockSumF64(eMag, sRed);
    const double dLog = blockSumF64(eLog, sRed);
    // the histogram is consumed: its consumer leaves it zero for the next pass (every thread of this block is past its reads: the block sums above
    // are barriers), so no memset node sits in front of the next table_scan
    for (int i = (int)tid * 4; i < kLog1pLut; i += 4 * (int)blockDim.x) *reinterpret_cast<uint4*>(W.countHist + i) = make_uint4(0u, 0u, 0u, 0u);
    if (tid != 0) return;
    fx128 w; w.lo = __ldcg(&a->wcDen[0]); w.hi = (i64)__ldcg(&a->wcDen[1]);
    SampleScalars S;
    S.readMagnitude = sqrt(magSqExact + dMag);
    S.logContDenom = logSumExact + dLog;
    S.wcDenom = fxToDouble(w);
    S.uniqueKept = (double)sKept;
    S.minSupport = resolveMinSupport(a->multiSum, a->multiCount, configuredMinSupport);
    S.uniqueSeeds = a->unique;
    S.uniqueKeptInt = sKept;
    S.totalFrequency = a->total;
    S.multiSum = a->multiSum; S.multiCount = a->multiCount;
    S.tableEntries = a->entries;
    S.overflow = a->overflow;
    *W.scalars = S;
    a->finDone = 0;   // ready for the next pass over the same accumulators (staged scoring may run again without a new sample)
}

void launchTableScan(WorkspaceView W, const u64* homo, int nSM, unsigned* nPartsOut, cudaStream_t st) {
    const u64 nBlocks = (W.tableCap + kScanSlots - 1) / kScanSlots;
    const unsigned g1 = (unsigned)std::min<u64>(std::min<u64>(nBlocks ? nBlocks : 1, (u64)nSM * 4), kMaxPartials - 1);
    noteLaunch(), table_scan<<<g1, 256, 0, st>>>(W, homo);
    *nPartsOut = g1;
}
void launchRootAndScalars(DevIndexView I, WorkspaceView W, PlaceOpts O, unsigned nFinParts, cudaStream_t st) {
    u64 gr = 1;
    if (I.hasRoot && I.rootDCount) { gr = ((u64)I.rootDCount + 255) / 256; if (gr > 148 * 8) gr = 148 * 8; }
    noteLaunch(), root_and_scalars<<<(unsigned)gr, 256, 0, st>>>(I, W, O.minReadSupport, nFinParts);
}
void launchFinalize(DevIndexView I, WorkspaceView W, PlaceOpts O, const u64* homo, u64 expectedEntries, int nSM, cudaStream_t st,
                    double seedMaskFraction, unsigned long long* maskScratch, cudaStream_t stSide, cudaEvent_t evFork, cudaEvent_t evJoin) {
    unsigned g1 = 0;
    launchTableScan(W, homo, nSM, &g1, st);
    if (seedMaskFraction > 0.0) { maskTopSeeds(W, seedMaskFraction, g1, maskScratch, st); ++g1; }   // + one partial of corrections
    // sized from the previous sample's entry count (the kernel grid-strides, so any grid is correct)
    u64 g2 = (expectedEntries + 255) / 256; if (g2 < 1) g2 = 1; if (g2 > (u64)nSM * 4) g2 = (u64)nSM * 4; if (g2 > kMaxPartials) g2 = kMaxPartials;
    noteLaunch(), entries_finalize<<<(unsigned)g2, 256, 0, st>>>(I, W, O.minReadSupport, g1);
    if (stSide && evFork && evJoin) {
        // the sample's scalars are needed by prefix_scores, not by node_deltas: the root / scalars kernel (a latency chain that ends in one block)
        // runs beside node_deltas on a side stream; the caller makes `st` wait for evJoin before prefix_scores
        cudaEventRecord(evFork, st);
        cudaStreamWaitEvent(stSide, evFork, 0);
        launchRootAndScalars(I, W, O, (unsigned)g2, stSide);
        cudaEventRecord(evJoin, stSide);
    } else launchRootAndScalars(I, W, O, (unsigned)g2, st);
}

// after a sample: clear exactly the ell entries it set (their ids sit next to the compacted entries) and the segment records
// that are combined with atomics
__global__ void __launch_bounds__(256) reset_sample(DevIndexView I, WorkspaceView W) {
    const unsigned n = W.acc->entCount;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u32 id = __ldcs(&W.entId[i]);
        if (id != kNone) W.ell[id] = 0;
    }
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < I.nBoundary; i += gridDim.x * blockDim.x)
        *reinterpret_cast<uint4*>(W.segRec + I.boundarySegs[i]) = make_uint4(0u, 0u, 0u, 0u);
}
void launchResetSample(DevIndexView I, WorkspaceView W, cudaStream_t st) { noteLaunch(), reset_sample<<<148 * 4, 256, 0, st>>>(I, W); }

// ------------------------------------------------------------------------------------------------------
// K1 node_deltas: one pass over the packed delta words (4 B per delta), no block barriers.
// A warp owns a chunk of 512 consecutive words; lane l owns words [16 l, 16 l + 16): four 16-byte loads, then 16
// independent gathers of ell[seed id] (word = 2 * seed id + lost; the entry is log1p(read count) as an exact integer, 0 when the
// seed is not in the reads).  The kernel is bound by L1 wavefronts (one per distinct line a gather touches), so (i) the chunk is
// stored lane-interleaved and read with fully coalesced 16-byte loads, and (ii) the low seed ids -- the root's and other
// ancestral seeds, which almost half of all deltas refer to -- are served from a shared-memory copy of ell[0 .. kHotIds).
// Node boundaries come as one 16-bit mask per lane (bit j = word j is the last delta of its node): no offset array, no search.
//   * a node that begins and ends inside a lane is stored directly (sums of <= 16 terms fit 64 bits);
//   * a node spread over several lanes is combined with a segmented warp scan (96-bit) and stored by the lane where it ends;
//   * a node spread over several chunks (listed at flatten time, zeroed by reset_sample) is combined with integer atomics.
// All sums are integers, hence exact and independent of how deltas are split over lanes, warps, shards or GPUs.
// Output: segRec[segment] = {sum of +-ell, #gained - #lost}, read by prefix_scores through nodeSeg[].
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void segAtomicAdd(SegRec* r, u64 lo, int hi, int cnt) {
    if (lo | (u64)(u32)hi) {
        const u64 old = atomicAdd(reinterpret_cast<unsigned long long*>(&r->lo), (unsigned long long)lo);
        const int h = hi + ((old + lo < old) ? 1 : 0);
        if (h) atomicAdd(&r->hi, h);
    }
    if (cnt) atomicAdd(&r->cnt, cnt);
}
__device__ __forceinline__ void segStore(SegRec* r, u64 lo, int hi, int cnt) {
    *reinterpret_cast<uint4*>(r) = make_uint4((u32)lo, (u32)(lo >> 32), (u32)hi, (u32)cnt);
}

constexpr int kK1Threads = 768;  // one persistent block per SM: the hot table is filled once per SM
constexpr int kHotIds = 16384;   // seed ids below this (the root's seeds and other early, ancestral ones) are gathered from shared memory
__global__ void __launch_bounds__(kK1Threads, 1) node_deltas(DevIndexView I, WorkspaceView W, u32 chunksPerWarp) {
    extern __shared__ __align__(16) long long sHot[];   // [kHotIds] copy of ell[0 .. kHotIds)
    const long long* __restrict__ ell = W.ell;
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    const unsigned ltMask = (1u << lane) - 1u;
    {
        const u32 nHot = (u32)min((u64)kHotIds, I.nSeeds + 1);
        const longlong2* src = reinterpret_cast<const longlong2*>(ell);
        longlong2* dst = reinterpret_cast<longlong2*>(sHot);
        for (u32 i = tid; i < nHot / 2; i += kK1Threads) dst[i] = __ldg(src + i);
        if (tid == 0 && (nHot & 1u)) sHot[nHot - 1] = ell[nHot - 1];
    }
    __syncthreads();
    // a warp owns a run of consecutive chunks: the open run at the end of a chunk is carried in registers into the next one,
    // so atomics are only needed for a segment that crosses the first or the last chunk boundary of the run
    const u64 c0 = ((u64)blockIdx.x * (kK1Threads >> 5) + (tid >> 5)) * chunksPerWarp;
    if (c0 >= I.nDeltaChunks) return;
    const u64 c1 = min(c0 + (u64)chunksPerWarp, I.nDeltaChunks);
    const u32 cs0 = __ldg(&I.chunkSeg[c0]);
    u32 segBase = cs0 & 0x7FFFFFFFu;        // segments that end before the current chunk
    bool outside = (cs0 >> 31) != 0;        // the next segment end closes a segment that began before this warp's run
    u64 clo = 0; int chi = 0, ccn = 0;      // open run carried over from the previous chunk (warp-uniform)
    for (u64 c = c0; c < c1; ++c) {
        const unsigned F = __ldg(&I.endMask[c * 32 + lane]);   // bit j: word j of this lane is the last delta of its node
        if (c + 2 < c1) {   // pull the chunk after next towards L2 while this one is processed (16 lines of words + 1 line of masks)
            if (lane < 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(I.dw + (c + 2) * kChunkWords + lane * 32));
            else if (lane == 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(I.endMask + (c + 2) * 32));
        }
        u32 w[16];
        {   // the chunk is stored lane-interleaved: 16-byte piece q of lane l sits at uint4 index 32 q + l (coalesced)
            const uint4* p = reinterpret_cast<const uint4*>(I.dw + c * kChunkWords) + lane;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 a = __ldg(p + 32 * q);
                w[4 * q] = a.x; w[4 * q + 1] = a.y; w[4 * q + 2] = a.z; w[4 * q + 3] = a.w;
            }
        }
        long long e[16];   // w = 2 * seed id + lost
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const u32 id = w[j] >> 1;
            if (id < (u32)kHotIds) e[j] = sHot[id];
            else { const int2 t = tex1Dfetch<int2>(W.ellTex, (int)id); e[j] = (long long)(((u64)(u32)t.y << 32) | (u32)t.x); }
        }
        // segment index of this lane's first segment end: segments ending in earlier chunks + in earlier lanes
        const unsigned nEnd = __popc(F);
        unsigned endsBefore = nEnd;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, endsBefore, d); if (lane >= (unsigned)d) endsBefore += o; }
        const unsigned endsInChunk = __shfl_sync(0xffffffffu, endsBefore, 31);
        endsBefore -= nEnd;
        const u32 segFirst = segBase + endsBefore;
        const unsigned endMask = __ballot_sync(0xffffffffu, nEnd != 0);
        // ---- walk the 16 words.  The positions of the lane's first and last segment end are known up front, so the running sum
        //      is simply captured there: head = sum up to the first end, tail = total - sum up to the last end, and for a lane
        //      with two ends the segment between them is the difference.  Three or more ends in one lane are rare (a re-walk).
        const int first = F ? __ffs(F) - 1 : 99, last = F ? 31 - __clz(F) : 99;
        long long acc = 0, headV = 0, lastV = 0; int cn = 0, headC = 0, lastC = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int m = -(int)(w[j] & 1u);                      // -1 when the seed is lost, 0 when gained
            const long long mm = (long long)m;
            acc += (e[j] ^ mm) - mm;
            cn += (int)(e[j] >> 32) ? (m | 1) : 0;                // present <=> e >= ln2 * 2^53, i.e. the high word is non-zero
            if (j == first) { headV = acc; headC = cn; }
            if (j == last) { lastV = acc; lastC = cn; }
        }
        if (nEnd == 2) { const long long sv = lastV - headV; segStore(W.segRec + segFirst + 1, (u64)sv, (int)(sv >> 63), lastC - headC); }
        if (__any_sync(0xffffffffu, nEnd > 2)) {
            if (nEnd > 2) {   // segments that begin and end inside the lane, one by one
                long long a2 = 0; int c2 = 0; unsigned k = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int m = -(int)(w[j] & 1u);
                    const long long mm = (long long)m;
                    a2 += (e[j] ^ mm) - mm;
                    c2 += (int)(e[j] >> 32) ? (m | 1) : 0;
                    if (F & (1u << j)) {
                        if (k) segStore(W.segRec + segFirst + k, (u64)a2, (int)(a2 >> 63), c2);
                        ++k; a2 = 0; c2 = 0;
                    }
                }
            }
        }
        acc -= lastV; cn -= lastC;   // trailing partial: everything after the last end (the whole lane when it has none)
        // ---- segmented inclusive scan over the lanes' trailing partials (a lane with a segment end restarts the run) ----
        u64 lo = (u64)acc; int hi = (int)(acc >> 63); int rc = cn;
        if (lane == 0 && !nEnd) { const u64 t = lo + clo; hi += chi + (t < lo ? 1 : 0); lo = t; rc += ccn; }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 olo = shflUpU64(lo, d);
            const int ohi = __shfl_up_sync(0xffffffffu, hi, d), oc = __shfl_up_sync(0xffffffffu, rc, d);
            // lanes (lane-d, lane] must all be free of segment ends for the value of lane-d to flow into this lane
            if (lane >= (unsigned)d && !(endMask & (((1u << d) - 1u) << (lane - d + 1)))) {
                const u64 nlo = lo + olo;
                hi += ohi + (nlo < lo ? 1 : 0);
                lo = nlo; rc += oc;
            }
        }
        // carry into this lane's first segment end = inclusive value of the previous lane (lane 0: the run carried in)
        u64 plo = shflUpU64(lo, 1);
        int phi = __shfl_up_sync(0xffffffffu, hi, 1), pc = __shfl_up_sync(0xffffffffu, rc, 1);
        if (lane == 0) { plo = clo; phi = chi; pc = ccn; }
        if (nEnd) {
            u64 flo = (u64)headV; int fhi = (int)(headV >> 63); int fc = headC;
            { const u64 t = flo + plo; fhi += phi + (t < flo ? 1 : 0); flo = t; fc += pc; }
            SegRec* dst = W.segRec + segFirst;
            if (outside && !(endMask & ltMask)) segAtomicAdd(dst, flo, fhi, fc);   // the segment began before this warp's run
            else segStore(dst, flo, fhi, fc);
        }
        if (endMask) outside = false;
        // the open run after the chunk's last segment end travels on in registers
        const bool open = !((__shfl_sync(0xffffffffu, F, 31) >> 15) & 1u);
        clo = shflU64(lo, 31); chi = __shfl_sync(0xffffffffu, hi, 31); ccn = __shfl_sync(0xffffffffu, rc, 31);
        if (!open) { clo = 0; chi = 0; ccn = 0; }
        segBase += endsInChunk;
    }
    if (lane == 0 && segBase < I.nSeg) segAtomicAdd(W.segRec + segBase, clo, chi, ccn);   // run ends inside a segment (no-op when zero)
}
void launchDeltas(DevIndexView I, WorkspaceView W, int nSM, cudaStream_t st) {
    if (I.nDeltaChunks == 0) return;
    const u64 wpb = kK1Threads / 32;
    const u64 warps = (u64)nSM * wpb;   // one block per SM
    const u32 per = (u32)((I.nDeltaChunks + warps - 1) / warps);
    const u64 grid = ((I.nDeltaChunks + per - 1) / per + wpb - 1) / wpb;
    const size_t sm = (size_t)kHotIds * sizeof(long long);
    cudaFuncSetAttribute(node_deltas, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);   // per device: set on every launch (cheap)
    noteLaunch(), node_deltas<<<(unsigned)grid, kK1Threads, sm, st>>>(I, W, per);
}

// ------------------------------------------------------------------------------------------------------
// general deltas (a genome count >= 2 on either side; 76 of 2.4 M deltas on sars_20000): the reference's formula per delta
// (placement.cpp:282-339), exact fx128 atomics into the owning node's record, then the tree prefix of those records as an
// inclusive scan over DFS-interval events (+node at its DFS index, -node where its subtree ends).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ Acc5 accZero() { Acc5 a; a.f[0] = a.f[1] = a.f[2] = a.f[3] = fxZero(); a.pres = 0; return a; }
__device__ __forceinline__ Acc5 accAdd(const Acc5& a, const Acc5& b) {
    Acc5 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.f[i] = fxAdd(a.f[i], b.f[i]);
    r.pres = a.pres + b.pres; return r;
}
__device__ __forceinline__ Acc5 accNeg(const Acc5& a) {
    Acc5 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.f[i] = fxNeg(a.f[i]);
    r.pres = -a.pres; return r;
}
__device__ __forceinline__ Acc5 accShflUp(const Acc5& a, int d) {
    Acc5 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) { r.f[i].lo = shflUpU64(a.f[i].lo, d); r.f[i].hi = (i64)shflUpU64((u64)a.f[i].hi, d); }
    r.pres = (i64)shflUpU64((u64)a.pres, d); return r;
}
__device__ __forceinline__ void accStore(u64* p, const Acc5& a) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { p[2 * i] = a.f[i].lo; p[2 * i + 1] = (u64)a.f[i].hi; }
    p[8] = (u64)a.pres;
}
__device__ __forceinline__ Acc5 accLoad(const u64* p) {
    Acc5 a;
#pragma unroll
    for (int i = 0; i < 4; ++i) { a.f[i].lo = p[2 * i]; a.f[i].hi = (i64)p[2 * i + 1]; }
    a.pres = (i64)p[8]; return a;
}
__global__ void __launch_bounds__(256) gen_deltas(DevIndexView I, WorkspaceView W) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < I.nGenDeltas; i += gridDim.x * blockDim.x) {
        const long long e = __ldg(&W.ell[I.genId[i]]);
        if (!e) continue;
        const u32 pcv = I.genPc[i];
        const int p = (int)(short)(pcv & 0xFFFFu), c = (int)(short)(pcv >> 16);
        const DeltaTerms t = deltaTerms((double)e * kEllInvScale, p, c, p > 0 ? __ldg(&I.log1pSmall[p]) : 0.0, c > 0 ? __ldg(&I.log1pSmall[c]) : 0.0);
        u64* r = W.genRec + (size_t)I.genSlot[i] * kGenWords;
        fxAtomicAdd(r + 0, fxFromDouble(t.raw)); fxAtomicAdd(r + 2, fxFromDouble(t.cos));
        fxAtomicAdd(r + 4, fxFromDouble(t.wc)); fxAtomicAdd(r + 6, fxFromDouble(t.cont));
        if (t.pres) atomicAdd(reinterpret_cast<unsigned long long*>(r + 8), (unsigned long long)(long long)t.pres);
    }
}
// single block: inclusive scan of the signed node records in event order
__global__ void __launch_bounds__(256) gen_prefix(DevIndexView I, WorkspaceView W) {
    __shared__ u64 sWarp[8 * kGenWords];
    __shared__ u64 sCarry[kGenWords];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Acc5 carry = accZero();
    for (u32 base = 0; base < I.nEvents; base += 256) {
        const u32 j = base + threadIdx.x;
        Acc5 v = accZero();
        if (j < I.nEvents) {
            const u32 s = I.evSlot[j];
            v = accLoad(W.genRec + (size_t)(s & 0x7FFFFFFFu) * kGenWords);
            if (s >> 31) v = accNeg(v);
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const Acc5 o = accShflUp(v, d); if (lane >= d) v = accAdd(v, o); }
        if (lane == 31) accStore(sWarp + kGenWords * warp, v);
        __syncthreads();
        Acc5 pre = carry;
        for (int q = 0; q < warp; ++q) pre = accAdd(pre, accLoad(sWarp + kGenWords * q));
        v = accAdd(v, pre);
        if (j < I.nEvents) accStore(W.evPrefix + (size_t)j * kGenWords, v);
        if (threadIdx.x == 255) accStore(sCarry, v);
        __syncthreads();
        carry = accLoad(sCarry);
        __syncthreads();
    }
}
void launchGeneral(DevIndexView I, WorkspaceView W, cudaStream_t st) {
    if (I.nGenNodes == 0) return;
    cudaMemsetAsync(W.genRec, 0, (size_t)I.nGenNodes * kGenWords * sizeof(u64), st);
    unsigned g = (I.nGenDeltas + 255) / 256; if (g > 148 * 8) g = 148 * 8;
    noteLaunch(), gen_deltas<<<g, 256, 0, st>>>(I, W);
    noteLaunch(), gen_prefix<<<1, 256, 0, st>>>(I, W);
}

// ------------------------------------------------------------------------------------------------------
// K2 prefix_scores: A[v] = sum of the segment records over the root->v path, exact (96-bit integers + counts).
// Tile = kTileNodesK2 consecutive DFS nodes, 4 per thread.  The carry-in of a tile is the path root -> parent(first node),
// whose records were all written by K1, so every tile is independent (no inter-CTA dependency):
//   1. inclusive scan along the precomputed ancestor chain -> A[ancestor j]   (shared memory; global for very deep chains)
//   2. d'[w] = rec[w] (+ A[parent(w)] when the parent lies outside the tile)
//   3. Euler-tour difference inside the tile: diff[w] = d'[w] - sum_{u in tile, subtree(u) ends right before w} d'[u],
//      built with one shared-memory reduction per node (u subtracts itself at subEnd[u]): perfectly balanced, no lists.
//      The 96-bit values are split into carry-free limbs {sum of low 32-bit pieces, sum of the upper pieces} for this.
//   4. inclusive scan of diff = A[w]; add the general-delta prefix (event lookup) when the index has any; scores; store
// ------------------------------------------------------------------------------------------------------
struct Seg3 { u64 lo; int hi; int cnt; };
__device__ __forceinline__ Seg3 s3Zero() { Seg3 z; z.lo = 0; z.hi = 0; z.cnt = 0; return z; }
__device__ __forceinline__ Seg3 s3Add(const Seg3& a, const Seg3& b) {
    Seg3 r; r.lo = a.lo + b.lo; r.hi = a.hi + b.hi + (r.lo < a.lo ? 1 : 0); r.cnt = a.cnt + b.cnt; return r;
}
__device__ __forceinline__ Seg3 s3Sub(const Seg3& a, const Seg3& b) {
    Seg3 r; r.lo = a.lo - b.lo; r.hi = a.hi - b.hi - (a.lo < b.lo ? 1 : 0); r.cnt = a.cnt - b.cnt; return r;
}
__device__ __forceinline__ Seg3 s3ShflUp(const Seg3& a, int d) {
    Seg3 r; r.lo = shflUpU64(a.lo, d); r.hi = __shfl_up_sync(0xffffffffu, a.hi, d); r.cnt = __shfl_up_sync(0xffffffffu, a.cnt, d); return r;
}
__device__ __forceinline__ Seg3 s3Load(const SegRec* p) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    Seg3 r; r.lo = (u64)v.x | ((u64)v.y << 32); r.hi = (int)v.z; r.cnt = (int)v.w; return r;
}
__device__ __forceinline__ void s3Store(SegRec* p, const Seg3& a) {
    *reinterpret_cast<uint4*>(p) = make_uint4((u32)a.lo, (u32)(a.lo >> 32), (u32)a.hi, (u32)a.cnt);
}
__device__ __forceinline__ Seg3 s3OfNode(const DevIndexView& I, const WorkspaceView& W, u32 v) {
    const u32 s = __ldg(&I.nodeSeg[v]);
    return s == kNone ? s3Zero() : s3Load(W.segRec + s);
}
// block-wide inclusive scan over 256 threads (one Seg3 each); sWarp: 8 records of shared scratch
__device__ __forceinline__ Seg3 blockInclusiveScan(Seg3 v, SegRec* sWarp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Seg3 o = s3ShflUp(v, d);
        if (lane >= d) v = s3Add(v, o);
    }
    __syncthreads();
    if (lane == 31) s3Store(sWarp + warp, v);
    __syncthreads();
    Seg3 pre = s3Zero();
    for (int q = 0; q < warp; ++q) pre = s3Add(pre, s3Load(sWarp + q));
    return s3Add(v, pre);
}
constexpr int kChainSmem = 256;   // ancestor chains up to this length are kept in shared memory
constexpr int kK2Per = kTileNodesK2 / 256;

__global__ void __launch_bounds__(256, 4) prefix_scores(DevIndexView I, WorkspaceView W, PlaceOpts O) {
    __shared__ long long sA[kTileNodesK2];   // diff, low limb: sum of the low 32-bit pieces
    __shared__ long long sB[kTileNodesK2];   // diff, upper limb: sum of (value >> 32)
    __shared__ int sC[kTileNodesK2];         // diff, count
    __shared__ SegRec sChain[kChainSmem];
    __shared__ SegRec sWarp[8];
    __shared__ SegRec sCarry;
    const u32 tile = blockIdx.x;
    const u32 a0 = I.nodeBegin + tile * kTileNodesK2;
    const u32 a1 = min(a0 + (u32)kTileNodesK2, I.nodeEnd);
    const int tid = threadIdx.x;

    // 1. ancestor chain
    const u32 cb = I.chainOff[tile], ce = I.chainOff[tile + 1];
    SegRec* const chainA = (ce - cb <= (u32)kChainSmem) ? sChain : W.chainA + cb;
    if (ce - cb <= 32u) {   // the common case: one warp, no block barriers
        if (tid < 32) {
            Seg3 v = s3Zero();
            if (cb + tid < ce) v = s3OfNode(I, W, I.chainNodes[cb + tid]);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const Seg3 o = s3ShflUp(v, d); if (tid >= d) v = s3Add(v, o); }
            if (cb + tid < ce) s3Store(chainA + tid, v);
        }
    } else {
        Seg3 carry = s3Zero();
        for (u32 base = cb; base < ce; base += 256) {
            Seg3 v = s3Zero();
            const u32 j = base + tid;
            if (j < ce) v = s3OfNode(I, W, I.chainNodes[j]);
            v = blockInclusiveScan(v, sWarp);
            v = s3Add(v, carry);
            if (j < ce) s3Store(chainA + (j - cb), v);
            if (tid == 255) s3Store(&sCarry, v);
            __syncthreads();
            carry = s3Load(&sCarry);
            __syncthreads();
        }
    }
    // 2. d' of this thread's consecutive nodes; their own records are fetched while the chain settles
    const u32 w0 = a0 + kK2Per * tid;
    Seg3 d[kK2Per]; u32 cslot[kK2Per], send[kK2Per];
#pragma unroll
    for (int q = 0; q < kK2Per; ++q) {
        const u32 w = w0 + q;
        d[q] = s3Zero(); cslot[q] = kNone; send[q] = kNone;
        if (w < a1) { d[q] = s3OfNode(I, W, w); cslot[q] = __ldg(&I.carrySlot[w]); send[q] = __ldg(&I.subEnd[w]); }
    }
    __syncthreads();
    if (ce - cb > (u32)kChainSmem) __threadfence_block();
#pragma unroll
    for (int q = 0; q < kK2Per; ++q) {
        if (cslot[q] != kNone) d[q] = s3Add(d[q], s3Load(chainA + cslot[q]));
        const int li = kK2Per * tid + q;
        const long long up = (long long)((d[q].lo >> 32) | ((u64)(i64)d[q].hi << 32));
        sA[li] = (long long)(d[q].lo & 0xFFFFFFFFULL); sB[li] = up; sC[li] = d[q].cnt;
    }
    __syncthreads();
    // 3. every node subtracts itself where its subtree ends (if that is inside the tile)
#pragma unroll
    for (int q = 0; q < kK2Per; ++q) {
        if (send[q] < a1) {
            const int li = (int)(send[q] - a0);
            const long long up = (long long)((d[q].lo >> 32) | ((u64)(i64)d[q].hi << 32));
            atomicAdd(reinterpret_cast<unsigned long long*>(&sA[li]), (unsigned long long)(-(long long)(d[q].lo & 0xFFFFFFFFULL)));
            atomicAdd(reinterpret_cast<unsigned long long*>(&sB[li]), (unsigned long long)(-up));
            if (d[q].cnt) atomicAdd(&sC[li], -d[q].cnt);
        }
    }
    __syncthreads();
    // 4. scan of the differences
    Seg3 mine = s3Zero();
#pragma unroll
    for (int q = 0; q < kK2Per; ++q) {
        const int li = kK2Per * tid + q;
        const long long a = sA[li], b = sB[li];
        // value = a + b * 2^32 as a signed 96-bit integer
        Seg3 x; x.lo = (u64)a; x.hi = (int)(a >> 63); x.cnt = sC[li];
        Seg3 y; y.lo = (u64)b << 32; y.hi = (int)(b >> 32); y.cnt = 0;
        d[q] = s3Add(x, y);
        mine = s3Add(mine, d[q]);
    }
    const Seg3 incl = blockInclusiveScan(mine, sWarp);
    Seg3 run = s3Sub(incl, mine);
    const SampleScalars S = *W.scalars;
    double out[kK2Per * 5];   // the thread's 4 nodes are 160 contiguous bytes of the score array
#pragma unroll
    for (int q = 0; q < kK2Per; ++q) {
        const u32 w = w0 + q;
#pragma unroll
        for (int m = 0; m < 5; ++m) out[5 * q + m] = 0.0;
        if (w < a1) {
            run = s3Add(run, d[q]);
            double num[5];
            const u32 ne = I.nGenNodes ? __ldg(&I.evIdx[w]) : 0u;
            if (ne) { const Acc5 g = accLoad(W.evPrefix + (size_t)(ne - 1) * kGenWords); nodeNumerators(run.lo, run.hi, run.cnt, &g, I.ln2, num); }
            else nodeNumerators(run.lo, run.hi, run.cnt, nullptr, I.ln2, num);
            nodeScores(num[0], num[1], num[2], num[3], num[4], I.gMag[w], S, out + 5 * q);
            if (W.metrics) {
                double* m = W.metrics + (size_t)w * 5;
                m[0] = num[0]; m[1] = num[1]; m[2] = num[2]; m[3] = num[3]; m[4] = num[4];
            }
        }
    }
    double* o = W.scores + (size_t)w0 * 5;
    if (w0 + kK2Per <= a1 && (w0 & 1u) == 0) {   // 16-byte aligned: ten 16-byte stores instead of twenty 8-byte ones
#pragma unroll
        for (int i = 0; i < kK2Per * 5 / 2; ++i) reinterpret_cast<double2*>(o)[i] = make_double2(out[2 * i], out[2 * i + 1]);
    } else {
#pragma unroll
        for (int q = 0; q < kK2Per; ++q)
            if (w0 + q < a1) {
#pragma unroll
                for (int m = 0; m < 5; ++m) o[5 * q + m] = out[5 * q + m];
            }
    }
}
void launchPrefixScores(DevIndexView I, WorkspaceView W, PlaceOpts O, cudaStream_t st) {
    if (I.nK2Tiles == 0) return;
    noteLaunch(), prefix_scores<<<I.nK2Tiles, 256, 0, st>>>(I, W, O);
}

// ------------------------------------------------------------------------------------------------------
// selection.  The reference scores nodes in BFS order and keeps (best, tie list) with a relative tolerance:
// a node becomes the new best only if score > best + max(best*1e-4, 1e-9) (placement.cpp:355-371).  Every such
// "event" is a strict prefix maximum of the BFS-ordered score sequence, so:
//   bfs_gather   scores -> BFS order (ineligible nodes get -1), per-block maxima
//   bfs_records  prefix maxima ("records"), a short list per metric
//   chain        replays the tolerance chain over the records only
//   ties         nodes after the last event with score >= best - tol and > 0
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warpMaxF64(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmax(v, shflXorF64(v, d));
    return v;
}

__global__ void __launch_bounds__(256) bfs_gather(DevIndexView I, WorkspaceView W, PlaceOpts O, double* __restrict__ bfsScores) {
    __shared__ double sMax[8][5];
    const u32 base = blockIdx.x * kBfsBlock;
    double mx[5] = {-1.0, -1.0, -1.0, -1.0, -1.0};
    for (int q = 0; q < 4; ++q) {
        const u32 r = base + q * 256 + threadIdx.x;
        if (r < I.nShardNodes) {
            const u32 v = I.bfsNodes[r];
            const bool ok = (v != O.skipNode) && (!O.forceLeaf || I.isLeaf[v]);
            const double* s = W.scores + (size_t)v * 5;
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                const double x = ok ? s[m] : -1.0;
                bfsScores[(size_t)m * I.nShardNodes + r] = x;
                mx[m] = fmax(mx[m], x);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < 5; ++m) mx[m] = warpMaxF64(mx[m]);
    if ((threadIdx.x & 31) == 0)
        for (int m = 0; m < 5; ++m) sMax[threadIdx.x >> 5][m] = mx[m];
    __syncthreads();
    if (threadIdx.x < 5) {
        double v = sMax[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) v = fmax(v, sMax[w][threadIdx.x]);
        W.blockMax[(size_t)blockIdx.x * 5 + threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(256) bfs_records(DevIndexView I, WorkspaceView W, const double* __restrict__ bfsScores) {
    // grid: (nBfsBlocks, 5)
    __shared__ double sRed[8];
    __shared__ double sPrev;
    __shared__ double sThread[256];
    const int m = blockIdx.y;
    const u32 blk = blockIdx.x;
    // exclusive prefix maximum over earlier blocks (records must also be > 0)
    double pm = 0.0;
    for (u32 j = threadIdx.x; j < blk; j += 256) pm = fmax(pm, W.blockMax[(size_t)j * 5 + m]);
    pm = warpMaxF64(pm);
    if ((threadIdx.x & 31) == 0) sRed[threadIdx.x >> 5] = pm;
    __syncthreads();
    if (threadIdx.x == 0) { double v = sRed[0]; for (int w = 1; w < 8; ++w) v = fmax(v, sRed[w]); sPrev = v; }
    __syncthreads();
    const double prev = sPrev;
    if (W.blockMax[(size_t)blk * 5 + m] <= prev) return;  // no record in this block
    // 4 consecutive positions per thread
    const u32 r0 = blk * kBfsBlock + threadIdx.x * 4;
    double x[4]; double tmax = -1.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const u32 r = r0 + q;
        x[q] = r < I.nShardNodes ? bfsScores[(size_t)m * I.nShardNodes + r] : -1.0;
        tmax = fmax(tmax, x[q]);
    }
    sThread[threadIdx.x] = tmax;
    __syncthreads();
    double run = prev;
    for (int j = 0; j < (int)threadIdx.x; ++j) run = fmax(run, sThread[j]);  // 256x256/2 smem reads, rare blocks only
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const u32 r = r0 + q;
        if (r < I.nShardNodes && x[q] > run) {
            const unsigned o = atomicAdd(&W.acc->recordCount[m], 1u);
            if (o < W.recCap) {
                W.recRank[(size_t)m * W.recCap + o] = I.bfsRanks[r];
                W.recNode[(size_t)m * W.recCap + o] = I.bfsNodes[r];
                W.recScore[(size_t)m * W.recCap + o] = x[q];
            } else raiseFlag(W.acc, kOvfTable);
        }
        run = fmax(run, x[q]);
    }
}

// tolerance chain over the records of one metric (block m).  Records are unordered: every step finds the
// lowest-rank record after the last event that beats best + tol.
__global__ void __launch_bounds__(256) chain_select(WorkspaceView W, const u32* __restrict__ countOverride) {
    __shared__ ChainShared S;
    const int m = blockIdx.x;
    unsigned n = countOverride ? countOverride[m] : W.acc->recordCount[m];
    if (n > W.recCap) n = W.recCap;
    const u32* rk = W.recRank + (size_t)m * W.recCap;
    const u32* nd = W.recNode + (size_t)m * W.recCap;
    const double* sc = W.recScore + (size_t)m * W.recCap;
    const Selection s = chainReplay(S, n, [&](unsigned i, u32& r, double& x, u32& v) { r = rk[i]; x = sc[i]; v = nd[i]; return true; });
    if (threadIdx.x == 0) W.sel[m] = s;
}
void launchChain(WorkspaceView W, const u32* recCountOverride, cudaStream_t st) { noteLaunch(), chain_select<<<5, 256, 0, st>>>(W, recCountOverride); }

__global__ void __launch_bounds__(256) collect_ties(DevIndexView I, WorkspaceView W, const double* __restrict__ bfsScores) {
    const int m = blockIdx.y;
    const Selection s = W.sel[m];
    const double tol = fmax(s.best * 0.0001, 1e-9);
    const double lo = s.best - tol;
    if (W.blockMax[(size_t)blockIdx.x * 5 + m] < lo) return;
    for (int q = 0; q < 4; ++q) {
        const u32 r = blockIdx.x * kBfsBlock + q * 256 + threadIdx.x;
        if (r >= I.nShardNodes) continue;
        const u32 rank = I.bfsRanks[r];
        if (s.lastRank != kNone && rank <= s.lastRank) continue;
        const double x = bfsScores[(size_t)m * I.nShardNodes + r];
        if (x >= lo && x > 0.0) {
            const unsigned o = atomicAdd(&W.acc->tieCount[m], 1u);
            if (o < W.tieCap) W.tieNode[(size_t)m * W.tieCap + o] = I.bfsNodes[r]; else raiseFlag(W.acc, kOvfTable);
            if (o < (unsigned)kTieHead) W.tieHead[m * kTieHead + o] = I.bfsNodes[r];
        }
    }
}

// bfsScores lives right after blockMax in the workspace (allocated by the host as one buffer)
static double* bfsScoresOf(const DevIndexView& I, const WorkspaceView& W) { return W.blockMax + (size_t)I.nBfsBlocks * 5; }

void launchRecords(DevIndexView I, WorkspaceView W, PlaceOpts O, cudaStream_t st) {
    if (I.nBfsBlocks == 0) return;
    double* bs = bfsScoresOf(I, W);
    noteLaunch(), bfs_gather<<<I.nBfsBlocks, 256, 0, st>>>(I, W, O, bs);
    noteLaunch(), bfs_records<<<dim3(I.nBfsBlocks, 5), 256, 0, st>>>(I, W, bs);
}
void launchTies(DevIndexView I, WorkspaceView W, PlaceOpts O, cudaStream_t st) {
    if (I.nBfsBlocks == 0) return;
    noteLaunch(), collect_ties<<<dim3(I.nBfsBlocks, 5), 256, 0, st>>>(I, W, bfsScoresOf(I, W));
}

}  // namespace pm
