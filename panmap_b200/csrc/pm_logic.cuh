// pm_logic.cuh -- arithmetic shared by every kernel of the placement path, written as PM_HD
// (__host__ __device__) so tests/hostcheck can run the exact same code on the CPU against the oracle
// before it ever reaches a GPU.  Nothing here is a CPU fallback: the product library only instantiates it
// inside __global__ kernels.
//
// Reference semantics restated here (paths relative to /root/reference/):
//   seeding.hpp:86-120        chash / comp / rol / ror
//   seeding.cpp:47-229        rollingSyncmers (rolling k-mer and s-mer hashes on both strands, window minima)
//   placement.cpp:1598-1686   per-read k-min-mer construction (l > 1) and trim filter
//   placement.cpp:242-345     per-delta contributions of computeChildMetrics
//   placement.hpp:120-149     the five score getters
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PM_HD __host__ __device__ __forceinline__
#else
#define PM_HD inline
#endif

namespace pm {

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;

constexpr u64 kHashA = 0x3c8bfbb395c60474ULL;
constexpr u64 kHashC = 0x3193c18562a02b4cULL;
constexpr u64 kHashG = 0x20323ed082572324ULL;
constexpr u64 kHashT = 0x295549f54be24456ULL;
constexpr u64 kEmptyKey = 0xFFFFFFFFFFFFFFFFULL;
constexpr int kMaxK = 32;  // 4-bit base codes in a 128-bit history register
constexpr int kTileNodesK2 = 1024;  // prefix_scores: consecutive DFS nodes per tile (4 per thread)

PM_HD u64 rol64(u64 h, unsigned r) { r &= 63u; return r ? (h << r) | (h >> (64u - r)) : h; }
PM_HD u64 ror64(u64 h, unsigned r) { r &= 63u; return r ? (h >> r) | (h << (64u - r)) : h; }
PM_HD u64 rol1(u64 h) { return (h << 1) | (h >> 63); }
PM_HD u64 ror1(u64 h) { return (h >> 1) | (h << 63); }
PM_HD u64 umin64(u64 a, u64 b) { return a < b ? a : b; }

// base code: A 0, C 1, G 2, T 3 (either case), anything else 4 (ambiguous, hash contribution 0)
PM_HD unsigned baseCode(unsigned char c) {
    switch (c) {
        case 'A': case 'a': return 0; case 'C': case 'c': return 1;
        case 'G': case 'g': return 2; case 'T': case 't': return 3;
        default: return 4;
    }
}
PM_HD u64 codeHash(unsigned code) {
    return code == 0 ? kHashA : code == 1 ? kHashC : code == 2 ? kHashG : code == 3 ? kHashT : 0ULL;
}

// Per-(k,s) rotated constants, 8 entries each (4..7 are zero so that "ambiguous / not yet seen" needs no branch).
struct SeedTables {
    u64 fwdNew[8];   // c(b)
    u64 fwdOldK[8];  // rol(c(b), k)
    u64 fwdOldS[8];  // rol(c(b), s)
    u64 revNewK[8];  // rol(c(comp b), k-1)
    u64 revNewS[8];  // rol(c(comp b), s-1)
    u64 revOld[8];   // ror(c(comp b), 1)
};
inline void buildSeedTables(SeedTables& T, int k, int s) {
    for (unsigned c = 0; c < 8; ++c) {
        const u64 f = codeHash(c), r = c < 4 ? codeHash(3 - c) : 0ULL;
        T.fwdNew[c] = f;
        T.fwdOldK[c] = rol64(f, (unsigned)k);
        T.fwdOldS[c] = rol64(f, (unsigned)s);
        T.revNewK[c] = rol64(r, (unsigned)(k - 1));
        T.revNewS[c] = rol64(r, (unsigned)(s - 1));
        T.revOld[c] = ror64(r, 1);
    }
}

// device image of the tables: element 0 = the SeedTables, the elements behind it hold the s = 8 rank table (syncmers_rank; zeros for s != 8)
constexpr size_t kSeedTableElems = 1 + ((size_t)(1u << 16) * 2 + sizeof(SeedTables) - 1) / sizeof(SeedTables);

struct SeederParams {
    int k, s, t, l, open;
    int w;           // k - s + 1
    unsigned rotK;   // k mod 64
    unsigned rotKL;  // k*l mod 64
    unsigned rotKL1; // k*(l-1) mod 64
    int trimStart, trimEnd;
    int maxLen;      // longest read of the launch when the host knows it (0 = unknown): lets syncmers_rank stage the reads in shared memory
};
inline SeederParams makeSeederParams(int k, int s, int t, int l, int open, int trimStart, int trimEnd) {
    SeederParams p;
    p.k = k; p.s = s; p.t = t; p.l = l; p.open = open; p.w = k - s + 1;
    p.rotK = (unsigned)k & 63u;
    p.rotKL = (unsigned)((long long)k * l) & 63u;
    p.rotKL1 = (unsigned)((long long)k * (l > 0 ? l - 1 : 0)) & 63u;
    p.trimStart = trimStart; p.trimEnd = trimEnd; p.maxLen = 0;
    return p;
}

// Sequential per-read seeder: feed bases one at a time; emits syncmers (and k-min-mers for l > 1).
// Ring storage (4*w + max(l,1) u64 words) is supplied by the caller as a strided view so that a CUDA block can
// lay the rings out [slot][thread] in shared memory (conflict-free) and the host check can use a plain array.
// STRIDE is a compile-time constant (the CUDA block size) so that ring addressing is shifts, not 64-bit multiplies.
template <int STRIDE>
struct ReadSeederT {
    u64 fk, rk, fs, rs;      // rolling k-mer / s-mer hashes, both strands
    u64 hist2;               // last 32 base codes, 2 bits each, newest in bits 0..1
    unsigned histAmb;        // last 32 "ambiguous" flags, newest in bit 0
    u64 preF, preR;          // running minima of the current block (van Herk / Gil-Werman sliding minimum)
    u64 kmF, kmR;            // rolling k-min-mer hashes
    int lastAmb;             // index of the most recent ambiguous base
    int slot;                // (s-mer index) mod w
    int synCount;            // in-range syncmers seen so far (saturates at l)
    int hslot;               // synCount mod l
    unsigned kmRot;          // (k * synCount) mod 64 during the first l syncmers
    u64 *pF, *pFsuf, *pR, *pRsuf, *pH;  // ring views

    PM_HD void reset(u64* ringBase, int w) {
        fk = rk = fs = rs = 0; kmF = kmR = 0;
        hist2 = 0; histAmb = 0xFFFFFFFFu;  // every past base "ambiguous": table entries 4..7 are zero
        preF = preR = kEmptyKey;
        lastAmb = -1; slot = -1; synCount = 0; hslot = 0; kmRot = 0;
        pF = ringBase; pFsuf = ringBase + (size_t)w * STRIDE; pR = ringBase + (size_t)2 * w * STRIDE;
        pRsuf = ringBase + (size_t)3 * w * STRIDE; pH = ringBase + (size_t)4 * w * STRIDE;
    }
    PM_HD unsigned codeBack(int dist) const {  // table index of the base `dist` positions before the one being pushed (1 <= dist <= 32)
        return (unsigned)((hist2 >> (2 * (dist - 1))) & 3ULL) | (((histAmb >> (dist - 1)) & 1u) << 2);
    }

    // Push base `code` (0..3 ACGT, >=4 ambiguous) at read position i (0-based).  Returns true when k-mer window i-k+1 is a
    // syncmer; hash / isReverse are then set.  (seeding.cpp:147-226 restated with no per-step rescans.)
    PM_HD bool pushBase(int i, unsigned code, const SeedTables& T, const SeederParams& P, u64& hash, bool& isReverse) {
        const unsigned oldK = codeBack(P.k), oldS = codeBack(P.s);
        const unsigned tc = code & 7u;
        fk = rol1(fk) ^ T.fwdOldK[oldK] ^ T.fwdNew[tc];
        rk = ror1(rk) ^ T.revOld[oldK] ^ T.revNewK[tc];
        fs = rol1(fs) ^ T.fwdOldS[oldS] ^ T.fwdNew[tc];
        rs = ror1(rs) ^ T.revOld[oldS] ^ T.revNewS[tc];
        hist2 = (hist2 << 2) | (u64)(code & 3u);
        histAmb = (histAmb << 1) | (code >= 4 ? 1u : 0u);
        if (code >= 4) lastAmb = i;
        if (i < P.s - 1) return false;
        const int w = P.w;
        slot = (slot + 1 == w) ? 0 : slot + 1;
        pF[slot * STRIDE] = fs; pR[slot * STRIDE] = rs;
        if (slot == 0) { preF = fs; preR = rs; } else { preF = umin64(preF, fs); preR = umin64(preR, rs); }
        bool syn = false;
        if (i >= P.k - 1) {
            const int pslot = (slot + 1 == w) ? 0 : slot + 1;  // slot of the oldest s-mer of the window
            u64 mf = preF, mr = preR;
            if (pslot != 0) { mf = umin64(mf, pFsuf[pslot * STRIDE]); mr = umin64(mr, pRsuf[pslot * STRIDE]); }
            int ia = pslot + P.t; if (ia >= w) ia -= w;            // s-mer p+t
            int ib = slot - P.t; if (ib < 0) ib += w;              // s-mer p+k-s-t
            bool fsyn, rsyn;
            if (P.open) { fsyn = pF[ia * STRIDE] == mf; rsyn = pR[ib * STRIDE] == mr; }
            else {
                fsyn = (pF[ia * STRIDE] == mf) || (pF[ib * STRIDE] == mf);
                rsyn = (pR[ib * STRIDE] == mr) || (pR[ia * STRIDE] == mr);
            }
            syn = (i - lastAmb >= P.k) && (fsyn || rsyn) && (fk != rk);
            hash = umin64(fk, rk);
            isReverse = rk < fk;
        }
        if (slot == w - 1) {  // block complete: suffix minima for the windows that straddle into the next block
            u64 a = kEmptyKey, b = kEmptyKey;
            for (int q = w - 1; q >= 0; --q) {
                a = umin64(a, pF[q * STRIDE]); pFsuf[q * STRIDE] = a;
                b = umin64(b, pR[q * STRIDE]); pRsuf[q * STRIDE] = b;
            }
        }
        return syn;
    }

    // A syncmer (start position pos, canonical hash h) of a read of length len was found: apply the trim filter
    // and the k-min-mer construction.  Returns true when a seed is to be counted (placement.cpp:1627-1682).
    PM_HD bool pushSyncmer(int pos, int len, u64 h, const SeederParams& P, u64& seed) {
        if (pos < P.trimStart || pos > len - P.trimEnd - P.k) return false;
        if (P.l <= 1) { seed = h; return true; }
        const int l = P.l;
        if (synCount < l) {
            kmF = rol64(kmF, P.rotK) ^ h;
            kmR ^= rol64(h, kmRot);
            kmRot = (kmRot + P.rotK) & 63u;
            ++synCount;
        } else {
            const u64 prev = pH[hslot * STRIDE];
            kmF = rol64(kmF, P.rotK) ^ rol64(prev, P.rotKL) ^ h;
            kmR = ror64(kmR, P.rotK) ^ ror64(prev, P.rotK) ^ rol64(h, P.rotKL1);
        }
        pH[hslot * STRIDE] = h;
        hslot = (hslot + 1 == l) ? 0 : hslot + 1;
        if (synCount >= l && kmF != kmR) { seed = umin64(kmF, kmR); return true; }
        return false;
    }
};
PM_HD int seederRingWords(int k, int s, int l) { return 4 * (k - s + 1) + (l > 1 ? l : 1); }

// ---- closed syncmers with s = 8 by RANK (syncmers_rank kernel; seeding.cpp:147-226) --------------------------------------------
// Whether a k-mer is a closed syncmer depends only on the ORDER of its s-mer hashes (does the oldest or the newest of the k-s+1 s-mers
// attain the window minimum), never on their values.  For s = 8 there are 4^8 = 65,536 s-mers without an ambiguous base, so the order is
// a table: rank[idx] = dense rank of the forward hash of the 8-mer idx (first base in the two top bits) among all 8-mer hashes -- equal
// hashes get equal ranks, so every <= / == the reference evaluates on 64-bit hashes gives the same answer on 16-bit ranks.  The reverse
// strand needs no second table: the reverse hash of an s-mer IS the forward hash of its reverse complement (XOR_i rol(c(comp b_i), i)),
// so its rank is rank[index of the reverse complement].  s-mers that contain an ambiguous base have no rank; every k-mer window that
// holds one is rejected anyway (seeding.cpp: the window must be free of ambiguous bases), and a window's test only looks at its own s-mers.
// Both strands' ranks travel in one 32-bit word (forward low, reverse high): one 16x2 minimum instruction (VIMNMX.U16x2 on sm_90+,
// which also returns the two "a <= b" predicates) serves both strands.
constexpr int kRankS = 8;
constexpr unsigned kRankEntries = 1u << (2 * kRankS);
inline void buildSmerRanks(uint16_t* rank) {   // host, once per index: 65,536 hashes, sort, dense rank
    struct E { u64 h; unsigned idx; };
    E* e = new E[kRankEntries];
    for (unsigned idx = 0; idx < kRankEntries; ++idx) {
        u64 h = 0;
        for (int i = 0; i < kRankS; ++i) h ^= rol64(codeHash((idx >> (2 * (kRankS - 1 - i))) & 3u), (unsigned)(kRankS - 1 - i));
        e[idx].h = h; e[idx].idx = idx;
    }
    // plain shell sort keeps this header free of <algorithm>; 65,536 entries once per index
    for (unsigned gap = kRankEntries / 2; gap > 0; gap /= 2)
        for (unsigned i = gap; i < kRankEntries; ++i) {
            const E x = e[i]; unsigned j = i;
            for (; j >= gap && e[j - gap].h > x.h; j -= gap) e[j] = e[j - gap];
            e[j] = x;
        }
    unsigned r = 0;
    for (unsigned i = 0; i < kRankEntries; ++i) { if (i && e[i].h != e[i - 1].h) ++r; rank[e[i].idx] = (uint16_t)r; }
    delete[] e;
}
inline void buildSeedTableImage(SeedTables* img, int k, int s) {   // img[kSeedTableElems]
    memset(img, 0, kSeedTableElems * sizeof(SeedTables));
    buildSeedTables(img[0], k, s);
    if (s == kRankS) buildSmerRanks(reinterpret_cast<uint16_t*>(img + 1));
}
// cur holds four base codes, one per byte (byte j = base 4n + j; codes >= 4 are ambiguous and enter as code & 3).  HF collects the last
// 16 bases as 2-bit codes, newest in the low bits; HR their complements, newest in the HIGH bits.  One multiply gathers the four 2-bit
// fields of a word into its top byte (the partial products land on disjoint bits), one byte permute shifts it into the history.
PM_HD unsigned bytePerm(unsigned a, unsigned b, unsigned sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const u64 v = ((u64)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
#endif
}
PM_HD void rankPushWord(unsigned cur, unsigned& HF, unsigned& HR) {
    const unsigned x = cur & 0x03030303u;
    HF = bytePerm(HF, x * 0x40100401u, 0x2107u);                    // HF << 8 | (c0 c1 c2 c3)
    HR = bytePerm(HR, (x ^ 0x03030303u) * 0x01041040u, 0x7321u);    // HR >> 8 | (c3' c2' c1' c0') << 24
}
// byte offsets into the u16 rank table of the s-mer that ENDS at base ph (0..3) of the newest word
PM_HD unsigned rankAddrF(unsigned HF, int ph) { return (ph == 3 ? HF << 1 : HF >> (5 - 2 * ph)) & 0x1FFFEu; }
PM_HD unsigned rankAddrR(unsigned HR, int ph) { return (HR >> (9 + 2 * ph)) & 0x1FFFEu; }
PM_HD unsigned minU16x2(unsigned a, unsigned b) {
#if defined(__CUDA_ARCH__)
    return __vminu2(a, b);
#else
    const unsigned lo = (a & 0xFFFFu) < (b & 0xFFFFu) ? (a & 0xFFFFu) : (b & 0xFFFFu), hi = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
    return lo | (hi << 16);
#endif
}
PM_HD unsigned minU16x2Le(unsigned a, unsigned b, bool& leHi, bool& leLo) {   // per-half minimum and "a <= b"
#if defined(__CUDA_ARCH__)
    return __vibmin_u16x2(a, b, &leHi, &leLo);
#else
    leHi = (a >> 16) <= (b >> 16); leLo = (a & 0xFFFFu) <= (b & 0xFFFFu);
    return minU16x2(a, b);
#endif
}
// Sliding minimum over the last W s-mers in blocks of W (van Herk / Gil-Werman): raw[] = the packed ranks of the current block so far
// (slots > j still hold the previous block's), suf[q] = minimum of the previous block's slots q..W-1, pre = minimum of the current block.
// step(j, v): s-mer v lands in slot j; true when the window of W s-mers ending here has its minimum at the oldest or the newest s-mer, on
// either strand.  j must be a compile-time constant after unrolling so that raw / suf stay in registers.
template <int W>
struct RankWindow {
    unsigned raw[W], suf[W], pre;
    PM_HD void reset() { for (int q = 0; q < W; ++q) { raw[q] = 0; suf[q] = 0; } pre = 0; }
    PM_HD bool step(int j, unsigned v) {
        const int pslot = (j + 1 == W) ? 0 : j + 1;                 // slot of the oldest s-mer of the window
        const unsigned t = j == 0 ? suf[1] : (pslot == 0 ? pre : minU16x2(pre, suf[pslot]));   // everything but the newest
        bool nh, nl, oh, ol;
        const unsigned m = minU16x2Le(v, t, nh, nl);                // newest <= rest: it attains the minimum
        (void)minU16x2Le(raw[pslot], m, oh, ol);                    // oldest <= minimum: it attains it (slot 0: this block's first)
        pre = j == 0 ? v : minU16x2(pre, v);
        raw[j] = v;
        if (j == W - 1) {
            unsigned a = v; suf[W - 1] = a;
            for (int q = W - 2; q >= 1; --q) { a = minU16x2(raw[q], a); suf[q] = a; }
        }
        return nh || nl || oh || ol;
    }
};

// ---- exact accumulation: signed 128-bit fixed point, 64 fractional bits --------------------------------
// Every f64 quantity that is summed across nodes (tree prefix) or across table slots (read magnitudes) is first
// truncated to a multiple of 2^-64 and then added as an integer, so sums are associative: the Euler-tour
// +delta/-delta cancellation is exact and results do not depend on tile shape, GPU count or atomics order.
// |x| < 2^62 is required (numerators are bounded by (#seeds) * log1p(count) << 2^62).
struct fx128 { u64 lo; i64 hi; };
PM_HD fx128 fxZero() { fx128 z; z.lo = 0; z.hi = 0; return z; }
PM_HD fx128 fxAdd(fx128 a, fx128 b) {
    fx128 r; r.lo = a.lo + b.lo; r.hi = (i64)((u64)a.hi + (u64)b.hi + (r.lo < a.lo ? 1ULL : 0ULL)); return r;
}
PM_HD fx128 fxNeg(fx128 a) {
    fx128 r; r.lo = ~a.lo + 1ULL; r.hi = (i64)(~(u64)a.hi + (r.lo == 0 ? 1ULL : 0ULL)); return r;
}
PM_HD fx128 fxSub(fx128 a, fx128 b) { return fxAdd(a, fxNeg(b)); }
PM_HD fx128 fxMulU64(fx128 a, u64 n) {   // a >= 0 times a count (the product must stay below 2^126)
#if defined(__CUDA_ARCH__)
    const u64 carry = __umul64hi(a.lo, n);
#else
    const u64 carry = (u64)(((unsigned __int128)a.lo * n) >> 64);
#endif
    fx128 r; r.lo = a.lo * n; r.hi = (i64)((u64)a.hi * n + carry); return r;
}
PM_HD fx128 fxFromInt(i64 v) { fx128 r; r.lo = 0; r.hi = v; return r; }
PM_HD u64 dblBits(double x) {
#if defined(__CUDA_ARCH__)
    return (u64)__double_as_longlong(x);
#else
    union { double d; u64 u; } c; c.d = x; return c.u;
#endif
}
PM_HD double bitsDbl(u64 b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    union { double d; u64 u; } c; c.u = b; return c.d;
#endif
}
PM_HD fx128 fxFromDouble(double x) {  // truncates |x| toward zero at 2^-64; fx(-x) == -fx(x)
    const u64 b = dblBits(x);
    const int e = (int)((b >> 52) & 0x7FF);
    if (e == 0 || e == 0x7FF) return fxZero();  // zero / subnormal / non-finite contribute nothing
    const u64 m = (b & 0xFFFFFFFFFFFFFULL) | 0x10000000000000ULL;
    const int sh = e - 1075 + 64;  // value * 2^64 = m * 2^sh
    fx128 r;
    if (sh >= 64) { r.lo = 0; r.hi = (sh - 64 < 10) ? (i64)(m << (sh - 64)) : (i64)0x3FFFFFFFFFFFFFFFLL; }
    else if (sh > 0) { r.lo = m << sh; r.hi = (i64)(m >> (64 - sh)); }
    else if (sh == 0) { r.lo = m; r.hi = 0; }
    else if (sh > -53) { r.lo = m >> (-sh); r.hi = 0; }
    else return fxZero();
    return (b >> 63) ? fxNeg(r) : r;
}
PM_HD int clz64(u64 x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
PM_HD double fxToDouble(fx128 a) {  // round-to-nearest-even of the exact value
    const bool neg = a.hi < 0;
    if (neg) a = fxNeg(a);
    const u64 hi = (u64)a.hi, lo = a.lo;
    if (hi == 0 && lo == 0) return 0.0;
    // normalise to a 64-bit mantissa t with sticky bit, value = t * 2^ex
    u64 t; int ex;
    if (hi == 0) { const int z = clz64(lo); t = lo << z; ex = -64 - z; }
    else {
        const int z = clz64(hi);
        t = z ? (hi << z) | (lo >> (64 - z)) : hi;
        const u64 rest = z ? (lo << z) : lo;
        if (rest) t |= 1ULL;
        ex = -z;
    }
    // t has its top bit set: keep 53 bits, round to nearest even using the low 11 bits
    u64 mant = t >> 11;
    const u64 rem = t & 0x7FFULL;
    if (rem > 0x400ULL || (rem == 0x400ULL && (mant & 1ULL))) ++mant;
    int e2 = ex + 11;  // value = mant * 2^e2, mant in [2^52, 2^53]
    if (mant == (1ULL << 53)) { mant >>= 1; ++e2; }
    const u64 bits = ((u64)(e2 + 52 + 1023) << 52) | (mant & 0xFFFFFFFFFFFFFULL);
    const double d = bitsDbl(bits);
    return neg ? -d : d;
}

// ---- expected rounding drift of a sequential f64 sum (finish_scalars; see the comment there) ----------
// While the running sum is in binade [2^e, 2^(e+1)) every addition of x is rounded to a multiple of u = 2^(e-52), i.e. contributes
// rint(x/u)*u - x, and (for addends in random order) a share (hi-lo)/T of the additions happens in that binade.
constexpr int kBinades = 10;  // each lower binade carries 1/4 of the drift of the one above it
struct Binades { double fr[kBinades], u[kBinades], iu[kBinades]; };
PM_HD void makeBinades(double T, Binades& B) {   // T = the exact total
    const int eTop = (int)((dblBits(T) >> 52) & 0x7FF) - 1023;
    for (int j = 0; j < kBinades; ++j) {   // share of the additions that land in each binade, its ulp and 1/ulp
        const int ex = eTop - j;
        const double lo = j == kBinades - 1 ? 0.0 : bitsDbl((u64)(ex + 1023) << 52);
        const double top = bitsDbl((u64)(ex + 1024) << 52);
        const double hi = T < top ? T : top;
        B.fr[j] = (hi - lo) / T;
        B.u[j] = bitsDbl((u64)(ex - 52 + 1023) << 52);
        B.iu[j] = bitsDbl((u64)(52 - ex + 1023) << 52);
    }
}
PM_HD double driftOf(const Binades& B, double x) {   // expected drift contributed by ONE addition of x
    double a = 0.0;
    for (int j = 0; j < kBinades; ++j) a += B.fr[j] * (rint(x * B.iu[j]) * B.u[j] - x);
    return a;
}

// ---- per-node numerators from the two accumulator families --------------------------------------------
// Deltas with genome counts 0 <-> 1 ("fast", all but a handful) are summed as a signed 96-bit integer in units of 2^-53
// (every log1p(readCount >= 1) is an exact multiple of 2^-53) plus a signed count; deltas with a genome count >= 2
// ("general") keep the reference's five separate sums in fx128.  Both families are exact, so the tree prefix is
// associative and the only roundings are the final conversions below.
struct Acc5 {    // general-delta accumulators: raw, cos, wc, cont numerators + presence
    fx128 f[4];
    i64 pres;
};
PM_HD fx128 fxFromSeg(u64 lo, int hi) {  // 96-bit integer in units of 2^-53 -> fx128 (units of 2^-64)
    fx128 r; r.lo = lo << 11; r.hi = (i64)((((u64)(i64)hi) << 11) | (lo >> 53)); return r;
}
// out: logRawNum, logCosNum, presence, weightedContainmentNum, logContainmentNum (placement.hpp:108-118)
PM_HD void nodeNumerators(u64 lo, int hi, int cnt, const Acc5* g, double ln2, double* out) {
    const fx128 S = fxFromSeg(lo, hi);
    const double sd = fxToDouble(S);
    if (!g) { out[0] = sd; out[1] = sd * ln2; out[2] = (double)cnt; out[3] = (double)cnt; out[4] = sd; return; }
    out[0] = fxToDouble(fxAdd(S, g->f[0]));
    out[1] = fxToDouble(fxAdd(fxFromDouble(sd * ln2), g->f[1]));
    out[2] = (double)((i64)cnt + g->pres);
    out[3] = fxToDouble(fxAdd(fxFromInt((i64)cnt), g->f[2]));
    out[4] = fxToDouble(fxAdd(S, g->f[3]));
}

// ---- per-delta contribution (placement.cpp:282-339) ---------------------------------------------------
struct DeltaTerms { double raw, cos, wc, cont; int pres; };
// lr = log1p(readCount) of the seed (> 0), p/c = parent/child genome counts, logP/logC = log1p of them (0 when count<=0)
PM_HD DeltaTerms deltaTerms(double lr, int p, int c, double logP, double logC) {
    DeltaTerms d;
    d.pres = (int)((p == 0) & (c != 0)) - (int)((c == 0) & (p != 0));
    d.raw = (c > 0 ? lr / (double)c : 0.0) - (p > 0 ? lr / (double)p : 0.0);
    d.cos = lr * (logC - logP);
    d.wc = (c > 0 ? 1.0 / (double)c : 0.0) - (p > 0 ? 1.0 / (double)p : 0.0);
    d.cont = (double)d.pres * lr;
    return d;
}

// ---- the five scores (placement.hpp:120-149) ----
struct SampleScalars {
    double readMagnitude;       // sqrt(sum log1p(count)^2)
    double logContDenom;        // sum log1p(count)
    double wcDenom;             // sum over root seeds in reads of 1/childCount
    double uniqueKept;          // U' as double
    long long minSupport;
    long long uniqueSeeds;      // table entries after homopolymer removal / masking
    long long uniqueKeptInt;    // U'
    long long totalFrequency;   // sum of all counts (pre-filter)
    long long multiSum, multiCount;  // auto min-support statistics
    long long tableEntries;     // occupied slots before filters
    long long overflow;         // set when the table or a list ran out of room
};
PM_HD void nodeScores(double raw, double cosn, double pres, double wc, double cont, double gMag,
                      const SampleScalars& S, double* out) {
    // gMag = sqrt(genomeMagnitudeSquared), precomputed per node at index creation (sample independent)
    out[0] = S.readMagnitude > 0.0 ? raw / S.readMagnitude : 0.0;
    double v = 0.0;
    if (!(S.readMagnitude <= 0.0 || gMag <= 0.0)) {
        v = cosn / (S.readMagnitude * gMag);
        v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
    }
    out[1] = v;
    out[2] = S.uniqueKept > 0.0 ? pres / S.uniqueKept : 0.0;
    out[3] = S.wcDenom > 0.0 ? wc / S.wcDenom : 0.0;
    out[4] = S.logContDenom > 0.0 ? cont / S.logContDenom : 0.0;
}

// table slot hash (keys are already hashes, but their low bits come from XORs of rotations: mix once)
PM_HD u64 mixKey(u64 h) { h ^= h >> 32; h *= 0x9E3779B97F4A7C15ULL; h ^= h >> 29; return h; }

// sharded samples: owner rank of a seed = the high bits of its mixed hash scaled to [0, nRanks) (table slots use the low bits)
PM_HD u32 seedOwner(u64 h, u32 nRanks) { return (u32)(((mixKey(h) >> 32) * (u64)nRanks) >> 32); }

}  // namespace pm
