// hostcheck.cpp -- TEST INFRASTRUCTURE.  Runs the PM_HD device logic of panmap_b200/csrc/pm_logic.cuh and the host
// flattener on the CPU, emulating what the kernels do tile by tile, so the algorithms (rolling seeder, exact
// fixed-point tree prefix with closers / ancestor chains / carry slots, record-based selection) can be checked
// against the oracle without a GPU.  Never linked into libpanmap_b200.so.
#include "../../panmap_b200/csrc/pm_host.h"
#include "../../panmap_b200/csrc/pm_logic.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

using namespace pm;
static std::string g_err;

extern "C" {
const char* hc_last_error() { return g_err.c_str(); }

// mode 1: syncmers (hash,rev,pos) ; mode 2: seeds.  returns count
int64_t hc_seed(const char* seq, int64_t len, int k, int s, int t, int l, int open, int trimS, int trimE, int mode,
                uint64_t* outHash, uint8_t* outRev, int64_t* outPos, int64_t cap) {
    SeedTables T; buildSeedTables(T, k, s);
    const SeederParams P = makeSeederParams(k, s, t, mode == 1 ? 1 : l, open, trimS, trimE);
    std::vector<u64> ring((size_t)seederRingWords(k, s, l) * 3, 0);
    ReadSeederT<3> sd; sd.reset(ring.data() + 1, k - s + 1);  // stride 3 to exercise the strided view
    int64_t n = 0;
    if (len < k) return 0;
    for (int i = 0; i < (int)len; ++i) {
        u64 h; bool rev;
        if (sd.pushBase(i, baseCode((unsigned char)seq[i]), T, P, h, rev)) {
            const int pos = i - k + 1;
            if (mode == 1) { if (n < cap) { outHash[n] = h; outRev[n] = rev; outPos[n] = pos; } ++n; }
            else { u64 seed; if (sd.pushSyncmer(pos, (int)len, h, P, seed)) { if (n < cap) outHash[n] = seed; ++n; } }
        }
    }
    return n;
}

}  // extern "C"
// the syncmers_rank kernel's decision logic on the host (pm_logic.cuh: rank table, 2-bit histories, 16x2 sliding minimum), with the
// k-mer hashes taken from their definition: (hash, pos) of every closed syncmer, t = 0, s = 8
template <int K>
static int64_t seedRankT(const char* seq, int64_t len, int trimS, int trimE, const uint16_t* rank, uint64_t* outHash, int64_t* outPos, int64_t cap) {
    constexpr int S = kRankS, W = K - S + 1;
    if (len < K) return 0;
    RankWindow<W> win; win.reset();
    unsigned HF = 0, HR = 0, cur = 0;
    int iLo = trimS + K - 1; const int iHi = (int)len - trimE - 1;
    int64_t n = 0;
    auto codeAt = [&](int64_t i) -> unsigned { return i < len ? baseCode((unsigned char)seq[i]) : 4u; };
    for (int i = 0; i < (int)len; ++i) {
        const int ph = i & 3;
        if (ph == 0) {
            cur = codeAt(i) | (codeAt(i + 1) << 8) | (codeAt(i + 2) << 16) | (codeAt(i + 3) << 24);
            rankPushWord(cur, HF, HR);
        }
        if ((cur & 0x04040404u) && (cur & (4u << (8 * ph)))) iLo = std::max(iLo, i + K);   // an ambiguous base closes every window that holds it
        if (i < S - 1) continue;
        const unsigned rf = rank[rankAddrF(HF, ph) >> 1], rr = rank[rankAddrR(HR, ph) >> 1];
        const int j = (i - (S - 1)) % W;
        bool syn = false;
        // step() wants a constant slot after unrolling on the device; here a switch stands in for the unrolled loop
        switch (j) {
#define PM_CASE(J) case J: if (J < W) syn = win.step(J < W ? J : 0, rf | (rr << 16)); break;
            PM_CASE(0) PM_CASE(1) PM_CASE(2) PM_CASE(3) PM_CASE(4) PM_CASE(5) PM_CASE(6) PM_CASE(7) PM_CASE(8) PM_CASE(9) PM_CASE(10) PM_CASE(11)
#undef PM_CASE
        }
        if (!syn || i < iLo || i > iHi) continue;
        u64 fk = 0, rk = 0;
        for (int q = 0; q < K; ++q) {
            const unsigned c = baseCode((unsigned char)seq[i - K + 1 + q]);
            fk ^= rol64(codeHash(c), (unsigned)(K - 1 - q));
            rk ^= rol64(c < 4 ? codeHash(3 - c) : 0ULL, (unsigned)q);
        }
        if (fk == rk) continue;
        if (n < cap) { outHash[n] = umin64(fk, rk); outPos[n] = i - K + 1; }
        ++n;
    }
    return n;
}
extern "C" {
int64_t hc_seed_rank(const char* seq, int64_t len, int k, int trimS, int trimE, uint64_t* outHash, int64_t* outPos, int64_t cap) {
    static std::vector<uint16_t> rank;
    if (rank.empty()) { rank.resize(kRankEntries); buildSmerRanks(rank.data()); }
    if (k == 19) return seedRankT<19>(seq, len, trimS, trimE, rank.data(), outHash, outPos, cap);
    if (k == 15) return seedRankT<15>(seq, len, trimS, trimE, rank.data(), outHash, outPos, cap);
    return -1;
}
// number of distinct ranks (65,536 unless two 8-mers share a hash)
uint32_t hc_rank_distinct() { std::vector<uint16_t> r(kRankEntries); buildSmerRanks(r.data()); uint32_t mx = 0; for (auto v : r) mx = std::max<uint32_t>(mx, v); return mx + 1; }

// owner rank of every seed of a sharded sample (pm_logic.cuh seedOwner, the function partition_export uses on the device)
void hc_seed_owner(const uint64_t* h, int64_t n, uint32_t nRanks, uint32_t* out) { for (int64_t i = 0; i < n; ++i) out[i] = seedOwner(h[i], nRanks); }
// exact magnitude sums the way gathered_finalize + finish_scalars form them (minus the rounding-drift term): per-entry fixed-point terms
// of the listed counts plus n1 times the count-1 term; out: sum log1p^2, sum log1p
void hc_gathered_sums(const uint32_t* cnt, int64_t n, uint64_t n1, double* out) {
    fx128 m = fxZero(), l = fxZero();
    for (int64_t i = 0; i < n; ++i) { const double x = std::log1p((double)cnt[i]); m = fxAdd(m, fxFromDouble(x * x)); l = fxAdd(l, fxFromDouble(x)); }
    const double l1 = std::log1p(1.0);
    m = fxAdd(m, fxMulU64(fxFromDouble(l1 * l1), n1)); l = fxAdd(l, fxMulU64(fxFromDouble(l1), n1));
    out[0] = fxToDouble(m); out[1] = fxToDouble(l);
}

// the device's read-magnitude scalars on the host: exact fixed-point sums of log1p(count) and its square over the kept counts plus the
// expected sequential-summation drift from the count histogram (finish_scalars); out: sum log1p^2, sum log1p -- both WITH the drift term
void hc_magnitude_model(const uint32_t* cnt, int64_t n, double* out) {
    fx128 m = fxZero(), l = fxZero();
    std::vector<uint32_t> sorted(cnt, cnt + n);
    std::sort(sorted.begin(), sorted.end());
    for (int64_t i = 0; i < n; ++i) { const double x = std::log1p((double)cnt[i]); m = fxAdd(m, fxFromDouble(x * x)); l = fxAdd(l, fxFromDouble(x)); }
    const double M = fxToDouble(m), L = fxToDouble(l);
    double dM = 0.0, dL = 0.0;
    if (M > 0.0 && L > 0.0) {
        Binades BM, BL; makeBinades(M, BM); makeBinades(L, BL);
        for (int64_t i = 0; i < n;) {
            int64_t j = i; while (j < n && sorted[j] == sorted[i]) ++j;
            if (sorted[i] < 65536u) { const double x = std::log1p((double)sorted[i]); dM += driftOf(BM, x * x) * (double)(j - i); dL += driftOf(BL, x) * (double)(j - i); }
            i = j;
        }
    }
    out[0] = M + dM; out[1] = L + dL;
}

double hc_fx_roundtrip(double x) { return fxToDouble(fxFromDouble(x)); }
// exact sum of doubles through the fixed-point accumulator, in the given order and in reverse: both must agree
int hc_fx_sum(const double* x, int64_t n, double* fwd, double* rev) {
    fx128 a = fxZero(), b = fxZero();
    for (int64_t i = 0; i < n; ++i) a = fxAdd(a, fxFromDouble(x[i]));
    for (int64_t i = n; i-- > 0;) b = fxAdd(b, fxFromDouble(x[i]));
    *fwd = fxToDouble(a); *rev = fxToDouble(b);
    return (a.lo == b.lo && a.hi == b.hi) ? 1 : 0;
}

static Acc5 accZ() { Acc5 a; for (auto& x : a.f) x = fxZero(); a.pres = 0; return a; }
static Acc5 accAdd(const Acc5& a, const Acc5& b) { Acc5 r; for (int i = 0; i < 4; ++i) r.f[i] = fxAdd(a.f[i], b.f[i]); r.pres = a.pres + b.pres; return r; }
static Acc5 accNeg(const Acc5& a) { Acc5 r; for (int i = 0; i < 4; ++i) r.f[i] = fxNeg(a.f[i]); r.pres = -a.pres; return r; }
struct Seg { u64 lo; int hi; int cnt; };
static Seg segZ() { return Seg{0, 0, 0}; }
static Seg segAdd(const Seg& a, const Seg& b) { Seg r; r.lo = a.lo + b.lo; r.hi = a.hi + b.hi + (r.lo < a.lo ? 1 : 0); r.cnt = a.cnt + b.cnt; return r; }
static Seg segOf(long long v, int c) { return Seg{(u64)v, (int)(v >> 63), c}; }

// CPU emulation of finalize + K1 (node_deltas, warp by warp) + general deltas + K2 (prefix_scores, tile by tile) for shard
// `shard` of `nShards`.  table: (hash, logv = log1p(count) or 0).  metrics out: [N][5] (only shard nodes are written),
// scores out: [N][5].  scal: U', magnitude, logSum, wcDenom.
int hc_emulate_scoring(const pm_index_desc* d, uint32_t shard, uint32_t nShards, const uint64_t* tHash, const double* tLog, int64_t U,
                       double U1, double mag, double denL, double* metrics, double* scores, double* wcDenOut,
                       uint32_t* nodeBegin, uint32_t* nodeEnd, uint32_t chunksPerWarp) {
    try {
        FlatIndex F; flattenIndex(*d, shard, nShards, F);
        *nodeBegin = F.nodeBegin; *nodeEnd = F.nodeEnd;
        // ell by seed id through the dictionary table (as table_finalize does): log1p(count) * 2^53, an exact integer
        std::vector<long long> ell(F.S + 2, 0);
        for (int64_t i = 0; i < U; ++i) {
            if (!(tLog[i] > 0.0)) continue;
            u64 s = mixKey(tHash[i]) & F.dictMask;
            while (true) {
                if (F.dictKeys[s] == tHash[i]) {
                    if (F.dictVals[s] != 0xFFFFFFFFu) {
                        const double sc = tLog[i] * 9007199254740992.0;
                        if (sc != std::floor(sc) || sc >= 9.2e18) throw std::runtime_error("log1p(count) is not a multiple of 2^-53");
                        ell[F.dictVals[s]] = (long long)sc;
                    }
                    break;
                }
                if (F.dictVals[s] == 0xFFFFFFFFu && F.dictKeys[s] == kEmptyKey) break;
                s = (s + 1) & F.dictMask;
            }
        }
        // root denominator
        fx128 wc = fxZero();
        for (size_t i = 0; i < F.rootId.size(); ++i) {
            const int c = (int)F.rootChild[i];
            if (c > 0 && ell[F.rootId[i]] != 0) wc = fxAdd(wc, fxFromDouble(1.0 / (double)c));
        }
        const double denW = fxToDouble(wc);
        *wcDenOut = denW;
        // ---- K1: 32 lanes x 16 words per chunk ----
        if (F.dw.size() != F.nDeltaChunks * 512 || F.chunkSeg.size() != F.nDeltaChunks + 1 || F.endMask.size() != F.nDeltaChunks * 32)
            throw std::runtime_error("chunk schedule size mismatch");
        std::vector<Seg> segRec(F.nSeg + 1, segZ());
        std::vector<int> stored(F.nSeg + 1, 0);
        std::vector<char> isBoundary(F.nSeg + 1, 0);
        for (uint32_t b : F.boundarySegs) isBoundary[b] = 1;
        std::vector<char> atomicked(F.nSeg + 1, 0);
        auto store = [&](uint32_t sg, const Seg& v) {
            if (sg >= F.nSeg) throw std::runtime_error("segment index out of range");
            if (atomicked[sg]) throw std::runtime_error("plain store to a segment that also receives atomic adds");
            if (stored[sg]++) throw std::runtime_error("segment stored twice");
            segRec[sg] = v;
        };
        auto atomic = [&](uint32_t sg, const Seg& v) {
            if (sg >= F.nSeg) throw std::runtime_error("atomic segment index out of range");
            if (!isBoundary[sg]) throw std::runtime_error("atomic add to a segment that is not zeroed per sample");
            if (stored[sg]) throw std::runtime_error("atomic add to a segment that was plainly stored");
            atomicked[sg] = 1;
            segRec[sg] = segAdd(segRec[sg], v);
        };
        const u64 per = chunksPerWarp ? chunksPerWarp : 1;
        for (u64 c0 = 0; c0 < F.nDeltaChunks; c0 += per) {   // one warp = a run of `per` consecutive chunks
          const u64 c1 = std::min<u64>(c0 + per, F.nDeltaChunks);
          uint32_t segBase = F.chunkSeg[c0] & 0x7FFFFFFFu;
          bool outside = (F.chunkSeg[c0] >> 31) != 0;
          Seg carry = segZ(); bool open = false;
          for (u64 c = c0; c < c1; ++c) {
            if (segBase != (F.chunkSeg[c] & 0x7FFFFFFFu)) throw std::runtime_error("running segment base disagrees with chunkSeg");
            unsigned nEnd[32], endsBefore[32]; Seg tail[32], head[32]; unsigned endMask = 0, total = 0;
            for (int lane = 0; lane < 32; ++lane) {
                uint32_t w[16];   // lane-interleaved storage: 16-byte piece q of lane l at uint4 index 32 q + l
                for (int q = 0; q < 4; ++q) for (int r = 0; r < 4; ++r) w[4 * q + r] = F.dw[c * 512 + 4 * (32 * q + lane) + r];
                const unsigned Fm = F.endMask[c * 32 + lane];
                if (Fm >> 16) throw std::runtime_error("end mask has bits above 15");
                nEnd[lane] = (unsigned)__builtin_popcount(Fm);
                endsBefore[lane] = total; total += nEnd[lane];
                if (nEnd[lane]) endMask |= 1u << lane;
                const uint32_t segFirst = segBase + endsBefore[lane];
                long long P[16]; int Cn[16]; long long acc = 0; int cn = 0;
                for (int j = 0; j < 16; ++j) {   // running sums, staged in shared memory by the kernel
                    const long long e = ell[w[j] >> 1], mm = -(long long)(w[j] & 1u);
                    const long long v = (e ^ mm) - mm;
                    acc += v; cn += (v > 0) - (v < 0);
                    P[j] = acc; Cn[j] = cn;
                }
                head[lane] = segZ();
                if (Fm) {
                    const int first = __builtin_ctz(Fm), last = 31 - __builtin_clz(Fm);
                    head[lane] = segOf(P[first], Cn[first]);
                    long long pv = P[first]; int pcn = Cn[first]; unsigned k = 1;
                    for (unsigned m = Fm & (Fm - 1); m; m &= m - 1, ++k) {
                        const int j = __builtin_ctz(m);
                        store(segFirst + k, segOf(P[j] - pv, Cn[j] - pcn));
                        pv = P[j]; pcn = Cn[j];
                    }
                    acc -= P[last]; cn -= Cn[last];
                }
                tail[lane] = segOf(acc, cn);
            }
            Seg incl[32];
            for (int lane = 0; lane < 32; ++lane) incl[lane] = tail[lane];
            if (!nEnd[0]) incl[0] = segAdd(incl[0], carry);
            for (int dd = 1; dd < 32; dd <<= 1) {   // Hillis-Steele, all lanes read the previous step's values
                Seg prev[32]; for (int lane = 0; lane < 32; ++lane) prev[lane] = incl[lane];
                for (int lane = dd; lane < 32; ++lane) {
                    const unsigned span = ((1u << dd) - 1u) << (lane - dd + 1);
                    if (!(endMask & span)) incl[lane] = segAdd(incl[lane], prev[lane - dd]);
                }
            }
            for (int lane = 0; lane < 32; ++lane) {
                if (!nEnd[lane]) continue;
                const Seg f = segAdd(head[lane], lane ? incl[lane - 1] : carry);
                const uint32_t sg = segBase + endsBefore[lane];
                if (outside && !(endMask & ((1u << lane) - 1u))) atomic(sg, f); else store(sg, f);
            }
            if (endMask) outside = false;
            open = !((F.endMask[c * 32 + 31] >> 15) & 1u);
            carry = open ? incl[31] : segZ();
            segBase += total;
          }
          if (!open) continue;   // (the kernel issues a no-op: its atomic add skips zero values)
          if (segBase < F.nSeg) atomic(segBase, carry);
          else if (carry.lo || carry.hi || carry.cnt) throw std::runtime_error("open run after the last segment");
        }
        for (uint32_t sg = 0; sg < F.nSeg; ++sg)
            if (!atomicked[sg] && stored[sg] != 1) throw std::runtime_error("a segment was not stored exactly once");
        // ---- general deltas + event prefix ----
        std::vector<double> l1p(32768); for (int c = 0; c < 32768; ++c) l1p[c] = std::log1p((double)c);
        const double ln2 = std::log1p(1.0);
        std::vector<Acc5> genRec(F.nGenNodes, accZ()), evPrefix(F.evSlot.size(), accZ());
        for (size_t i = 0; i < F.genSlot.size(); ++i) {
            const long long e = ell[F.genId[i]];
            if (!e) continue;
            const int p = (int)(short)(F.genPc[i] & 0xFFFF), cc = (int)(short)(F.genPc[i] >> 16);
            const DeltaTerms t = deltaTerms((double)e / 9007199254740992.0, p, cc, p > 0 ? l1p[p] : 0.0, cc > 0 ? l1p[cc] : 0.0);
            Acc5 g; g.f[0] = fxFromDouble(t.raw); g.f[1] = fxFromDouble(t.cos); g.f[2] = fxFromDouble(t.wc); g.f[3] = fxFromDouble(t.cont); g.pres = t.pres;
            genRec[F.genSlot[i]] = accAdd(genRec[F.genSlot[i]], g);
        }
        {
            Acc5 run = accZ();
            for (size_t e = 0; e < F.evSlot.size(); ++e) {
                const Acc5 v = genRec[F.evSlot[e] & 0x7FFFFFFFu];
                run = accAdd(run, (F.evSlot[e] >> 31) ? accNeg(v) : v);
                evPrefix[e] = run;
            }
        }
        // ---- K2 tile algorithm ----
        auto recOf = [&](uint32_t v) { return F.nodeSeg[v] == 0xFFFFFFFFu ? segZ() : segRec[F.nodeSeg[v]]; };
        SampleScalars S; std::memset(&S, 0, sizeof(S));
        S.readMagnitude = mag; S.logContDenom = denL; S.wcDenom = denW; S.uniqueKept = U1;
        for (uint32_t tile = 0; tile < F.nK2Tiles; ++tile) {
            const uint32_t T = kTileNodesK2;
            const uint32_t a0 = F.nodeBegin + tile * T, a1 = std::min(a0 + T, F.nodeEnd);
            const uint32_t cb = F.chainOff[tile], ce = F.chainOff[tile + 1];
            std::vector<Seg> chainA(ce - cb);
            Seg run = segZ();
            for (uint32_t j = cb; j < ce; ++j) { run = segAdd(run, recOf(F.chainNodes[j])); chainA[j - cb] = run; }
            // d' split into carry-free limbs; every node subtracts itself where its subtree ends (shared-memory reductions)
            std::vector<long long> sA(a1 - a0), sB(a1 - a0); std::vector<int> sC(a1 - a0);
            std::vector<Seg> dp(a1 - a0);
            for (uint32_t w = a0; w < a1; ++w) {
                Seg v = recOf(w);
                const uint32_t cs = F.carrySlot[w];
                if (cs != 0xFFFFFFFFu) { if (cs >= ce - cb) throw std::runtime_error("carry slot outside chain"); v = segAdd(v, chainA[cs]); }
                dp[w - a0] = v;
                sA[w - a0] = (long long)(v.lo & 0xFFFFFFFFULL); sB[w - a0] = (long long)((v.lo >> 32) | ((u64)(i64)v.hi << 32)); sC[w - a0] = v.cnt;
            }
            for (uint32_t w = a1; w-- > a0;) {   // any order
                if (F.subEnd[w] >= a1) continue;
                const Seg& v = dp[w - a0]; const uint32_t li = F.subEnd[w] - a0;
                sA[li] -= (long long)(v.lo & 0xFFFFFFFFULL); sB[li] -= (long long)((v.lo >> 32) | ((u64)(i64)v.hi << 32)); sC[li] -= v.cnt;
            }
            Seg pre = segZ();
            for (uint32_t w = a0; w < a1; ++w) {
                const long long a = sA[w - a0], b = sB[w - a0];
                const Seg x{(u64)a, (int)(a >> 63), sC[w - a0]}, y{(u64)b << 32, (int)(b >> 32), 0};
                pre = segAdd(pre, segAdd(x, y));
                double* m = metrics + (size_t)w * 5;
                const uint32_t ne = F.nGenNodes ? F.evIdx[w] : 0u;
                nodeNumerators(pre.lo, pre.hi, pre.cnt, ne ? &evPrefix[ne - 1] : nullptr, ln2, m);
                nodeScores(m[0], m[1], m[2], m[3], m[4], F.gMag[w], S, scores + (size_t)w * 5);
            }
        }
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// CPU emulation of the record-based selection for one metric over nodes [nodeBegin,nodeEnd) of possibly several
// shards merged: input = scores of all N nodes + eligibility + BFS ranks; emulates bfs_gather/bfs_records (per
// 1024-block prefix maxima), the chain over unordered records and the ties pass.
int64_t hc_emulate_selection(const double* score, const uint8_t* eligible, const uint32_t* bfsRank, uint64_t N, uint32_t nShards,
                             double* bestScore, uint32_t* bestIdx, uint32_t* tied, int64_t tiedCap) {
    struct Rec { uint32_t rank, node; double s; };
    std::vector<Rec> recs;
    std::vector<std::vector<uint32_t>> shardNodes(nShards);
    for (uint64_t v = 0; v < N; ++v) shardNodes[(size_t)(v * nShards / N)].push_back((uint32_t)v);
    for (auto& nodes : shardNodes) {
        std::sort(nodes.begin(), nodes.end(), [&](uint32_t a, uint32_t b) { return bfsRank[a] < bfsRank[b]; });
        double run = 0.0;
        for (uint32_t v : nodes) {
            const double x = eligible[v] ? score[v] : -1.0;
            if (x > run) recs.push_back(Rec{bfsRank[v], v, x});
            run = std::max(run, x);
        }
    }
    std::reverse(recs.begin(), recs.end());  // the chain must not depend on the order of the record list
    double best = 0.0; uint32_t bn = 0xFFFFFFFFu; long long last = -1;
    while (true) {
        const double thr = best + std::fmax(best * 0.0001, 1e-9);
        long long pick = -1;
        for (size_t i = 0; i < recs.size(); ++i)
            if ((long long)recs[i].rank > last && recs[i].s > thr && (pick < 0 || recs[i].rank < recs[pick].rank)) pick = (long long)i;
        if (pick < 0) break;
        best = recs[pick].s; bn = recs[pick].node; last = recs[pick].rank;
    }
    std::vector<uint32_t> t;
    const double lo = best - std::fmax(best * 0.0001, 1e-9);
    for (uint64_t v = 0; v < N; ++v) {
        if (!eligible[v]) continue;
        if (last >= 0 && (long long)bfsRank[v] <= last) continue;
        if (score[v] >= lo && score[v] > 0.0) t.push_back((uint32_t)v);
    }
    if (bn != 0xFFFFFFFFu || !t.empty()) t.push_back(bn);
    std::sort(t.begin(), t.end()); t.erase(std::unique(t.begin(), t.end()), t.end());
    *bestScore = best; *bestIdx = t.empty() ? bn : t.front();
    for (size_t i = 0; i < t.size() && (int64_t)i < tiedCap; ++i) tied[i] = t[i];
    return (int64_t)t.size();
}
}  // extern "C"

// ---- cached index image (panmap_b200/csrc/pm_image.cpp) on the host: flatten -> write -> read back -> write again; the two files must be
// byte-identical (every field of FlatIndex survives), a wrong stamp / a flipped byte / a truncated file must be refused.
// returns 0 when everything holds, a positive step number otherwise
#include <cstdio>
static std::vector<unsigned char> slurp(const std::string& p) {
    std::vector<unsigned char> b; FILE* f = std::fopen(p.c_str(), "rb"); if (!f) return b;
    std::fseek(f, 0, SEEK_END); const long n = std::ftell(f); std::fseek(f, 0, SEEK_SET); b.resize((size_t)n);
    if (n && std::fread(b.data(), 1, (size_t)n, f) != (size_t)n) b.clear();
    std::fclose(f); return b;
}
extern "C" int hc_image_roundtrip(const pm_index_desc* d, uint32_t shard, uint32_t nShards, const char* const* ids, const char* dir) {
    try {
        FlatIndex F; flattenIndex(*d, shard, nShards, F);
        std::vector<std::string> nodeIds;
        if (ids) for (uint64_t i = 0; i < d->n_nodes; ++i) nodeIds.emplace_back(ids[i]);
        ImageStamp st; st.srcSize = 1234; st.srcMtimeNs = 99; st.srcHeader[3] = 7; st.shard = shard; st.nShards = nShards;
        const std::string a = std::string(dir) + "/a.pmflat", b = std::string(dir) + "/b.pmflat", c = std::string(dir) + "/c.pmflat";
        writeFlatImage(a, F, nodeIds, st);
        FlatIndex G; std::vector<std::string> idsBack; std::string why;
        if (!readFlatImage(a, G, idsBack, &st, &why)) { g_err = why; return 1; }
        if (idsBack != nodeIds) return 2;
        if (G.N != F.N || G.D != F.D || G.S != F.S || G.dw != F.dw || G.endMask != F.endMask || G.dictKeys != F.dictKeys || G.gMag != F.gMag ||
            G.chainNodes != F.chainNodes || G.bfsNodes != F.bfsNodes || G.nK2Tiles != F.nK2Tiles || std::memcmp(G.homo, F.homo, sizeof(F.homo)) != 0) return 3;
        writeFlatImage(b, G, idsBack, st);
        const auto fa = slurp(a), fb = slurp(b);
        if (fa.empty() || fa != fb) return 4;
        ImageStamp other = st; other.srcMtimeNs += 1;
        FlatIndex H; std::vector<std::string> x;
        if (readFlatImage(a, H, x, &other, &why)) return 5;                       // the source changed
        other = st; other.shard += 1;
        if (readFlatImage(a, H, x, &other, &why)) return 6;                       // another shard's image
        if (!readFlatImage(a, H, x, nullptr, &why)) return 7;                     // no expectation: accepted
        auto bad = fa; bad[bad.size() / 2] ^= 0x10;                                // one flipped bit in the middle
        { FILE* f = std::fopen(c.c_str(), "wb"); std::fwrite(bad.data(), 1, bad.size(), f); std::fclose(f); }
        if (readFlatImage(c, H, x, &st, &why)) return 8;
        { FILE* f = std::fopen(c.c_str(), "wb"); std::fwrite(fa.data(), 1, fa.size() - 4096 > 0 ? fa.size() / 2 : 1, f); std::fclose(f); }
        if (readFlatImage(c, H, x, &st, &why)) return 9;                          // truncated
        if (readFlatImage(std::string(dir) + "/missing.pmflat", H, x, &st, &why)) return 10;
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// ---- index builder: what genome_materialize (pm_build_kernels.cu) does with the flattened tree, step for step on the host ----
// template copy, block state replayed along the root -> node path, point edits level by level (root first), compaction in aligned
// order with inverted blocks mirrored and complemented.  Every node's genome, concatenated in pre-order; returns the node count or -1.
static char hcComplement(char c) {
    switch (c) {
        case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
        case 'R': return 'Y'; case 'Y': return 'R'; case 'K': return 'M'; case 'M': return 'K';
        case 'B': return 'V'; case 'V': return 'B'; case 'D': return 'H'; case 'H': return 'D';
        default: return c;
    }
}
extern "C" int64_t hc_flat_genomes(const char* panman_path, char* out, uint64_t cap, uint64_t* offsets /* [n + 1] */, uint64_t maxNodes) {
    try {
        PanmanTree T; readPanman(panman_path, T);
        PanmanFlat F; flattenPanman(T, F);
        const size_t N = std::min<size_t>(T.nodes.size(), (size_t)maxNodes), B = F.blockStart.size() - 1, A = F.tmpl.size();   // the first maxNodes nodes in pre-order
        uint64_t total = 0;
        std::vector<uint32_t> path;
        for (size_t v = 0; v < N; ++v) {
            path.clear();
            for (uint32_t u = (uint32_t)v;; u = F.parent[u]) { path.push_back(u); if (F.parent[u] == kNoNode) break; }
            if (path.size() > F.maxDepth) { g_err = "maxDepth too small"; return -1; }
            std::string al = F.tmpl;
            std::vector<unsigned char> st(B, 2);
            for (size_t lv = path.size(); lv-- > 0;) {
                const uint32_t u = path[lv];
                for (uint32_t i = F.blockMutBegin[u]; i < F.blockMutBegin[u + 1]; ++i) {
                    const uint32_t m = F.blockMut[i], b = m >> 2;
                    if (m & 1u) st[b] = (unsigned char)(1u | ((m & 2u) ? 0u : 2u));
                    else if (m & 2u) st[b] ^= 2u;
                    else st[b] = 2u;
                }
            }
            for (size_t lv = path.size(); lv-- > 0;) {
                const uint32_t u = path[lv];
                for (uint32_t i = F.editBegin[u]; i < F.editBegin[u + 1]; ++i) al[F.editSlot[i]] = F.editChar[i];
            }
            offsets[v] = total;
            for (size_t q = 0; q < A; ++q) {
                const uint32_t b = F.slotBlock[q];
                if (!(st[b] & 1u)) continue;
                char c;
                if (st[b] & 2u) c = al[q];
                else c = hcComplement(al[F.blockStart[b] + (F.blockStart[b + 1] - 1u - q)]);
                if (c == '-') continue;
                if (total < cap) out[total] = c;
                ++total;
            }
        }
        offsets[N] = total;
        if (total > cap) { g_err = "output buffer too small"; return -1; }
        return (int64_t)N;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
