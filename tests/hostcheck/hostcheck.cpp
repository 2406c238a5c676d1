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

double hc_fx_roundtrip(double x) { return fxToDouble(fxFromDouble(x)); }
// exact sum of doubles through the fixed-point accumulator, in the given order and in reverse: both must agree
int hc_fx_sum(const double* x, int64_t n, double* fwd, double* rev) {
    fx128 a = fxZero(), b = fxZero();
    for (int64_t i = 0; i < n; ++i) a = fxAdd(a, fxFromDouble(x[i]));
    for (int64_t i = n; i-- > 0;) b = fxAdd(b, fxFromDouble(x[i]));
    *fwd = fxToDouble(a); *rev = fxToDouble(b);
    return (a.lo == b.lo && a.hi == b.hi) ? 1 : 0;
}

struct Acc { fx128 f[4]; i64 pres; };
static Acc accZ() { Acc a; for (auto& x : a.f) x = fxZero(); a.pres = 0; return a; }
static Acc accAdd(const Acc& a, const Acc& b) { Acc r; for (int i = 0; i < 4; ++i) r.f[i] = fxAdd(a.f[i], b.f[i]); r.pres = a.pres + b.pres; return r; }
static Acc accSub(const Acc& a, const Acc& b) { Acc r; for (int i = 0; i < 4; ++i) r.f[i] = fxSub(a.f[i], b.f[i]); r.pres = a.pres - b.pres; return r; }

// CPU emulation of finalize + K1 + K2 for shard `shard` of `nShards`.
// table: (hash sorted ascending, logv = log1p(count) or 0).  metrics out: [N][5] (only shard nodes are written),
// scores out: [N][5].  scal: U', magnitude, logSum, wcDenom.
int hc_emulate_scoring(const pm_index_desc* d, uint32_t shard, uint32_t nShards, const uint64_t* tHash, const double* tLog, int64_t U,
                       double U1, double mag, double denL, double* metrics, double* scores, double* wcDenOut,
                       uint32_t* nodeBegin, uint32_t* nodeEnd) {
    try {
        FlatIndex F; flattenIndex(*d, shard, nShards, F);
        *nodeBegin = F.nodeBegin; *nodeEnd = F.nodeEnd;
        // ell by seed id through the dictionary table (as table_finalize does)
        std::vector<double> ell(F.S, 0.0);
        for (int64_t i = 0; i < U; ++i) {
            if (!(tLog[i] > 0.0)) continue;
            u64 s = mixKey(tHash[i]) & F.dictMask;
            while (true) {
                if (F.dictKeys[s] == tHash[i]) { if (F.dictVals[s] != 0xFFFFFFFFu) ell[F.dictVals[s]] = tLog[i]; break; }
                if (F.dictVals[s] == 0xFFFFFFFFu && F.dictKeys[s] == kEmptyKey) break;
                s = (s + 1) & F.dictMask;
            }
        }
        // root denominator
        fx128 wc = fxZero();
        for (uint32_t i = 0; i < F.rootDCount; ++i) {
            const int c = (int)(short)(F.pc[F.rootDBegin + i] >> 16);
            if (c > 0 && ell[F.seedId[F.rootDBegin + i]] > 0.0) wc = fxAdd(wc, fxFromDouble(1.0 / (double)c));
        }
        const double denW = fxToDouble(wc);
        *wcDenOut = denW;
        // K1 (node_deltas): warp-per-chunk emulation -- 32 lanes x 16 deltas, interior stores, segmented scan over the lanes'
        // trailing partials, boundary nodes through (here: plain) accumulation, repeats added afterwards
        struct Tot { fx128 raw, cos, wc, cont; long long pres; };
        auto tz = []() { Tot t; t.raw = t.cos = t.wc = t.cont = fxZero(); t.pres = 0; return t; };
        auto tadd = [](const Tot& x, const Tot& y) { Tot r; r.raw = fxAdd(x.raw, y.raw); r.cos = fxAdd(x.cos, y.cos); r.wc = fxAdd(x.wc, y.wc); r.cont = fxAdd(x.cont, y.cont); r.pres = x.pres + y.pres; return r; };
        std::vector<Tot> delta(F.N, tz());
        std::vector<int> written(F.nLocal, 0);
        std::vector<double> l1p(32768); for (int c = 0; c < 32768; ++c) l1p[c] = std::log1p((double)c);
        const double ln2 = std::log1p(1.0);
        const u64 dReal = F.nLocalDeltas;
        auto nodeOf = [&](uint32_t lo, uint32_t hi, u64 d) { while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (F.lOff[mid + 1] > d) hi = mid; else lo = mid + 1; } return lo; };
        auto emit = [&](uint32_t ln, const Tot& t) {
            if (F.isBoundary[ln]) delta[F.lNode[ln]] = tadd(delta[F.lNode[ln]], t);
            else { if (written[ln]++) throw std::runtime_error("non-boundary node emitted twice"); delta[F.lNode[ln]] = t; }
        };
        if (F.seedId.size() != F.nDeltaChunks * 512 || F.chunkNode.size() != F.nDeltaChunks + 1) throw std::runtime_error("chunk schedule size mismatch");
        for (u64 c = 0; c < F.nDeltaChunks; ++c) {
            bool has[32], single[32]; uint32_t nf[32], nl[32]; Tot acc[32], tFirst[32];
            std::vector<std::pair<u64, int>> general;
            const uint32_t n0 = F.chunkNode[c], n1 = F.chunkNode[c + 1];
            for (int lane = 0; lane < 32; ++lane) {
                const u64 d0 = c * 512 + (u64)lane * 16;
                has[lane] = d0 < dReal; acc[lane] = tz(); tFirst[lane] = tz(); single[lane] = true; nf[lane] = nl[lane] = 0;
                if (!has[lane]) continue;
                uint32_t node = nodeOf(n0, n1, d0);
                if (!(F.lOff[node] <= d0 && d0 < F.lOff[node + 1])) throw std::runtime_error("nodeOfDelta landed on the wrong node");
                nf[lane] = node;
                u64 nextOff = F.lOff[node + 1];
                bool firstDone = false;
                for (int j = 0; j < 16; ++j) {
                    const u64 idx = d0 + j;
                    if (idx >= dReal) continue;
                    if (idx >= nextOff) {
                        if (!firstDone) { tFirst[lane] = acc[lane]; firstDone = true; }
                        else { if (F.isBoundary[node]) throw std::runtime_error("interior node flagged boundary"); if (written[node]++) throw std::runtime_error("interior node written twice"); delta[F.lNode[node]] = acc[lane]; }
                        acc[lane] = tz();
                        do { ++node; nextOff = F.lOff[node + 1]; } while (idx >= nextOff);
                    }
                    const uint32_t pcv = F.pc[idx]; const int p = (int)(short)(pcv & 0xFFFF), cc = (int)(short)(pcv >> 16);
                    const double lr = ell[F.seedId[idx]];
                    if (p != cc && lr > 0.0) {
                        if ((unsigned)p <= 1u && (unsigned)cc <= 1u) {
                            const fx128 fl = fxFromDouble(lr), fc = fxFromDouble(lr * ln2);
                            Tot& a = acc[lane];
                            if (cc > p) { a.raw = fxAdd(a.raw, fl); a.cont = fxAdd(a.cont, fl); a.cos = fxAdd(a.cos, fc); a.wc.hi += 1; a.pres += 1; }
                            else { a.raw = fxSub(a.raw, fl); a.cont = fxSub(a.cont, fl); a.cos = fxSub(a.cos, fc); a.wc.hi -= 1; a.pres -= 1; }
                        } else general.push_back({idx, lane});
                    }
                }
                nl[lane] = node; single[lane] = !firstDone;
            }
            Tot incl[32]; int head[32]; bool contPrev[32];
            for (int lane = 0; lane < 32; ++lane) {
                contPrev[lane] = has[lane] && lane > 0 && has[lane - 1] && nl[lane - 1] == nf[lane];
                head[lane] = (single[lane] && contPrev[lane]) ? 0 : 1;
                incl[lane] = acc[lane];
            }
            for (int d = 1; d < 32; d <<= 1) {   // Hillis-Steele segmented scan, all lanes read the previous step's values
                Tot up[32]; int hup[32];
                for (int lane = 0; lane < 32; ++lane) { up[lane] = incl[lane >= d ? lane - d : lane]; hup[lane] = head[lane >= d ? lane - d : lane]; }
                for (int lane = 0; lane < 32; ++lane) if (lane >= d && !head[lane]) { incl[lane] = tadd(incl[lane], up[lane]); head[lane] = hup[lane]; }
            }
            for (int lane = 0; lane < 32; ++lane) {
                if (!has[lane]) continue;
                if (!single[lane]) emit(nf[lane], contPrev[lane] ? tadd(incl[lane - 1], tFirst[lane]) : tFirst[lane]);
                const bool hasNext = lane < 31 && has[lane + 1];
                if (!(hasNext && nf[lane + 1] == nl[lane])) emit(nl[lane], incl[lane]);
            }
            for (auto& g : general) {
                const u64 idx = g.first;
                const uint32_t pcv = F.pc[idx]; const int p = (int)(short)(pcv & 0xFFFF), cc = (int)(short)(pcv >> 16);
                const DeltaTerms t = deltaTerms(ell[F.seedId[idx]], p, cc, p > 0 ? l1p[p] : 0.0, cc > 0 ? l1p[cc] : 0.0);
                Tot gt; gt.raw = fxFromDouble(t.raw); gt.cos = fxFromDouble(t.cos); gt.wc = fxFromDouble(t.wc); gt.cont = fxFromDouble(t.cont); gt.pres = t.pres;
                const uint32_t ln = nodeOf(n0, n1, idx);
                delta[F.lNode[ln]] = tadd(delta[F.lNode[ln]], gt);
            }
        }
        for (uint32_t i = 0; i < F.nLocal; ++i)
            if (F.lOff[i + 1] > F.lOff[i] && !F.isBoundary[i] && written[i] != 1) throw std::runtime_error("a node with deltas was not written exactly once");
        // K2 tile algorithm
        auto fromND = [](const Tot& n) { Acc a; a.f[0] = n.raw; a.f[1] = n.cos; a.f[2] = n.wc; a.f[3] = n.cont; a.pres = n.pres; return a; };
        SampleScalars S; std::memset(&S, 0, sizeof(S));
        S.readMagnitude = mag; S.logContDenom = denL; S.wcDenom = denW; S.uniqueKept = U1;
        for (uint32_t tile = 0; tile < F.nK2Tiles; ++tile) {
            const uint32_t a0 = F.nodeBegin + tile * 512, a1 = std::min(a0 + 512u, F.nodeEnd);
            const uint32_t cb = F.chainOff[tile], ce = F.chainOff[tile + 1];
            std::vector<Acc> chainA(ce - cb);
            Acc run = accZ();
            for (uint32_t j = cb; j < ce; ++j) { run = accAdd(run, fromND(delta[F.chainNodes[j]])); chainA[j - cb] = run; }
            std::vector<Acc> dp(a1 - a0);
            for (uint32_t w = a0; w < a1; ++w) {
                Acc v = fromND(delta[w]);
                const uint32_t cs = F.carrySlot[w];
                if (cs != 0xFFFFFFFFu) { if (cs >= ce - cb) throw std::runtime_error("carry slot outside chain"); v = accAdd(v, chainA[cs]); }
                dp[w - a0] = v;
            }
            Acc pre = accZ();
            for (uint32_t w = a0; w < a1; ++w) {
                Acc dd = dp[w - a0];
                for (uint32_t c = F.closeOff[w]; c < F.closeOff[w + 1]; ++c) { const uint32_t u = F.closeList[c]; if (u >= a0) dd = accSub(dd, dp[u - a0]); }
                pre = accAdd(pre, dd);
                double* m = metrics + (size_t)w * 5;
                m[0] = fxToDouble(pre.f[0]); m[1] = fxToDouble(pre.f[1]); m[2] = (double)pre.pres; m[3] = fxToDouble(pre.f[2]); m[4] = fxToDouble(pre.f[3]);
                nodeScores(m[0], m[1], (double)(u64)pre.pres, m[3], m[4], F.gMag[w], S, scores + (size_t)w * 5);
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
