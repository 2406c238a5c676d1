/*
 * panmap_oracle.c -- CPU restatement of panmap's placement hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under panmap_b200/ may include, link or call this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and only
 * as the checker.  Parity pinning: tests/test_oracle_vs_reference.py checks every function here against
 * the reference's own translation units compiled unmodified into oracle/_ref/libpanmap_ref.so
 * (oracle/ref_build/Makefile), and tests/golden/ holds vectors generated from that library
 * (tools/make_golden.py), including the reference's one numeric golden for this path,
 * examples/expected/single_sample/isolate.placement.tsv.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference/).
 * Plain C11, no dependencies beyond libm.  Compile with -ffp-contract=off so that f64 sums associate
 * exactly as written (the reference's release build may contract a*a-b*b; see SURVEY.md hard part 3).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;

#define ORC_NONE 0xFFFFFFFFu

/* ---- seeding.hpp:86-120 : per-base constants, complement, rotates (rotate count taken mod 64) ---- */
u64 orc_chash(char c) {
    switch (c) {
        case 'a': case 'A': return 0x3c8bfbb395c60474ULL;
        case 'c': case 'C': return 0x3193c18562a02b4cULL;
        case 'g': case 'G': return 0x20323ed082572324ULL;
        case 't': case 'T': return 0x295549f54be24456ULL;
        default: return 0;
    }
}
char orc_comp(char c) {
    switch (c) {
        case 'A': return 'T'; case 'a': return 't'; case 'C': return 'G'; case 'c': return 'g';
        case 'G': return 'C'; case 'g': return 'c'; case 'T': return 'A'; case 't': return 'a';
        default: return 'N';
    }
}
u64 orc_rol(u64 h, u64 r) { r &= 63; return r ? (h << r) | (h >> (64 - r)) : h; }
u64 orc_ror(u64 h, u64 r) { r &= 63; return r ? (h >> r) | (h << (64 - r)) : h; }

/* ---- seeding.cpp:20-30 hashSeq : returns 0, or -1 where the reference throws (non-ACGT base) ---- */
int orc_hash_seq(const char* s, int k, u64* f, u64* r) {
    u64 fh = 0, rh = 0;
    for (int i = 0; i < k; i++) {
        if (orc_chash(s[i]) == 0) return -1;
        fh ^= orc_rol(orc_chash(s[i]), (u64)(k - i - 1));
        rh ^= orc_rol(orc_chash(orc_comp(s[k - i - 1])), (u64)(k - i - 1));
    }
    *f = fh; *r = rh;
    return 0;
}

/* ---- seeding.cpp:47-229 rollingSyncmers, restated per window in closed form ----
 * For k-mer start p, w = k-s+1 s-mers F[p..p+w), Rc[p..p+w) (forward hash of the s-mer / of its reverse
 * complement).  closed: fs = F[p+t]==min || F[p+k-s-t]==min ; rs = Rc[p+k-s-t]==min || Rc[p+t]==min
 * (the reference's reverse ring is indexed newest-first, seeding.cpp:190, hence the mirrored indices);
 * open: fs = F[p+t]==min ; rs = Rc[p+k-s-t]==min.  Windows containing a non-ACGT base are never syncmers
 * (seeding.cpp:115,196); f==r is dropped (seeding.cpp:138,219).  Output tuple = (hash,isReverse,isSyncmer,pos).
 * Returns the number of tuples (all windows if returnAll, else syncmers only); writes at most cap. */
i64 orc_rolling_syncmers(const char* seq, i64 len, int k, int s, int open, int t, int returnAll,
                         u64* hash, uint8_t* isRev, uint8_t* isSync, i64* pos, i64 cap) {
    if (len < k) return 0;
    const i64 nS = len - s + 1;
    u64* F = (u64*)malloc(sizeof(u64) * (size_t)nS);
    u64* Rc = (u64*)malloc(sizeof(u64) * (size_t)nS);
    for (i64 p = 0; p < nS; p++) {
        u64 f = 0, r = 0;
        for (int i = 0; i < s; i++) {
            f ^= orc_rol(orc_chash(seq[p + i]), (u64)(s - 1 - i));
            r ^= orc_rol(orc_chash(orc_comp(seq[p + s - 1 - i])), (u64)(s - 1 - i));
        }
        F[p] = f; Rc[p] = r;
    }
    const int w = k - s + 1;
    i64 n = 0;
    for (i64 p = 0; p + k <= len; p++) {
        u64 f = 0, r = 0; int amb = 0;
        for (int i = 0; i < k; i++) {
            const u64 c = orc_chash(seq[p + i]);
            if (c == 0) amb = 1;
            f ^= orc_rol(c, (u64)(k - 1 - i));
            r ^= orc_rol(orc_chash(orc_comp(seq[p + k - 1 - i])), (u64)(k - 1 - i));
        }
        u64 mf = UINT64_MAX, mr = UINT64_MAX;
        for (int j = 0; j < w; j++) { if (F[p + j] < mf) mf = F[p + j]; if (Rc[p + j] < mr) mr = Rc[p + j]; }
        int fs, rs;
        if (open) { fs = F[p + t] == mf; rs = Rc[p + k - s - t] == mr; }
        else {
            fs = (F[p + t] == mf) || (F[p + k - s - t] == mf);
            rs = (Rc[p + k - s - t] == mr) || (Rc[p + t] == mr);
        }
        const int syn = !amb && (fs || rs) && f != r;
        if (syn || returnAll) {
            if (n < cap) {
                hash[n] = syn ? (f < r ? f : r) : UINT64_MAX;
                isRev[n] = (uint8_t)(syn && r < f);
                isSync[n] = (uint8_t)syn;
                pos[n] = p;
            }
            n++;
        }
    }
    free(F); free(Rc);
    return n;
}

/* ---- placement.cpp:1598-1686 (l>=1) and :1335-1368 (l==0): seeds of ONE read -------------------------
 * l<=1: every syncmer whose start lies in [trimStart, len-trimEnd-k] (l==1 additionally requires at least
 * one syncmer, :1625, which is vacuous).  l>1: k-min-mers over the trim-restricted syncmer run, provided the
 * UNRESTRICTED list has >= l entries (:1625) and the restricted one too (:1648):
 *   Fw = XOR_w rol(h[j+w], k*(l-1-w)), Rw = XOR_w rol(h[j+w], k*w), seed = min(Fw,Rw) unless Fw==Rw.
 * Returns the number of seed instances; writes at most cap hashes. */
i64 orc_read_seeds(const char* seq, i64 len, int k, int s, int t, int l, int open, int trimStart, int trimEnd,
                   u64* out, i64 cap) {
    if (len < k) return 0;
    const i64 maxw = len - k + 1;
    u64* h = (u64*)malloc(sizeof(u64) * (size_t)maxw);
    uint8_t* b1 = (uint8_t*)malloc((size_t)maxw);
    uint8_t* b2 = (uint8_t*)malloc((size_t)maxw);
    i64* ps = (i64*)malloc(sizeof(i64) * (size_t)maxw);
    const i64 m = orc_rolling_syncmers(seq, len, k, s, open, t, 0, h, b1, b2, ps, maxw);
    const int validStart = trimStart;
    const int validEnd = (int)len - trimEnd - k;
    i64 n = 0;
    if (l <= 1) {
        for (i64 j = 0; j < m; j++) {
            const int sp = (int)ps[j];
            if (sp < validStart || sp > validEnd) continue;
            if (n < cap) out[n] = h[j];
            n++;
        }
    } else if (m >= l) {
        i64 lo = 0, hi = m;
        while (lo < hi && (int)ps[lo] < validStart) lo++;
        while (hi > lo && (int)ps[hi - 1] > validEnd) hi--;
        if (hi - lo >= l) {
            for (i64 j = lo; j + l <= hi; j++) {
                u64 fw = 0, rw = 0;
                for (int w = 0; w < l; w++) {
                    fw ^= orc_rol(h[j + w], (u64)k * (u64)(l - 1 - w));
                    rw ^= orc_rol(h[j + w], (u64)k * (u64)w);
                }
                if (fw != rw) { if (n < cap) out[n] = fw < rw ? fw : rw; n++; }
            }
        }
    }
    free(h); free(b1); free(b2); free(ps);
    return n;
}

/* ---- placement.cpp:41-76 : canonical hashes of the four homopolymer k-mers ---- */
void orc_homopolymer_hashes(int k, u64 out[4]) {
    const char b[4] = {'A', 'C', 'G', 'T'};
    for (int j = 0; j < 4; j++) {
        const u64 bv = orc_chash(b[j]), cv = orc_chash(orc_comp(b[j]));
        u64 f = 0, r = 0;
        for (int i = 0; i < k; i++) { f ^= orc_rol(bv, (u64)(k - i - 1)); r ^= orc_rol(cv, (u64)(k - i - 1)); }
        out[j] = f < r ? f : r;
    }
}

static int cmp_u64(const void* a, const void* b) {
    const u64 x = *(const u64*)a, y = *(const u64*)b;
    return x < y ? -1 : x > y;
}

/* ---- placement.cpp:79-88 avgPhredQuality: mean of (signed char) qual - 33 over the k-mer span; 0.0 when the span
 * does not fit the quality string ---- */
double orc_avg_phred(const char* qual, i64 qlen, i64 startPos, int k) {
    if (qlen == 0 || startPos < 0 || startPos + k > qlen) return 0.0;
    i64 sum = 0;
    for (int i = 0; i < k; i++) sum += (int)qual[startPos + i] - 33;
    return (double)sum / k;
}

/* ---- placement.cpp:1179-1240 (l==0) and :1388-1533 (l>=1) with --min-seed-quality > 0: seeds of ONE read ------
 * A syncmer "passes" when its start lies in [trimStart, len-trimEnd-k] and the average Phred quality over its k
 * bases is >= minQ.  l<=1: every passing syncmer is a seed.  l>1: k-min-mers over windows of l CONSECUTIVE syncmers of
 * the unrestricted list, kept only when all l of them pass (unlike the default path, which closes the list up after the
 * trim filter); needs >= l syncmers in the read (:1428).  Reads are never deduplicated on this path. */
i64 orc_read_seeds_quality(const char* seq, const char* qual, i64 len, int k, int s, int t, int l, int open, int trimStart,
                           int trimEnd, int minQ, u64* out, i64 cap) {
    if (len < k) return 0;
    const i64 maxw = len - k + 1;
    u64* h = (u64*)malloc(sizeof(u64) * (size_t)maxw);
    uint8_t* b1 = (uint8_t*)malloc((size_t)maxw);
    uint8_t* b2 = (uint8_t*)malloc((size_t)maxw);
    i64* ps = (i64*)malloc(sizeof(i64) * (size_t)maxw);
    uint8_t* pass = (uint8_t*)malloc((size_t)maxw);
    const i64 m = orc_rolling_syncmers(seq, len, k, s, open, t, 0, h, b1, b2, ps, maxw);
    const int validStart = trimStart;
    const int validEnd = (int)len - trimEnd - k;
    for (i64 j = 0; j < m; j++) {
        const int sp = (int)ps[j];
        pass[j] = (uint8_t)(sp >= validStart && sp <= validEnd && !(orc_avg_phred(qual, len, ps[j], k) < (double)minQ));
    }
    i64 n = 0;
    if (l <= 1) {
        for (i64 j = 0; j < m; j++) if (pass[j]) { if (n < cap) out[n] = h[j]; n++; }
    } else if (m >= l) {
        for (i64 j = 0; j + l <= m; j++) {
            int ok = 1;
            for (int w = 0; w < l; w++) ok &= pass[j + w];
            if (!ok) continue;
            u64 fw = 0, rw = 0;
            for (int w = 0; w < l; w++) {
                fw ^= orc_rol(h[j + w], (u64)k * (u64)(l - 1 - w));
                rw ^= orc_rol(h[j + w], (u64)k * (u64)w);
            }
            if (fw != rw) { if (n < cap) out[n] = fw < rw ? fw : rw; n++; }
        }
    }
    free(h); free(b1); free(b2); free(ps); free(pass);
    return n;
}


/* ---- placement.cpp:1550-1722 : reads -> seedFreqInReads (hash -> count), sorted by hash -----------------
 * reads = concatenated bytes, off[n_reads+1].  Counting every read occurrence equals the reference's
 * dedup-then-multiply (:1550-1620); with dedup!=0 each distinct sequence counts once (:1619).
 * The four homopolymer k-mer hashes are erased afterwards (:1708-1718).  Output arrays are malloc'd
 * (free with orc_free); returns the number of unique seeds U. */
static int cmp_read(const void* a, const void* b, void* ctx);
struct read_ctx { const char* reads; const u64* off; };
static int cmp_read(const void* a, const void* b, void* ctx) {
    const struct read_ctx* c = (const struct read_ctx*)ctx;
    const u64 i = *(const u64*)a, j = *(const u64*)b;
    const u64 li = c->off[i + 1] - c->off[i], lj = c->off[j + 1] - c->off[j];
    const int r = memcmp(c->reads + c->off[i], c->reads + c->off[j], li < lj ? li : lj);
    if (r) return r;
    return li < lj ? -1 : li > lj;
}
void orc_free(void* p) { free(p); }

/* seed instances -> (hash, count) sorted by hash, minus the four homopolymer k-mer hashes (:1708-1718); frees inst */
static i64 table_from_instances(u64* inst, u64 nInst, int k, u64** outHash, i64** outCount) {
    qsort(inst, (size_t)nInst, sizeof(u64), cmp_u64);
    u64 homo[4];
    orc_homopolymer_hashes(k, homo);
    u64* H = (u64*)malloc(sizeof(u64) * (size_t)(nInst ? nInst : 1));
    i64* C = (i64*)malloc(sizeof(i64) * (size_t)(nInst ? nInst : 1));
    i64 U = 0;
    for (u64 i = 0; i < nInst;) {
        u64 j = i;
        while (j < nInst && inst[j] == inst[i]) j++;
        const u64 hsh = inst[i];
        if (hsh != homo[0] && hsh != homo[1] && hsh != homo[2] && hsh != homo[3]) { H[U] = hsh; C[U] = (i64)(j - i); U++; }
        i = j;
    }
    free(inst);
    *outHash = H; *outCount = C;
    return U;
}

i64 orc_seed_table(const char* reads, const u64* off, u64 n_reads, int k, int s, int t, int l, int open,
                   int trimStart, int trimEnd, int dedup, u64** outHash, i64** outCount) {
    u64 capInst = 0;
    for (u64 i = 0; i < n_reads; i++) { const u64 L = off[i + 1] - off[i]; if (L >= (u64)k) capInst += L - (u64)k + 1; }
    u64* inst = (u64*)malloc(sizeof(u64) * (size_t)(capInst ? capInst : 1));
    u64 nInst = 0;
    if (!dedup) {
        for (u64 i = 0; i < n_reads; i++)
            nInst += (u64)orc_read_seeds(reads + off[i], (i64)(off[i + 1] - off[i]), k, s, t, l, open, trimStart, trimEnd,
                                         inst + nInst, (i64)(capInst - nInst));
    } else {
        u64* order = (u64*)malloc(sizeof(u64) * (size_t)(n_reads ? n_reads : 1));
        for (u64 i = 0; i < n_reads; i++) order[i] = i;
        struct read_ctx c = {reads, off};
        qsort_r(order, (size_t)n_reads, sizeof(u64), cmp_read, &c);
        for (u64 a = 0; a < n_reads; a++) {
            if (a > 0 && cmp_read(&order[a - 1], &order[a], &c) == 0) continue;
            const u64 i = order[a];
            nInst += (u64)orc_read_seeds(reads + off[i], (i64)(off[i + 1] - off[i]), k, s, t, l, open, trimStart, trimEnd,
                                         inst + nInst, (i64)(capInst - nInst));
        }
        free(order);
    }
    return table_from_instances(inst, nInst, k, outHash, outCount);
}

/* same with --min-seed-quality > 0 (quals: one byte per base, same offsets as the reads) */
i64 orc_seed_table_quality(const char* reads, const char* quals, const u64* off, u64 n_reads, int k, int s, int t, int l, int open,
                           int trimStart, int trimEnd, int minQ, u64** outHash, i64** outCount) {
    u64 capInst = 0;
    for (u64 i = 0; i < n_reads; i++) { const u64 L = off[i + 1] - off[i]; if (L >= (u64)k) capInst += L - (u64)k + 1; }
    u64* inst = (u64*)malloc(sizeof(u64) * (size_t)(capInst ? capInst : 1));
    u64 nInst = 0;
    for (u64 i = 0; i < n_reads; i++)
        nInst += (u64)orc_read_seeds_quality(reads + off[i], quals + off[i], (i64)(off[i + 1] - off[i]), k, s, t, l, open, trimStart, trimEnd,
                                             minQ, inst + nInst, (i64)(capInst - nInst));
    return table_from_instances(inst, nInst, k, outHash, outCount);
}

/* ---- placement.cpp:1748-1799 : mask the top floor(frac*U) seeds by count --------------------------------
 * The reference sorts (hash,count) pairs by count only with std::sort, so which seeds fall at a tied cut is
 * unspecified there; this restatement (and the CUDA path) break ties by ascending hash.  In place; returns new U. */
struct hc { u64 h; i64 c; };
static int cmp_hc_count_desc(const void* a, const void* b) {
    const struct hc* x = (const struct hc*)a; const struct hc* y = (const struct hc*)b;
    if (x->c != y->c) return x->c > y->c ? -1 : 1;
    return x->h < y->h ? -1 : x->h > y->h;
}
static int cmp_hc_hash(const void* a, const void* b) {
    const struct hc* x = (const struct hc*)a; const struct hc* y = (const struct hc*)b;
    return x->h < y->h ? -1 : x->h > y->h;
}
i64 orc_mask_top_seeds(u64* hash, i64* count, i64 U, double frac) {
    if (!(frac > 0.0) || U == 0) return U;
    const i64 numToMask = (i64)(size_t)(frac * (double)U);
    if (numToMask <= 0) return U;
    struct hc* v = (struct hc*)malloc(sizeof(struct hc) * (size_t)U);
    for (i64 i = 0; i < U; i++) { v[i].h = hash[i]; v[i].c = count[i]; }
    qsort(v, (size_t)U, sizeof(struct hc), cmp_hc_count_desc);
    const i64 drop = numToMask < U ? numToMask : U;
    qsort(v + drop, (size_t)(U - drop), sizeof(struct hc), cmp_hc_hash);
    for (i64 i = drop; i < U; i++) { hash[i - drop] = v[i].h; count[i - drop] = v[i].c; }
    free(v);
    return U - drop;
}

/* ---- placement.cpp:931-955 resolveMinReadSupport ---- */
i64 orc_resolve_min_read_support(const i64* count, i64 U, int configured) {
    if (configured >= 0) return configured;
    u64 sum = 0, n = 0;
    for (i64 i = 0; i < U; i++) if (count[i] >= 2) { sum += (u64)count[i]; n++; }
    const double est = n > 0 ? (double)sum / (double)n : 0.0;
    return est > 3.0 ? 2 : 1;
}

/* ---- placement.cpp:957-984 computeReadSeedMagnitudes -------------------------------------------------
 * logv[i] = log1p(count) for kept seeds, 0 for dropped ones.  Sums run in ascending-hash order (the
 * reference's order is its hash map's iteration order, i.e. unspecified; results agree to ~1e-15).
 * scal: [0]=U' [1]=logReadMagnitude [2]=logContainmentDenominator [3]=totalReadSeedFrequency [4]=dropped */
void orc_read_magnitudes(const i64* count, i64 U, i64 minSupport, double* logv, double* scal) {
    double magSq = 0.0, sum = 0.0; i64 kept = 0, low = 0, total = 0;
    for (i64 i = 0; i < U; i++) {
        total += count[i];
        if (count[i] < minSupport) { logv[i] = 0.0; low++; continue; }
        const double lc = log1p((double)count[i]);
        logv[i] = lc; magSq += lc * lc; sum += lc; kept++;
    }
    scal[0] = (double)kept; scal[1] = sqrt(magSq); scal[2] = sum; scal[3] = (double)total; scal[4] = (double)low;
}

static double lookup_log(const u64* hash, const double* logv, i64 U, u64 key) {
    i64 lo = 0, hi = U;
    while (lo < hi) { const i64 mid = (lo + hi) >> 1; if (hash[mid] < key) lo = mid + 1; else hi = mid; }
    return (lo < U && hash[lo] == key) ? logv[lo] : 0.0;  /* 0 == "not in logReadCounts" (log1p(c>=1) > 0) */
}

/* ---- placement.cpp:1863-1876 : weighted-containment denominator from the ROOT's deltas ---- */
double orc_weighted_denominator(const u64* dHash, const int16_t* dChild, u64 rootBegin, u64 rootEnd,
                                const u64* hash, const double* logv, i64 U) {
    double d = 0.0;
    for (u64 i = rootBegin; i < rootEnd; i++)
        if (dChild[i] > 0 && lookup_log(hash, logv, U, dHash[i]) > 0.0) d += 1.0 / (double)dChild[i];
    return d;
}

/* ---- placement.cpp:242-345 computeChildMetrics, one node's delta range, accumulators in m[7] ----------
 * m: 0 logRawNum, 1 logCosNum, 2 presence (integer held in a double), 3 wcNum, 4 logContNum, 5 gMagSq, 6 gUnique */
void orc_child_metrics(double* m, const u64* dHash, const int16_t* dParent, const int16_t* dChild, u64 begin, u64 end,
                       const u64* hash, const double* logv, i64 U) {
    for (u64 i = begin; i < end; i++) {
        const i64 p = dParent[i], c = dChild[i];
        const double logC = c > 0 ? log1p((double)c) : 0.0;
        const double logP = p > 0 ? log1p((double)p) : 0.0;
        m[5] += logC * logC - logP * logP;
        m[6] += (double)((c > 0) - (p > 0));
        if (c - p == 0) continue;
        const double lr = lookup_log(hash, logv, U, dHash[i]);
        if (!(lr > 0.0)) continue;
        const i64 pd = (i64)((p == 0) & (c != 0)) - (i64)((c == 0) & (p != 0));
        m[2] += (double)pd;
        m[0] += (c > 0 ? lr / (double)c : 0.0) - (p > 0 ? lr / (double)p : 0.0);
        m[1] += lr * (logC - logP);
        m[3] += (c > 0 ? 1.0 / (double)c : 0.0) - (p > 0 ? 1.0 / (double)p : 0.0);
        m[4] += (double)pd * lr;
    }
}

/* ---- placement.hpp:120-149 score getters ---- */
void orc_scores(const double* m, double U1, double mag, double denL, double denW, double* s) {
    s[0] = mag > 0.0 ? m[0] / mag : 0.0;
    {
        const double g = sqrt(m[5]);
        double v = (mag <= 0.0 || g <= 0.0) ? 0.0 : m[1] / (mag * g);
        if (!(mag <= 0.0 || g <= 0.0)) { if (v < 0.0) v = 0.0; if (v > 1.0) v = 1.0; }
        s[1] = v;
    }
    s[2] = U1 > 0.0 ? m[2] / U1 : 0.0;
    s[3] = denW > 0.0 ? m[3] / denW : 0.0;
    s[4] = denL > 0.0 ? m[4] / denL : 0.0;
}

/* ---- panmap_utils.cpp:260-289 + placement.cpp:742-827 : BFS visit order ---------------------------------
 * children are appended in ascending DFS index, levels are visited in order => order[] lists DFS indices in
 * the exact sequence the 1-thread reference scores them. */
void orc_bfs_order(const u32* parentIdx, u64 N, u32* order) {
    if (N == 0) return;
    u32* childCount = (u32*)calloc((size_t)N + 1, sizeof(u32));
    for (u64 v = 1; v < N; v++) childCount[parentIdx[v] + 1]++;
    for (u64 v = 0; v < N; v++) childCount[v + 1] += childCount[v];     /* CSR offsets */
    u32* fill = (u32*)malloc(sizeof(u32) * (size_t)N);
    memcpy(fill, childCount, sizeof(u32) * (size_t)N);
    u32* kids = (u32*)malloc(sizeof(u32) * (size_t)N);
    for (u64 v = 1; v < N; v++) kids[fill[parentIdx[v]]++] = (u32)v;     /* ascending v per parent */
    u64 head = 0, tail = 0;
    order[tail++] = 0;
    while (head < tail) {
        const u32 v = order[head++];
        for (u32 e = childCount[v]; e < childCount[v + 1]; e++) order[tail++] = kids[e];
    }
    free(childCount); free(fill); free(kids);
}

/* ---- placement.cpp:701-827 : per-node accumulators (parent copy + own deltas) and scores, all nodes -----
 * metrics [N][7], scores [N][5], indexed by DFS index. */
void orc_node_metrics(const u64* dHash, const int16_t* dParent, const int16_t* dChild, const u64* off, const u32* parentIdx,
                      u64 N, const u64* hash, const double* logv, i64 U, double U1, double mag, double denL, double denW,
                      double* metrics, double* scores) {
    u32* order = (u32*)malloc(sizeof(u32) * (size_t)(N ? N : 1));
    orc_bfs_order(parentIdx, N, order);
    for (u64 i = 0; i < N; i++) {
        const u32 v = order[i];
        double* m = metrics + 7 * (size_t)v;
        if (v == 0) memset(m, 0, 7 * sizeof(double));
        else memcpy(m, metrics + 7 * (size_t)parentIdx[v], 7 * sizeof(double));
        orc_child_metrics(m, dHash, dParent, dChild, off[v], off[v + 1], hash, logv, U);
        orc_scores(m, U1, mag, denL, denW, scores + 5 * (size_t)v);
    }
    free(order);
}

/* ---- placement.cpp:355-371 update*Score chain + :395-401 finalizeTiedIndices -----------------------------
 * order/score: the visit sequence (eligible nodes only).  tied (cap tiedCap) receives the sorted unique tie
 * list; returns its length.  bestIdx = ORC_NONE when no node ever scored above tolerance. */
i64 orc_select_chain(const u32* order, const double* score, i64 n, double* bestScore, u32* bestIdx, u32* tied, i64 tiedCap) {
    double best = 0.0; u32 bi = ORC_NONE;
    u32* tv = (u32*)malloc(sizeof(u32) * (size_t)(2 * n + 2));
    i64 tn = 0;
    for (i64 i = 0; i < n; i++) {
        const double sc = score[i]; const u32 v = order[i];
        const double tol = fmax(best * 0.0001, 1e-9);
        if (sc > best + tol) { best = sc; bi = v; tn = 0; tv[tn++] = v; }
        else if (sc >= best - tol && sc > 0) {
            if (tn == 0 || tv[tn - 1] != bi) tv[tn++] = bi;
            if (v != bi) tv[tn++] = v;
        }
    }
    i64 m = 0;
    if (tn > 0) {
        /* sort + unique */
        for (i64 i = 1; i < tn; i++) { u32 x = tv[i]; i64 j = i - 1; while (j >= 0 && tv[j] > x) { tv[j + 1] = tv[j]; j--; } tv[j + 1] = x; }
        for (i64 i = 0; i < tn; i++) if (i == 0 || tv[i] != tv[i - 1]) { if (m < tiedCap) tied[m] = tv[i]; m++; }
        bi = tv[0];
    }
    free(tv);
    *bestScore = best; *bestIdx = bi;
    return m;
}

/* ---- placement.cpp:986-1950 placeLite, compute part: reads + flat index -> 5 x (score, best, ties) ------
 * leafMask may be NULL; with forceLeaf only nodes without children are eligible (:794-795); skipNode is the
 * leave-one-out index (ORC_NONE = none).  tied: [5][tiedCap]; nodeScores (may be NULL): [N][5].
 * stats: [0]=U (unique seeds after homopolymer/mask) [1]=minSupport [2]=U' [3]=logReadMagnitude
 *        [4]=logContDenom [5]=wcDenom [6]=totalReadSeedFrequency */
int orc_place_q(const char* reads, const char* quals, const u64* readOff, u64 n_reads,
                const u64* dHash, const int16_t* dParent, const int16_t* dChild, const u64* off, const u32* parentIdx, u64 N,
                int k, int s, int t, int l, int open, int trimStart, int trimEnd, int dedup, int minReadSupport,
                double seedMaskFraction, int forceLeaf, u32 skipNode, int minSeedQuality,
                double* bestScore, u32* bestIdx, i64* tiedCount, u32* tied, i64 tiedCap, double* nodeScores, double* stats) {
    u64* H; i64* C;
    i64 U = (minSeedQuality > 0 && quals)   /* placement.cpp:1179,1388: the quality path replaces the seed construction (and ignores dedup) */
                ? orc_seed_table_quality(reads, quals, readOff, n_reads, k, s, t, l, open, trimStart, trimEnd, minSeedQuality, &H, &C)
                : orc_seed_table(reads, readOff, n_reads, k, s, t, l, open, trimStart, trimEnd, dedup, &H, &C);
    if (n_reads > 0) U = orc_mask_top_seeds(H, C, U, seedMaskFraction);
    const i64 ms = orc_resolve_min_read_support(C, U, minReadSupport);
    double* logv = (double*)malloc(sizeof(double) * (size_t)(U ? U : 1));
    double scal[5];
    orc_read_magnitudes(C, U, ms, logv, scal);
    const double denW = N ? orc_weighted_denominator(dHash, dChild, off[0], off[1], H, logv, U) : 0.0;
    double* metrics = (double*)malloc(sizeof(double) * 7 * (size_t)(N ? N : 1));
    double* scores = nodeScores ? nodeScores : (double*)malloc(sizeof(double) * 5 * (size_t)(N ? N : 1));
    orc_node_metrics(dHash, dParent, dChild, off, parentIdx, N, H, logv, U, scal[0], scal[1], scal[2], denW, metrics, scores);
    u32* order = (u32*)malloc(sizeof(u32) * (size_t)(N ? N : 1));
    orc_bfs_order(parentIdx, N, order);
    uint8_t* hasChild = (uint8_t*)calloc((size_t)(N ? N : 1), 1);
    for (u64 v = 1; v < N; v++) hasChild[parentIdx[v]] = 1;
    u32* eo = (u32*)malloc(sizeof(u32) * (size_t)(N ? N : 1));
    double* es = (double*)malloc(sizeof(double) * (size_t)(N ? N : 1));
    for (int mtr = 0; mtr < 5; mtr++) {
        i64 n = 0;
        for (u64 i = 0; i < N; i++) {
            const u32 v = order[i];
            if (v == skipNode || (forceLeaf && hasChild[v])) continue;
            eo[n] = v; es[n] = scores[5 * (size_t)v + mtr]; n++;
        }
        tiedCount[mtr] = orc_select_chain(eo, es, n, &bestScore[mtr], &bestIdx[mtr], tied + (size_t)mtr * (size_t)tiedCap, tiedCap);
    }
    stats[0] = (double)U; stats[1] = (double)ms; stats[2] = scal[0]; stats[3] = scal[1]; stats[4] = scal[2]; stats[5] = denW; stats[6] = scal[3];
    free(H); free(C); free(logv); free(metrics); if (!nodeScores) free(scores);
    free(order); free(hasChild); free(eo); free(es);
    return 0;
}

int orc_place(const char* reads, const u64* readOff, u64 n_reads,
              const u64* dHash, const int16_t* dParent, const int16_t* dChild, const u64* off, const u32* parentIdx, u64 N,
              int k, int s, int t, int l, int open, int trimStart, int trimEnd, int dedup, int minReadSupport,
              double seedMaskFraction, int forceLeaf, u32 skipNode,
              double* bestScore, u32* bestIdx, i64* tiedCount, u32* tied, i64 tiedCap, double* nodeScores, double* stats) {
    return orc_place_q(reads, NULL, readOff, n_reads, dHash, dParent, dChild, off, parentIdx, N, k, s, t, l, open, trimStart, trimEnd, dedup,
                       minReadSupport, seedMaskFraction, forceLeaf, skipNode, 0, bestScore, bestIdx, tiedCount, tied, tiedCap, nodeScores, stats);
}
