// pm_synth.cpp -- synthetic PanMAN-shaped inputs for bench.py and the large-size parity tests:
// random genome, a tree generated directly in DFS pre-order, Poisson SNPs per edge, EXACT per-node seed deltas
// (difference of the k-min-mer / syncmer multisets of child vs parent genome, sorted by hash) and error-bearing reads
// from a truth leaf.  This is a miniature index builder, incremental per SNP, not part of the product library.
// Seeds follow the same definitions as the placement path (pm_logic.cuh == seeding.cpp / placement.cpp:1598-1686).
#include "../../panmap_b200/csrc/pm_logic.cuh"

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

using namespace pm;

namespace {
struct Gen {
    int k, s, t, l, open;
    std::string g;
    // syncmers of the current genome by start position (flat: the bacterial-scale workload keeps 1.5 M of them and touches them 10^8 times)
    std::vector<u64> synHash; std::vector<uint8_t> synHas;
    int nextSyn(int pos) const { const int G = (int)synHas.size(); while (pos < G && !synHas[pos]) ++pos; return pos < G ? pos : -1; }   // first syncmer at >= pos
    int prevSyn(int pos) const { while (pos >= 0 && !synHas[pos]) --pos; return pos; }                                                  // last syncmer at <= pos
    // seed hash -> multiplicity in the current genome: open addressing, no deletion (a count may return to 0 and stay)
    struct Counts {
        std::vector<u64> key; std::vector<int> val; std::vector<uint8_t> used; size_t n = 0, mask = 0;
        void init(size_t cap) { size_t c = 1024; while (c < cap) c <<= 1; key.assign(c, 0); val.assign(c, 0); used.assign(c, 0); mask = c - 1; n = 0; }
        size_t slot(u64 h) const { size_t i = (size_t)mixKey(h) & mask; while (used[i] && key[i] != h) i = (i + 1) & mask; return i; }
        int get(u64 h) const { const size_t i = slot(h); return used[i] ? val[i] : 0; }
        void set(u64 h, int v) {
            size_t i = slot(h);
            if (!used[i]) {
                if ((n + 1) * 10 > (mask + 1) * 6) { grow(); i = slot(h); }
                used[i] = 1; key[i] = h; ++n;
            }
            val[i] = v;
        }
        void grow() {
            std::vector<u64> k2; std::vector<int> v2; std::vector<uint8_t> u2;
            k2.swap(key); v2.swap(val); u2.swap(used);
            init((mask + 1) * 2);
            for (size_t j = 0; j < k2.size(); ++j) if (u2[j]) { const size_t i = slot(k2[j]); used[i] = 1; key[i] = k2[j]; val[i] = v2[j]; ++n; }
        }
    } counts;
    SeedTables T; SeederParams P; std::vector<u64> ring;

    void init(int k_, int s_, int t_, int l_, int open_) {
        k = k_; s = s_; t = t_; l = l_; open = open_;
        buildSeedTables(T, k, s); P = makeSeederParams(k, s, t, 1, open, 0, 0);
        ring.assign((size_t)seederRingWords(k, s, 1), 0);
    }
    // syncmers with start position in [a, b] (clamped) of the current genome
    void syncmersIn(int a, int b, std::vector<std::pair<int, u64>>& out) {
        out.clear();
        const int G = (int)g.size();
        a = std::max(a, 0); b = std::min(b, G - k);
        if (a > b) return;
        ReadSeederT<1> sd; sd.reset(ring.data(), k - s + 1);
        for (int i = a; i <= b + k - 1; ++i) {
            u64 h; bool rev;
            if (sd.pushBase(i - a, baseCode((unsigned char)g[i]), T, P, h, rev)) out.push_back({i - k + 1, h});
        }
    }
    // seeds (k-min-mers for l > 1, syncmers otherwise) of a run of consecutive syncmers
    void seedsOf(const std::vector<u64>& h, std::vector<u64>& out) const {
        if (l <= 1) { out.insert(out.end(), h.begin(), h.end()); return; }
        for (size_t j = 0; j + l <= h.size(); ++j) {
            u64 fw = 0, rw = 0;
            for (int w = 0; w < l; ++w) { fw ^= rol64(h[j + w], (unsigned)((k * (l - 1 - w)) & 63)); rw ^= rol64(h[j + w], (unsigned)((k * w) & 63)); }
            if (fw != rw) out.push_back(fw < rw ? fw : rw);
        }
    }
};
struct Undo {
    uint32_t node;
    std::vector<std::pair<int, char>> bases;
    std::vector<std::pair<int, u64>> synRemoved; std::vector<int> synAdded;
    std::vector<std::pair<u64, int>> countOld;  // (hash, previous count)
};
struct Out {
    std::vector<u64> hash; std::vector<int16_t> par, chi; std::vector<u64> off; std::vector<uint32_t> parent;
    std::string reads; std::vector<u64> readOff; uint32_t truth = 0; std::string truthGenome;
};
Out* g_out = nullptr;
}  // namespace

extern "C" {

// returns 0; results fetched with synth_get_* until the next call / synth_free()
int synth_generate(uint64_t nNodes, uint64_t genomeLen, double lambda, uint64_t nReads, int readLen, double subRate, double nRate,
                   int k, int s, int t, int l, int open, uint64_t seed, double truthFrac) {
    delete g_out; g_out = new Out();
    Out& O = *g_out;
    std::mt19937_64 rg(seed + 42), rt(seed + 43), rm(seed + 44), rr(seed + 45);
    Gen G; G.init(k, s, t, l, open);
    G.g.resize(genomeLen);
    G.synHash.assign(genomeLen, 0); G.synHas.assign(genomeLen, 0); G.counts.init(genomeLen);
    const char B[4] = {'A', 'C', 'G', 'T'};
    for (auto& c : G.g) c = B[rg() & 3];
    O.parent.assign(nNodes, 0); O.off.assign(nNodes + 1, 0);
    std::vector<Undo> stack;
    std::poisson_distribution<int> pois(lambda > 0 ? lambda : 1e-9);
    std::vector<std::pair<int, u64>> tmpSyn;
    std::vector<u64> oldH, newH, oldSeeds, newSeeds;
    const uint64_t truthTarget = (uint64_t)(truthFrac * (double)(nNodes - 1));
    std::string truthGenome; bool truthFixed = false; uint32_t truthCand = 0;

    auto applyCounts = [&](Undo& u, std::vector<u64>& minus, std::vector<u64>& plus, std::vector<std::pair<u64, int>>& net) {
        for (u64 h : minus) net.push_back({h, -1});
        for (u64 h : plus) net.push_back({h, 1});
        (void)u;
    };

    for (uint64_t v = 0; v < nNodes; ++v) {
        // ---- attach ----
        if (v > 0) {
            uint32_t p;
            if (stack.size() <= 1 || (rt() & 1)) p = (uint32_t)(v - 1);
            else { p = stack[rt() % (stack.size() - 1)].node; }
            // a node is a leaf iff the next node does not hang under it
            if (!truthFixed && truthCand == v - 1 && v - 1 >= truthTarget && p != v - 1) { truthFixed = true; O.truth = (uint32_t)(v - 1); }
            while (!stack.empty() && stack.back().node != p) {  // backtrack: undo everything below p
                Undo& u = stack.back();
                for (auto it = u.countOld.rbegin(); it != u.countOld.rend(); ++it) G.counts.set(it->first, it->second);
                for (int pos : u.synAdded) G.synHas[pos] = 0;
                for (auto& pr : u.synRemoved) { G.synHas[pr.first] = 1; G.synHash[pr.first] = pr.second; }
                for (auto it = u.bases.rbegin(); it != u.bases.rend(); ++it) G.g[it->first] = it->second;
                stack.pop_back();
            }
            O.parent[v] = p;
        }
        stack.push_back(Undo()); Undo& U = stack.back(); U.node = (uint32_t)v;
        std::vector<std::pair<u64, int>> net;   // (seed, +-1) of this node; summed per seed below
        if (v == 0) {
            // root: all seeds of the genome from the empty genome
            G.syncmersIn(0, (int)genomeLen - k, tmpSyn);
            newH.clear();
            for (auto& pr : tmpSyn) { G.synHas[pr.first] = 1; G.synHash[pr.first] = pr.second; newH.push_back(pr.second); }
            newSeeds.clear(); G.seedsOf(newH, newSeeds);
            for (u64 h : newSeeds) net.push_back({h, 1});
        } else {
            int nm = lambda > 0 ? pois(rm) : 0;
            std::vector<int> posv;
            for (int i = 0; i < nm; ++i) {
                const int x = (int)(rm() % genomeLen);
                const char old = G.g[x];
                char alt = B[rm() & 3];
                while (alt == old) alt = B[rm() & 3];
                U.bases.push_back({x, old}); G.g[x] = alt; posv.push_back(x);
            }
            std::sort(posv.begin(), posv.end());
            // affected window ranges [x-k+1, x], merged; then extended by l-1 flanking syncmers and merged again
            std::vector<std::pair<int, int>> rng;
            for (int x : posv) {
                const int a = std::max(0, x - k + 1), b = std::min(x, (int)genomeLen - k);
                if (a > b) continue;
                if (!rng.empty() && a <= rng.back().second + 1) rng.back().second = std::max(rng.back().second, b); else rng.push_back({a, b});
            }
            const int fl = l > 1 ? l - 1 : 0;
            struct Seg { int a, b, lo, hi; };  // window range [a,b], extended syncmer position range [lo,hi]
            std::vector<Seg> segs;
            for (auto& r : rng) {
                Seg sg{r.first, r.second, r.first, r.second};
                int q = r.first;
                for (int i = 0; i < fl; ++i) { q = G.prevSyn(q - 1); if (q < 0) break; sg.lo = q; }
                q = r.second;
                for (int i = 0; i < fl; ++i) { q = G.nextSyn(q + 1); if (q < 0) break; sg.hi = q; }
                if (!segs.empty() && sg.lo <= segs.back().hi) {
                    // overlapping neighbourhoods: merge into one segment (windows between them cancel in the net diff)
                    segs.back().hi = std::max(segs.back().hi, sg.hi); segs.back().b = sg.b;
                    // remember the extra affected range by widening [a,b]; syncmers in the unaffected gap are recomputed too (same values)
                } else segs.push_back(sg);
            }
            for (auto& sg : segs) {
                oldH.clear(); newH.clear();
                std::vector<std::pair<int, u64>> oldIn;
                for (int q = G.nextSyn(sg.lo); q >= 0 && q <= sg.hi; q = G.nextSyn(q + 1)) { oldH.push_back(G.synHash[q]); if (q >= sg.a && q <= sg.b) oldIn.push_back({q, G.synHash[q]}); }
                // recompute the syncmers of every window in [a,b] on the mutated genome
                G.syncmersIn(sg.a, sg.b, tmpSyn);
                for (auto& pr : oldIn) { G.synHas[pr.first] = 0; U.synRemoved.push_back(pr); }
                for (auto& pr : tmpSyn) { G.synHas[pr.first] = 1; G.synHash[pr.first] = pr.second; U.synAdded.push_back(pr.first); }
                for (int q = G.nextSyn(sg.lo); q >= 0 && q <= sg.hi; q = G.nextSyn(q + 1)) newH.push_back(G.synHash[q]);
                oldSeeds.clear(); newSeeds.clear();
                G.seedsOf(oldH, oldSeeds); G.seedsOf(newH, newSeeds);
                applyCounts(U, oldSeeds, newSeeds, net);
            }
        }
        // ---- emit this node's deltas (sorted by hash) ----
        std::sort(net.begin(), net.end());
        std::vector<std::pair<u64, int>> ch;
        for (size_t i = 0; i < net.size();) {
            size_t j = i; int d = 0;
            while (j < net.size() && net[j].first == net[i].first) { d += net[j].second; ++j; }
            if (d != 0) ch.push_back({net[i].first, d});
            i = j;
        }
        for (auto& c : ch) {
            const int oldc = G.counts.get(c.first);
            const int newc = oldc + c.second;
            U.countOld.push_back({c.first, oldc});
            G.counts.set(c.first, newc);
            O.hash.push_back(c.first); O.par.push_back((int16_t)std::min(oldc, 32767)); O.chi.push_back((int16_t)std::min(newc, 32767));
        }
        O.off[v + 1] = O.hash.size();
        if (!truthFixed && v >= truthTarget) { truthCand = (uint32_t)v; truthGenome = G.g; }
    }
    if (!truthFixed) { O.truth = truthCand; }
    O.truthGenome = truthGenome;
    // ---- reads from the truth genome ----
    O.readOff.assign(nReads + 1, 0);
    O.reads.reserve(nReads * (size_t)readLen);
    std::uniform_real_distribution<double> uni(0.0, 1.0);
    const int RL = std::min<int>(readLen, (int)truthGenome.size());
    for (uint64_t r = 0; r < nReads; ++r) {
        const size_t st = (size_t)(rr() % (truthGenome.size() - RL + 1));
        std::string rd = truthGenome.substr(st, RL);
        if (rr() & 1) {  // reverse complement
            std::reverse(rd.begin(), rd.end());
            for (auto& c : rd) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
        }
        for (auto& c : rd) {
            const double u = uni(rr);
            if (u < nRate) c = 'N';
            else if (u < nRate + subRate) { char a = B[rr() & 3]; while (a == c) a = B[rr() & 3]; c = a; }
        }
        O.reads += rd; O.readOff[r + 1] = O.reads.size();
    }
    return 0;
}
uint64_t synth_num_deltas() { return g_out ? g_out->hash.size() : 0; }
uint64_t synth_reads_bytes() { return g_out ? g_out->reads.size() : 0; }
uint32_t synth_truth_node() { return g_out ? g_out->truth : 0; }
void synth_get_index(uint64_t* hash, int16_t* par, int16_t* chi, uint64_t* off, uint32_t* parent) {
    Out& O = *g_out;
    std::memcpy(hash, O.hash.data(), O.hash.size() * 8); std::memcpy(par, O.par.data(), O.par.size() * 2);
    std::memcpy(chi, O.chi.data(), O.chi.size() * 2); std::memcpy(off, O.off.data(), O.off.size() * 8);
    std::memcpy(parent, O.parent.data(), O.parent.size() * 4);
}
void synth_get_reads(char* reads, uint64_t* off) {
    Out& O = *g_out;
    std::memcpy(reads, O.reads.data(), O.reads.size()); std::memcpy(off, O.readOff.data(), O.readOff.size() * 8);
}
void synth_get_truth_genome(char* out) { std::memcpy(out, g_out->truthGenome.data(), g_out->truthGenome.size()); }
void synth_free() { delete g_out; g_out = nullptr; }
}  // extern "C"
