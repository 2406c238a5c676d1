// ORACLE SCAFFOLDING (test infrastructure, not product code; never linked into the product).
// extern "C" entry points over the reference's UNMODIFIED translation units
// (/root/reference/src/{seeding,placement,panmap_utils,index_single_mode}.cpp, compiled from where they
// lie via the symlink farm in oracle/_ref/src).  Used by tests/ to pin oracle/panmap_oracle.c, by
// tools/make_golden.py to generate tests/golden/*, and by bench.py --impl reference as the CPU arm.
#include "index_single_mode.hpp"
#include "logging.hpp"
#include "panmap_utils.hpp"
#include "placement.hpp"
#include "seeding.hpp"
#include "panman_loader.hpp"

#include <boost/iostreams/device/mapped_file.hpp>
#include <capnp/serialize.h>
#include <tbb/global_control.h>

#include <chrono>
#include <cstring>
#include <iostream>
#include <sstream>
#include <cstring>
#include <fstream>
#include <memory>
#include <queue>

// link stubs for symbols the compiled-in TUs reference but the oracle never reaches
extern "C" int64_t score_reads_vs_reference(const char*, int, const char**, const int*, int, bool) { return 0; }
namespace panmap_zstd {
bool compressToFile(const void*, size_t, const std::string&, int, int, size_t, const void*, size_t) { return false; }
bool decompressFromFile(const std::string&, std::vector<uint8_t>&, int, size_t) { return false; }
}  // namespace panmap_zstd

namespace {
struct LoadedIndex {
    boost::iostreams::mapped_file_source mm;
    std::unique_ptr<capnp::FlatArrayMessageReader> rd;
    panmapUtils::LiteTree tree;
};
capnp::ReaderOptions opts() {
    capnp::ReaderOptions o; o.traversalLimitInWords = kj::maxValue; o.nestingLimit = 1024; return o;
}
std::string g_err;
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void ref_hash_seq(const char* s, int n, uint64_t* f, uint64_t* r) {
    auto p = seeding::hashSeq(std::string(s, n)); *f = p.first; *r = p.second;
}

// reference seeding::rollingSyncmers; returns tuple count (writes at most cap tuples)
int64_t ref_rolling_syncmers(const char* seq, int64_t len, int k, int s, int open, int t, int returnAll,
                             uint64_t* hash, uint8_t* isRev, uint8_t* isSync, int64_t* pos, int64_t cap) {
    auto v = seeding::rollingSyncmers(std::string_view(seq, static_cast<size_t>(len)), k, s, open != 0, t, returnAll != 0);
    int64_t n = 0;
    for (auto& [h, rev, syn, p] : v) {
        if (n < cap) { hash[n] = h; isRev[n] = rev; isSync[n] = syn; pos[n] = p; }
        ++n;
    }
    return n;
}

// .panman -> reference IndexBuilder -> uncompressed .idx (reference writeIndex format)
int ref_build_index(const char* panmanPath, const char* idxPath, int k, int s, int t, int l, int open,
                    int flankMask, int hpc, int threads) {
    try {
        output::init(true, false, true);
        panmanUtils::Tree T;
        panman_loader::load(panmanPath, T);
        tbb::global_control gc(tbb::global_control::max_allowed_parallelism, threads > 0 ? threads : 1);
        index_single_mode::IndexBuilder b(&T, k, s, t, l, open != 0, flankMask, hpc != 0, false, false);
        b.buildIndexParallel(threads > 0 ? threads : 1);
        b.writeIndex(idxPath, 1, 7, true);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// genome string of one node via the reference's own reconstruction (panmap_utils.cpp:7-180)
int64_t ref_node_genome(const char* panmanPath, const char* nodeId, char* out, int64_t cap) {
    try {
        output::init(true, false, true);
        panmanUtils::Tree T;
        panman_loader::load(panmanPath, T);
        std::string g = panmapUtils::getStringFromReference(&T, nodeId, false);
        if (static_cast<int64_t>(g.size()) <= cap) std::memcpy(out, g.data(), g.size());
        return static_cast<int64_t>(g.size());
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

void* ref_index_open(const char* idxPath) {
    try {
        output::init(true, false, true);
        auto* L = new LoadedIndex();
        L->mm.open(idxPath);
        const auto* words = reinterpret_cast<const capnp::word*>(L->mm.data() + index_single_mode::kIndexHeaderSize);
        const size_t nw = (L->mm.size() - index_single_mode::kIndexHeaderSize) / sizeof(capnp::word);
        L->rd = std::make_unique<capnp::FlatArrayMessageReader>(kj::ArrayPtr<const capnp::word>(words, nw), opts());
        L->tree.initialize(L->rd->getRoot<LiteIndex>().getLiteTree());
        return L;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void ref_index_close(void* h) { delete static_cast<LoadedIndex*>(h); }
int64_t ref_index_num_nodes(void* h) { return static_cast<int64_t>(static_cast<LoadedIndex*>(h)->tree.dfsIndexToNode.size()); }

struct RefPlaceOut {
    double bestScore[5];
    uint32_t bestIndex[5];
    int64_t tiedCount[5];
    int64_t totalReads;
    uint64_t readUniqueSeedCount;
    int64_t totalReadSeedFrequency;
    double readMagnitude;
    int64_t uniqueSeedsAfterFilters;  // size of result.seedFreqInReads
    double seconds;                   // wall time of placeLite
};

// reference placement::placeLite with CLI-default params; tied index lists / seed table fetched afterwards
// Stage timers of the reference itself: placeLite logs "Total read processing", "K-minimizer extraction" / "Seed extraction", "Read
// deduplication" and "Tree traversal" in ms through output::debug / output::info (placement.cpp:1275,1376,1691,1701,1929), which write
// to std::cerr when verbose.  With the timers on, ref_place turns verbose on, captures std::cerr for the duration of the call and
// parses those lines; nothing in the reference is modified.
int g_stageTimers = 0;
double g_stageMs[5] = {0, 0, 0, 0, 0};   // read processing (parse + dedup + seeding), seeding alone, dedup, tree traversal, whole call
void ref_set_stage_timers(int on) { g_stageTimers = on; }
void ref_last_stage_ms(double* out) { for (int i = 0; i < 5; ++i) out[i] = g_stageMs[i]; }
static double grabMs(const std::string& log, const char* key) {
    const size_t p = log.find(key);
    if (p == std::string::npos) return -1.0;
    return std::atof(log.c_str() + p + std::strlen(key));
}
int g_minSeedQuality = 0;   // --min-seed-quality of the next ref_place calls (0 = off, the CLI default)
void ref_set_min_seed_quality(int q) { g_minSeedQuality = q; }
struct RefPlaceKeep { placement::PlacementResult res; };
void* ref_place(void* h, const char* r1, const char* r2, const char* outTsv, int threads, int minReadSupport,
                double seedMaskFraction, int trimStart, int trimEnd, int dedup, int forceLeaf, int storeDiag,
                RefPlaceOut* out) {
    try {
        auto* L = static_cast<LoadedIndex*>(h);
        tbb::global_control gc(tbb::global_control::max_allowed_parallelism, threads > 0 ? threads : 1);
        auto* K = new RefPlaceKeep();
        placement::TraversalParams tp;
        tp.seedMaskFraction = seedMaskFraction;  // CLI default is 0 (main.cpp:1967), struct default 0.001
        tp.minReadSupport = minReadSupport;
        tp.trimStart = trimStart; tp.trimEnd = trimEnd; tp.dedupReads = dedup != 0; tp.forceLeaf = forceLeaf != 0;
        tp.store_diagnostics = storeDiag != 0;
        tp.minSeedQuality = g_minSeedQuality;
        std::string o = outTsv ? outTsv : "";
        std::ostringstream captured;
        std::streambuf* oldBuf = nullptr;
        const bool wasVerbose = output::config().verbose, wasPlain = output::config().plain, wasQuiet = output::config().quiet;
        if (g_stageTimers) { output::config().verbose = true; output::config().plain = true; output::config().quiet = false; oldBuf = std::cerr.rdbuf(captured.rdbuf()); }
        auto t0 = std::chrono::steady_clock::now();
        try { placement::placeLite(K->res, &L->tree, *L->rd, r1 ? r1 : "", r2 ? r2 : "", o, tp, nullptr); }
        catch (...) { if (oldBuf) std::cerr.rdbuf(oldBuf); output::config().verbose = wasVerbose; output::config().plain = wasPlain; output::config().quiet = wasQuiet; throw; }
        auto t1 = std::chrono::steady_clock::now();
        if (g_stageTimers) {
            std::cerr.rdbuf(oldBuf); output::config().verbose = wasVerbose; output::config().plain = wasPlain; output::config().quiet = wasQuiet;
            const std::string log = captured.str();
            g_stageMs[0] = grabMs(log, "Total read processing: ");
            g_stageMs[1] = std::max(grabMs(log, "K-minimizer extraction: "), grabMs(log, "Seed extraction: "));
            g_stageMs[2] = grabMs(log, "Read deduplication: ");
            g_stageMs[3] = grabMs(log, "Tree traversal: ");
            g_stageMs[4] = 1e3 * std::chrono::duration<double>(t1 - t0).count();
        }
        auto& R = K->res;
        const double sc[5] = {R.bestLogRawScore, R.bestLogCosineScore, R.bestContainmentScore,
                              R.bestWeightedContainmentScore, R.bestLogContainmentScore};
        const uint32_t ix[5] = {R.bestLogRawNodeIndex, R.bestLogCosineNodeIndex, R.bestContainmentNodeIndex,
                                R.bestWeightedContainmentNodeIndex, R.bestLogContainmentNodeIndex};
        const std::vector<uint32_t>* td[5] = {&R.tiedLogRawNodeIndices, &R.tiedLogCosineNodeIndices,
                                              &R.tiedContainmentNodeIndices, &R.tiedWeightedContainmentNodeIndices,
                                              &R.tiedLogContainmentNodeIndices};
        for (int i = 0; i < 5; ++i) { out->bestScore[i] = sc[i]; out->bestIndex[i] = ix[i]; out->tiedCount[i] = td[i]->size(); }
        out->totalReads = R.totalReadsProcessed;
        out->readUniqueSeedCount = R.readUniqueSeedCount;
        out->totalReadSeedFrequency = R.totalReadSeedFrequency;
        out->readMagnitude = R.readMagnitude;
        out->uniqueSeedsAfterFilters = static_cast<int64_t>(R.seedFreqInReads.size());
        out->seconds = std::chrono::duration<double>(t1 - t0).count();
        return K;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void ref_place_free(void* k) { delete static_cast<RefPlaceKeep*>(k); }
void ref_place_tied(void* k, int metric, uint32_t* out) {
    auto& R = static_cast<RefPlaceKeep*>(k)->res;
    const std::vector<uint32_t>* td[5] = {&R.tiedLogRawNodeIndices, &R.tiedLogCosineNodeIndices,
                                          &R.tiedContainmentNodeIndices, &R.tiedWeightedContainmentNodeIndices,
                                          &R.tiedLogContainmentNodeIndices};
    std::copy(td[metric]->begin(), td[metric]->end(), out);
}
// the read seed table the reference ended with (after homopolymer removal / masking), unsorted
void ref_place_seed_table(void* k, uint64_t* hashes, int64_t* counts) {
    auto& R = static_cast<RefPlaceKeep*>(k)->res;
    size_t i = 0;
    for (auto& [h, c] : R.seedFreqInReads) { hashes[i] = h; counts[i] = c; ++i; }
}

// Per-node accumulators + scores by composing the reference's own computeChildMetrics along the tree in the
// reference's BFS order (placement.cpp:742-827), from a given read seed table (hash,count) with the
// reference's resolveMinReadSupport / computeReadSeedMagnitudes / root-denominator code (placement.cpp:1839-1876).
// metrics: [N][7] = logRawNum, logCosNum, presence, wcNum, logContNum, gMagSq, gUnique ; scores: [N][5]
// scalars: [0]=minSupport [1]=U' [2]=logReadMagnitude [3]=logContDenom [4]=wcDenom [5]=totalFreq
int ref_node_metrics(void* h, const uint64_t* hashes, const int64_t* counts, int64_t nSeeds, int minReadSupport,
                     double* metrics, double* scores, double* scalars) {
    try {
        auto* L = static_cast<LoadedIndex*>(h);
        auto* tree = &L->tree;
        if (!tree->seedChangesLoaded) {  // run the reference's own SoA hookup by placing zero reads
            placement::PlacementResult tmp; std::string o = "/dev/null";
            placement::TraversalParams tp; tp.seedMaskFraction = 0;
            placement::placeLite(tmp, tree, *L->rd, "", "", o, tp, nullptr);
        }
        placement::PlacementGlobalState st;
        for (int64_t i = 0; i < nSeeds; ++i) st.seedFreqInReads[hashes[i]] = counts[i];
        const int64_t ms = placement::resolveMinReadSupport(st.seedFreqInReads, minReadSupport);
        placement::computeReadSeedMagnitudes(st, ms);
        st.liteTree = tree; st.root = tree->root;
        tree->forEachSeedChange(tree->root->seedChangeOffset, tree->root->seedChangeSize,
                                [&](uint64_t sh, int64_t, int64_t cc) {
                                    if (cc > 0 && st.logReadCounts.contains(sh)) st.weightedContainmentDenominator += 1.0 / static_cast<double>(cc);
                                });
        scalars[0] = static_cast<double>(ms); scalars[1] = static_cast<double>(st.readUniqueSeedCount);
        scalars[2] = st.logReadMagnitude; scalars[3] = st.logContainmentDenominator;
        scalars[4] = st.weightedContainmentDenominator; scalars[5] = static_cast<double>(st.totalReadSeedFrequency);
        const size_t N = tree->dfsIndexToNode.size();
        std::vector<placement::NodeMetrics> M(N);
        std::queue<panmapUtils::LiteNode*> q; q.push(tree->root);
        while (!q.empty()) {
            auto* n = q.front(); q.pop();
            placement::NodeMetrics m = n->parent ? M[n->parent->nodeIndex] : placement::NodeMetrics{};
            placement::NodeMetrics::computeChildMetrics(m, n->seedChangeOffset, n->seedChangeSize, st);
            M[n->nodeIndex] = m;
            double* o = metrics + 7 * static_cast<size_t>(n->nodeIndex);
            o[0] = m.logRawNumerator; o[1] = m.logCosineNumerator; o[2] = static_cast<double>(static_cast<int64_t>(m.presenceIntersectionCount));
            o[3] = m.weightedContainmentNumerator; o[4] = m.logContainmentNumerator; o[5] = m.genomeMagnitudeSquared;
            o[6] = static_cast<double>(static_cast<int64_t>(m.genomeUniqueSeedCount));
            double* s = scores + 5 * static_cast<size_t>(n->nodeIndex);
            s[0] = m.getLogRawScore(st.logReadMagnitude); s[1] = m.getLogCosineScore(st.logReadMagnitude);
            s[2] = m.getContainmentScore(st.readUniqueSeedCount);
            s[3] = m.getWeightedContainmentScore(st.weightedContainmentDenominator);
            s[4] = m.getLogContainmentScore(st.logContainmentDenominator);
            for (auto* c : n->children) q.push(c);
        }
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// reference tolerance chain (placement.cpp:355-401) driven over an explicit score sequence, for unit tests
int64_t ref_select_chain(const uint32_t* order, const double* score, int64_t n, double* bestScore, uint32_t* bestIdx,
                         uint32_t* tied, int64_t tiedCap) {
    placement::PlacementResult R;
    for (int64_t i = 0; i < n; ++i) R.updateLogRawScore(order[i], score[i]);
    panmapUtils::LiteTree dummy;
    R.resolveNodeIds(&dummy);
    *bestScore = R.bestLogRawScore; *bestIdx = R.bestLogRawNodeIndex;
    int64_t m = static_cast<int64_t>(R.tiedLogRawNodeIndices.size());
    for (int64_t i = 0; i < m && i < tiedCap; ++i) tied[i] = R.tiedLogRawNodeIndices[i];
    return m;
}

// flat arrays -> a real uncompressed .idx (LiteIndex capnp message + PMI1 header, as IndexBuilder::writeIndex lays it out),
// so the reference's placeLite can consume synthetic indexes byte-for-byte identical to what the GPU path is given
int ref_write_index(const char* path, const uint64_t* hash, const int16_t* par, const int16_t* chi, const uint64_t* off,
                    const uint32_t* parent, uint64_t N, uint64_t D, int k, int s, int t, int l, int openAndHpc) {
    const int open = openAndHpc & 1; const bool hpc = (openAndHpc >> 1) & 1;   // bit 1: index built with --hpc
    try {
        capnp::MallocMessageBuilder msg;
        auto idx = msg.initRoot<LiteIndex>();
        idx.setK(k); idx.setS(s); idx.setT(t); idx.setL(l); idx.setOpen(open != 0); idx.setHpc(hpc);
        idx.setFormatVersion(panmapUtils::INDEX_FORMAT_VERSION);
        auto tree = idx.initLiteTree();
        auto nodes = tree.initLiteNodes(static_cast<unsigned>(N));
        for (uint64_t i = 0; i < N; ++i) {
            nodes[i].setId("node_" + std::to_string(i));
            nodes[i].setParentIndex(i ? parent[i] : 0);
        }
        tree.initBlockRanges(0);
        const uint64_t SEG = panmapUtils::LiteTree::SEED_CHANGE_SEGMENT;
        const unsigned nSeg = static_cast<unsigned>(D ? (D + SEG - 1) / SEG : 1);
        auto H = idx.initSeedChangeHashes(nSeg); auto P = idx.initSeedChangeParentCounts(nSeg); auto Cc = idx.initSeedChangeChildCounts(nSeg);
        for (unsigned g = 0; g < nSeg; ++g) {
            const uint64_t b = g * SEG, n = std::min<uint64_t>(SEG, D - b);
            auto h = H.init(g, static_cast<unsigned>(n)); auto p = P.init(g, static_cast<unsigned>(n)); auto c = Cc.init(g, static_cast<unsigned>(n));
            for (uint64_t i = 0; i < n; ++i) { h.set(i, hash[b + i]); p.set(i, par[b + i]); c.set(i, chi[b + i]); }
        }
        auto O = idx.initNodeChangeOffsets(static_cast<unsigned>(N + 1));
        for (uint64_t i = 0; i <= N; ++i) O.set(i, off[i]);
        kj::Array<capnp::word> flat = capnp::messageToFlatArray(msg);
        index_single_mode::IndexParamsHeader ph; ph.k = k; ph.s = s; ph.t = t; ph.l = l; ph.hpc = hpc; ph.open = open != 0; ph.uncompressed = true;
        const auto header = index_single_mode::encodeIndexHeader(ph);
        std::ofstream out(path, std::ios::binary | std::ios::trunc);
        out.write(reinterpret_cast<const char*>(header.data()), header.size());
        out.write(reinterpret_cast<const char*>(flat.begin()), static_cast<std::streamsize>(flat.size() * sizeof(capnp::word)));
        return out ? 0 : -1;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

}  // extern "C"
