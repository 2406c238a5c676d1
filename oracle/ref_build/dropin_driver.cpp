// TEST SCAFFOLDING around the compiled drop-in (panmap_b200/integration/panmap_adapter.cpp): the reference's caller sequence of
// runPlacement (src/main.cpp:1668-1750) -- map the .idx, FlatArrayMessageReader, LiteTree::initialize, placement::placeLite -- with
// the adapter's placeLite (GPU) or, in the same binary, the reference's own placeLite (compiled from the unmodified placement.cpp with
// the symbol renamed on the command line: -DplaceLite=placeLite_cpu_reference).  Returns every PlacementResult field so that
// tests/test_gpu_dropin.py can diff the two.
#include "index_single_mode.hpp"
#include "logging.hpp"
#include "panmap_utils.hpp"
#include "placement.hpp"

#include <boost/iostreams/device/mapped_file.hpp>
#include <capnp/serialize.h>
#include <tbb/global_control.h>

#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace placement {   // the reference's own implementation under its build-time alias
void placeLite_cpu_reference(PlacementResult& result, panmapUtils::LiteTree* liteTree, ::capnp::MessageReader& liteIndex, const std::string& reads1,
                             const std::string& reads2, std::string& outputPath, const TraversalParams& params, panmanUtils::Tree* fullTree);
}
// link stubs for symbols the compiled-in TUs reference but this path never reaches
extern "C" int64_t score_reads_vs_reference(const char*, int, const char**, const int*, int, bool) { return 0; }

namespace {
struct Loaded {
    boost::iostreams::mapped_file_source mm;
    std::unique_ptr<capnp::FlatArrayMessageReader> rd;
    panmapUtils::LiteTree tree;
};
struct Kept { placement::PlacementResult res; };
std::string g_err;
}  // namespace

extern "C" {

const char* dropin_last_error() { return g_err.c_str(); }

void* dropin_open(const char* idxPath) {
    try {
        output::init(true, false, true);
        auto* L = new Loaded();
        L->mm.open(idxPath);
        const auto* words = reinterpret_cast<const capnp::word*>(L->mm.data() + index_single_mode::kIndexHeaderSize);
        const size_t nw = (L->mm.size() - index_single_mode::kIndexHeaderSize) / sizeof(capnp::word);
        capnp::ReaderOptions o; o.traversalLimitInWords = kj::maxValue; o.nestingLimit = 1024;
        L->rd = std::make_unique<capnp::FlatArrayMessageReader>(kj::ArrayPtr<const capnp::word>(words, nw), o);
        L->tree.initialize(L->rd->getRoot<LiteIndex>().getLiteTree());
        return L;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void dropin_close(void* h) { delete static_cast<Loaded*>(h); }

struct DropinOut {
    double bestScore[5]; uint32_t bestIndex[5]; int64_t tiedCount[5];
    int64_t totalReads; uint64_t readUniqueSeedCount; int64_t totalReadSeedFrequency; double readMagnitude;
    int64_t seedTableSize; int32_t k, s, t, open; int64_t nodeScoreRows; char bestId[5][64];
};

// which: 1 = the adapter (GPU), 0 = the reference's own placeLite
void* dropin_place(void* h, int which, const char* r1, const char* r2, const char* outTsv, int threads, int minReadSupport, double seedMaskFraction,
                   int trimStart, int trimEnd, int dedup, int forceLeaf, int storeDiag, int minSeedQuality, DropinOut* out) {
    try {
        auto* L = static_cast<Loaded*>(h);
        tbb::global_control gc(tbb::global_control::max_allowed_parallelism, threads > 0 ? threads : 1);
        auto* K = new Kept();
        placement::TraversalParams tp;
        tp.seedMaskFraction = seedMaskFraction; tp.minReadSupport = minReadSupport; tp.trimStart = trimStart; tp.trimEnd = trimEnd;
        tp.dedupReads = dedup != 0; tp.forceLeaf = forceLeaf != 0; tp.store_diagnostics = storeDiag != 0; tp.minSeedQuality = minSeedQuality;
        std::string o = outTsv ? outTsv : "";
        if (which) placement::placeLite(K->res, &L->tree, *L->rd, r1 ? r1 : "", r2 ? r2 : "", o, tp, nullptr);
        else placement::placeLite_cpu_reference(K->res, &L->tree, *L->rd, r1 ? r1 : "", r2 ? r2 : "", o, tp, nullptr);
        auto& R = K->res;
        const double sc[5] = {R.bestLogRawScore, R.bestLogCosineScore, R.bestContainmentScore, R.bestWeightedContainmentScore, R.bestLogContainmentScore};
        const uint32_t ix[5] = {R.bestLogRawNodeIndex, R.bestLogCosineNodeIndex, R.bestContainmentNodeIndex, R.bestWeightedContainmentNodeIndex,
                                R.bestLogContainmentNodeIndex};
        const std::vector<uint32_t>* td[5] = {&R.tiedLogRawNodeIndices, &R.tiedLogCosineNodeIndices, &R.tiedContainmentNodeIndices,
                                              &R.tiedWeightedContainmentNodeIndices, &R.tiedLogContainmentNodeIndices};
        const std::string* id[5] = {&R.bestLogRawNodeId, &R.bestLogCosineNodeId, &R.bestContainmentNodeId, &R.bestWeightedContainmentNodeId,
                                    &R.bestLogContainmentNodeId};
        std::memset(out, 0, sizeof(*out));
        for (int i = 0; i < 5; ++i) {
            out->bestScore[i] = sc[i]; out->bestIndex[i] = ix[i]; out->tiedCount[i] = static_cast<int64_t>(td[i]->size());
            std::strncpy(out->bestId[i], id[i]->c_str(), 63);
        }
        out->totalReads = R.totalReadsProcessed; out->readUniqueSeedCount = R.readUniqueSeedCount; out->totalReadSeedFrequency = R.totalReadSeedFrequency;
        out->readMagnitude = R.readMagnitude; out->seedTableSize = static_cast<int64_t>(R.seedFreqInReads.size());
        out->k = R.k; out->s = R.s; out->t = R.t; out->open = R.open ? 1 : 0; out->nodeScoreRows = static_cast<int64_t>(R.nodeScores.size());
        return K;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void dropin_free(void* k) { delete static_cast<Kept*>(k); }
void dropin_tied(void* k, int metric, uint32_t* out) {
    auto& R = static_cast<Kept*>(k)->res;
    const std::vector<uint32_t>* td[5] = {&R.tiedLogRawNodeIndices, &R.tiedLogCosineNodeIndices, &R.tiedContainmentNodeIndices,
                                          &R.tiedWeightedContainmentNodeIndices, &R.tiedLogContainmentNodeIndices};
    std::copy(td[metric]->begin(), td[metric]->end(), out);
}
void dropin_seed_table(void* k, uint64_t* hashes, int64_t* counts) {
    size_t i = 0;
    for (auto& [h, c] : static_cast<Kept*>(k)->res.seedFreqInReads) { hashes[i] = h; counts[i] = c; ++i; }
}
void dropin_node_scores(void* k, float* out) {
    auto& v = static_cast<Kept*>(k)->res.nodeScores;
    for (size_t i = 0; i < v.size(); ++i) for (int m = 0; m < 5; ++m) out[5 * i + m] = v[i][m];
}
// the reference's batch caller shape (main.cpp:1574-1592): nThreads host threads place the same sample concurrently through the adapter
int dropin_place_concurrent(void* h, int nThreads, const char* r1, const char* r2, const char* outPrefix, uint32_t* bestLogContainment) {
    auto* L = static_cast<Loaded*>(h);
    std::vector<std::thread> th; std::vector<int> ok((size_t)nThreads, 0);
    for (int t = 0; t < nThreads; ++t)
        th.emplace_back([&, t] {
            try {
                placement::PlacementResult R; placement::TraversalParams tp; tp.seedMaskFraction = 0.0;
                std::string o = std::string(outPrefix) + std::to_string(t) + ".tsv";
                placement::placeLite(R, &L->tree, *L->rd, r1, r2 ? r2 : "", o, tp, nullptr);
                bestLogContainment[t] = R.bestLogContainmentNodeIndex; ok[t] = 1;
            } catch (...) { ok[t] = 0; }
        });
    for (auto& x : th) x.join();
    int all = 1; for (int v : ok) all &= v;
    return all;
}

}  // extern "C"
