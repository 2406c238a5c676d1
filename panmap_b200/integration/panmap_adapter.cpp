// panmap_adapter.cpp -- the drop-in: placement::placeLite with the reference's OWN signature (/root/reference/src/placement.hpp:237-244)
// on top of the C ABI of libpanmap_b200.so.  This is the file a panmap maintainer adds to the build in place of the body of placeLite in
// src/placement.cpp (:986-2032); it compiles against panmap's own headers (placement.hpp, panmap_utils.hpp, the generated
// index_lite.capnp.h) and is NOT part of libpanmap_b200.so.  oracle/ref_build/Makefile builds it together with the unmodified reference
// translation units (the reference's own placeLite renamed on the compiler command line so that both can live in one binary) into
// oracle/_ref/libpanmap_dropin.so, and tests/test_gpu_dropin.py runs the reference's caller sequence (IndexReader -> LiteTree::initialize ->
// placeLite, src/main.cpp:1668-1750) through it and diffs every PlacementResult field and the TSV against the reference's own run.
//
// Ownership and threading follow the reference: the caller owns the index reader and the LiteTree; the first call for a tree
// flattens the index into HBM (cached per LiteTree, like `seedChangesLoaded`, panmap_utils.hpp:100); every calling thread gets its own
// pm_workspace (own CUDA stream), so the TBB workers of runBatchPlacement (main.cpp:1574-1592) call concurrently.
#include "placement.hpp"

#include "index_lite.capnp.h"
#include "panmap_utils.hpp"

#include <panmap_b200.h>

#include <capnp/any.h>

#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

struct DeviceSide {
    pm_index* index = nullptr;
    std::vector<std::string> ids;        // LiteNode ids by DFS index (TSV writer)
    std::vector<const char*> idPtrs;
    pm_seed_params seed{};
};
std::mutex g_mu;
std::map<const panmapUtils::LiteTree*, DeviceSide*> g_byTree;   // one device index per loaded LiteTree

int deviceFromEnv() {
    const char* e = std::getenv("PANMAP_B200_DEVICE");
    return e ? std::atoi(e) : 0;
}

// the zero-copy views placement.cpp:1021-1092 builds, concatenated over the 5e8-element segments and handed to the GPU once
DeviceSide* deviceSideOf(panmapUtils::LiteTree* tree, ::capnp::MessageReader& liteIndex) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_byTree.find(tree);
    if (it != g_byTree.end()) return it->second;
    auto root = liteIndex.getRoot<LiteIndex>();
    if (root.getFormatVersion() != panmapUtils::INDEX_FORMAT_VERSION)    // same check and wording as placement.cpp:1013-1019
        throw std::runtime_error("Index format version " + std::to_string(root.getFormatVersion()) + " is incompatible with this panmap (expects " +
                                 std::to_string(panmapUtils::INDEX_FORMAT_VERSION) + "). Rebuild the index (delete the .idx and rerun).");
    if (!root.hasSeedChangeHashes() || !root.hasSeedChangeParentCounts() || !root.hasSeedChangeChildCounts() || !root.hasNodeChangeOffsets())
        throw std::runtime_error("Index missing required V3 fields (seedChangeHashes, etc). V2 is no longer supported.");
    const size_t N = tree->dfsIndexToNode.size();
    auto offs = root.getNodeChangeOffsets();
    if (offs.size() != N + 1)
        throw std::runtime_error("Struct-of-arrays format offsets size mismatch: " + std::to_string(offs.size()) + " vs " + std::to_string(N + 1));
    std::vector<uint64_t> nodeOffsets(N + 1);
    for (size_t i = 0; i <= N; ++i) nodeOffsets[i] = offs[i];
    const uint64_t D = nodeOffsets[N];
    std::vector<uint64_t> hash(D); std::vector<int16_t> par(D), chi(D);
    uint64_t done = 0;
    auto hs = root.getSeedChangeHashes(); auto ps = root.getSeedChangeParentCounts(); auto cs = root.getSeedChangeChildCounts();
    for (uint32_t seg = 0; seg < hs.size(); ++seg) {
        auto h = hs[seg]; auto p = ps[seg]; auto c = cs[seg];
        const uint64_t n = std::min<uint64_t>(h.size(), D - done);
        for (uint64_t i = 0; i < n; ++i) { hash[done + i] = h[i]; par[done + i] = p[i]; chi[done + i] = c[i]; }
        done += n;
    }
    if (done != D) throw std::runtime_error("seed-change arrays shorter than nodeChangeOffsets says");
    std::vector<uint32_t> parentIdx(N, 0);
    for (size_t i = 1; i < N; ++i) parentIdx[i] = tree->dfsIndexToNode[i]->parent->nodeIndex;
    std::unique_ptr<DeviceSide> ds(new DeviceSide());
    ds->seed = pm_seed_params{root.getK(), root.getS(), root.getT(), root.getL(), root.getOpen() ? 1 : 0, root.getHpc() ? 1 : 0};
    pm_index_desc d{N, D, hash.data(), par.data(), chi.data(), nodeOffsets.data(), parentIdx.data(), ds->seed};
    if (pm_index_create(&d, deviceFromEnv(), &ds->index) != PM_OK) throw std::runtime_error(pm_last_error());
    ds->ids.resize(N);
    for (size_t i = 0; i < N; ++i) ds->ids[i] = tree->resolveNodeId(static_cast<uint32_t>(i));
    for (auto& s : ds->ids) ds->idPtrs.push_back(s.c_str());
    DeviceSide* raw = ds.release();
    g_byTree[tree] = raw;
    return raw;
}

pm_workspace* workspaceOf(DeviceSide* ds) {
    thread_local std::map<pm_index*, pm_workspace*> mine;    // one stream per calling thread and index
    auto it = mine.find(ds->index);
    if (it != mine.end()) return it->second;
    pm_workspace* ws = nullptr;
    if (pm_workspace_create(ds->index, &ws) != PM_OK) throw std::runtime_error(pm_last_error());
    mine[ds->index] = ws;
    return ws;
}

}  // namespace

namespace placement {

void placeLite(PlacementResult& result, panmapUtils::LiteTree* liteTree, ::capnp::MessageReader& liteIndex, const std::string& reads1,
               const std::string& reads2, std::string& outputPath, const TraversalParams& params, panmanUtils::Tree* /*fullTree*/) {
    if (params.verify_scores) throw std::runtime_error("VERIFICATION MODE requires the CPU path (full tree): not available on the GPU place stage");
    DeviceSide* ds = deviceSideOf(liteTree, liteIndex);
    pm_workspace* ws = workspaceOf(ds);

    pm_place_params p{};
    p.trim_start = params.trimStart; p.trim_end = params.trimEnd; p.min_read_support = params.minReadSupport;
    p.dedup_reads = params.dedupReads ? 1 : 0; p.force_leaf = params.forceLeaf ? 1 : 0; p.skip_node_index = PM_NONE;
    p.seed_mask_fraction = params.seedMaskFraction;
    p.want_node_scores = (params.store_diagnostics || params.refineEnabled) ? 1 : 0;
    p.min_seed_quality = params.minSeedQuality;
    pm_place_result r{};
    char err[512] = {0};
    // files in, result + <outputPath> TSV out (extractReadSequences / extractFullFastqData, the place stage, the TSV writer of :1952-2006)
    const int rc = pm_place_files(ds->index, ws, ds->idPtrs.data(), ds->idPtrs.size(), reads1.c_str(), reads2.c_str(), outputPath.c_str(), &p, &r,
                                  err, sizeof(err));
    if (rc != PM_OK) throw std::runtime_error(err[0] ? err : pm_last_error());

    double* sc[5] = {&result.bestLogRawScore, &result.bestLogCosineScore, &result.bestContainmentScore, &result.bestWeightedContainmentScore,
                     &result.bestLogContainmentScore};
    uint32_t* ix[5] = {&result.bestLogRawNodeIndex, &result.bestLogCosineNodeIndex, &result.bestContainmentNodeIndex,
                       &result.bestWeightedContainmentNodeIndex, &result.bestLogContainmentNodeIndex};
    std::vector<uint32_t>* td[5] = {&result.tiedLogRawNodeIndices, &result.tiedLogCosineNodeIndices, &result.tiedContainmentNodeIndices,
                                    &result.tiedWeightedContainmentNodeIndices, &result.tiedLogContainmentNodeIndices};
    for (int m = 0; m < 5; ++m) {
        *sc[m] = r.best_score[m]; *ix[m] = r.best_index[m];
        td[m]->assign(r.tied_count[m], 0);
        if (r.tied_count[m] && pm_get_tied(ws, m, td[m]->data(), r.tied_count[m]) != PM_OK) throw std::runtime_error(pm_last_error());
    }
    result.resolveNodeIds(liteTree);   // unchanged reference code (placement.cpp:403-437)
    if (p.want_node_scores) {          // per-node scores for --dump-all-scores / refinement (placement.cpp:812-818: floats)
        const size_t N = liteTree->dfsIndexToNode.size();
        std::vector<double> flat(N * 5);
        if (pm_get_node_scores(ws, flat.data()) != PM_OK) throw std::runtime_error(pm_last_error());
        result.nodeScores.resize(N);
        for (size_t v = 0; v < N; ++v) for (int m = 0; m < 5; ++m) result.nodeScores[v][m] = static_cast<float>(flat[v * 5 + m]);
    }
    result.totalReadsProcessed = static_cast<int64_t>(r.total_reads);
    result.reads1Path = reads1; result.reads2Path = reads2;
    {   // the read seed table the alignment stage filters reference seeds with (placement.hpp:223)
        std::vector<uint64_t> h(r.unique_seeds + 1); std::vector<int64_t> c(r.unique_seeds + 1);
        const int n = pm_get_seed_table(ws, h.data(), c.data(), h.size());
        if (n < 0) throw std::runtime_error(pm_last_error());
        result.seedFreqInReads.clear();
        result.seedFreqInReads.reserve(static_cast<size_t>(n));
        for (int i = 0; i < n; ++i) result.seedFreqInReads[static_cast<size_t>(h[i])] = c[i];
    }
    result.k = ds->seed.k; result.s = ds->seed.s; result.t = ds->seed.t; result.open = ds->seed.open != 0;
    result.readUniqueSeedCount = r.read_unique_seed_count;
    result.totalReadSeedFrequency = r.total_read_seed_frequency;
    result.readMagnitude = r.read_magnitude;
}

}  // namespace placement
