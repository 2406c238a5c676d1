// placement.hpp -- host-side mirror of the reference's placement interface (/root/reference/src/placement.hpp) on top
// of the C ABI in include/panmap_b200.h.  Same names, same argument meaning, same error behaviour:
//   * TraversalParams / PlacementResult carry the fields of placement.hpp:28-54 and :157-235 that the place stage uses
//   * placeLite() == placement::placeLite (placement.hpp:237-244): reads FASTA/FASTQ(.gz) files (R1 then R2, pairs interleaved,
//     placement.cpp:164-197), places them, fills the result and writes <outputPath> exactly as placement.cpp:1952-1985 does
//   * errors are std::runtime_error, like the reference's (placement.cpp:1013-1046)
// The reference passes (LiteTree*, capnp::MessageReader&); inside panmap those are turned into a pm_index once (see
// INTEGRATION.md for the adapter).  Here the index is the already-created device index plus the node id table.
#pragma once
#include "../../include/panmap_b200.h"

#include <cstdint>
#include <string>
#include <vector>

namespace placement {

struct TraversalParams {  // placement.hpp:28-54 (k, s, t, l, open, hpc come from the index, placement.cpp:1094-1101)
    double seedMaskFraction = 0.001;  // struct default of the reference; the CLI passes 0 (main.cpp:1967)
    int minSeedQuality = 0;
    bool dedupReads = false;
    int trimStart = 0;
    int trimEnd = 0;
    int minReadSupport = -1;
    bool store_diagnostics = false;   // keep per-node scores (--dump-all-scores)
    bool forceLeaf = false;
};

struct PlacementResult {  // placement.hpp:157-235
    double bestLogRawScore = 0.0;                 uint32_t bestLogRawNodeIndex = UINT32_MAX;                 std::vector<uint32_t> tiedLogRawNodeIndices;
    double bestLogCosineScore = 0.0;              uint32_t bestLogCosineNodeIndex = UINT32_MAX;              std::vector<uint32_t> tiedLogCosineNodeIndices;
    double bestContainmentScore = 0.0;            uint32_t bestContainmentNodeIndex = UINT32_MAX;            std::vector<uint32_t> tiedContainmentNodeIndices;
    double bestWeightedContainmentScore = 0.0;    uint32_t bestWeightedContainmentNodeIndex = UINT32_MAX;    std::vector<uint32_t> tiedWeightedContainmentNodeIndices;
    double bestLogContainmentScore = 0.0;         uint32_t bestLogContainmentNodeIndex = UINT32_MAX;         std::vector<uint32_t> tiedLogContainmentNodeIndices;
    std::vector<std::vector<double>> nodeScores;  // [n_nodes][5] when store_diagnostics (f64 here, float in the reference)
    std::string bestLogRawNodeId, bestLogCosineNodeId, bestContainmentNodeId, bestWeightedContainmentNodeId, bestLogContainmentNodeId;
    int64_t totalReadsProcessed = 0;
    std::string reads1Path, reads2Path;
    size_t readUniqueSeedCount = 0;
    int64_t totalReadSeedFrequency = 0;
    double readMagnitude = 0.0;
    pm_place_result raw{};                        // everything the C ABI returned for this sample (tie counts, seed statistics, stage times)
};

// the device index + what LiteTree::resolveNodeId needs (panmap_utils.hpp:113-118)
struct DeviceIndex {
    pm_index* index = nullptr;
    pm_workspace* workspace = nullptr;           // one per concurrent caller (batch threads own one each)
    const std::vector<std::string>* nodeIds = nullptr;
};

void placeLite(PlacementResult& result, DeviceIndex& index, const std::string& reads1, const std::string& reads2,
               std::string& outputPath, const TraversalParams& params = {});

// extractReadSequences (placement.cpp:164-197): sequences of reads1 then reads2, pairs interleaved; throws on pair-count
// mismatch (the reference prints the message and exit(1)s)
// quals (optional) == extractFullFastqData (placement.cpp:199-238): the quality bytes at the same offsets, 'I' for records without
void extractReadSequences(const std::string& readPath1, const std::string& readPath2, std::string& bases, std::vector<uint64_t>& offsets,
                          std::string* quals = nullptr);

}  // namespace placement
