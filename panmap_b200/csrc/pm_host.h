// pm_host.h -- host-side structures shared by the .idx reader, the index flattener and the C ABI.
#pragma once
#include "../../include/panmap_b200.h"
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace pm {

struct IoError : std::runtime_error { using std::runtime_error::runtime_error; };
struct Unsupported : std::runtime_error { using std::runtime_error::runtime_error; };

// a parsed .idx (reference-native widths)
struct HostIndex {
    std::vector<uint8_t> raw;
    pm_seed_params sp{};
    std::vector<uint64_t> hash;
    std::vector<int16_t> parentCount, childCount;
    std::vector<uint64_t> nodeOffsets;
    std::vector<uint32_t> parentIndex;
    std::vector<std::string> nodeIds;
};
void readIdxFile(const std::string& path, HostIndex& out);


// everything pm_index_create derives from a pm_index_desc before uploading (see DESIGN.md "HBM layout")
struct FlatIndex {
    uint64_t N = 0, D = 0, S = 0;
    pm_seed_params sp{};
    uint32_t nodeBegin = 0, nodeEnd = 0, nAnc = 0, nLocal = 0;
    uint64_t nLocalDeltas = 0;
    // global tree arrays
    std::vector<uint32_t> parent, depth, subEnd, bfsRank;
    std::vector<uint8_t> isLeaf;
    std::vector<double> gMagSq, gMag;
    std::vector<int64_t> gUnique;
    std::vector<uint32_t> closeOff, closeList;
    // dictionary (id -> hash) and open-addressing table (hash -> id)
    std::vector<uint64_t> dictHash, dictKeys;
    std::vector<uint32_t> dictVals;
    uint64_t dictMask = 0;
    // shard-local delta storage
    std::vector<uint32_t> seedId, pc, lNode;
    std::vector<uint64_t> lOff;
    // delta kernel schedule: fixed chunks of 512 deltas (one warp each)
    uint64_t nDeltaChunks = 0;                 // seedId / pc are padded to nDeltaChunks * 512 entries
    std::vector<uint32_t> chunkNode;           // [nDeltaChunks+1] local node owning the chunk's first delta
    std::vector<uint8_t> isBoundary;           // [nLocal] node has deltas in more than one chunk (its sums are combined with atomics)
    std::vector<uint32_t> boundaryNodes;       // global ids of those nodes (their accumulators are zeroed per sample)
    // K2 tiles
    std::vector<uint32_t> carrySlot, chainOff, chainNodes;
    uint32_t nK2Tiles = 0;
    // selection
    std::vector<uint32_t> bfsNodes, bfsRanks;
    uint64_t rootDBegin = 0; uint32_t rootDCount = 0;
    uint64_t homo[4] = {0, 0, 0, 0};
};
// shard `shard` of `nShards` (contiguous DFS ranges balanced by delta count); throws std::runtime_error on bad input
void flattenIndex(const pm_index_desc& d, uint32_t shard, uint32_t nShards, FlatIndex& out);

void packBlockFirst(const uint64_t* packedOff, uint64_t nReads, uint64_t nChunks, uint32_t* out);

// reference BFS visit order (children ascending, level by level): rank of every node
void bfsRanks(const uint32_t* parent, uint64_t N, std::vector<uint32_t>& rank);

}  // namespace pm
