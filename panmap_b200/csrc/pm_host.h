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
    // dictionary (id -> hash) and open-addressing table (hash -> id)
    std::vector<uint64_t> dictHash, dictKeys;
    std::vector<uint32_t> dictVals;
    uint64_t dictMask = 0;
    // shard-local delta storage (local nodes = ancestors of nodeBegin, root first, then the shard's nodes)
    std::vector<uint32_t> lNode;               // [nLocal] global node id
    // "fast" deltas = genome count 0 <-> 1 (all but a handful): one 32-bit word each, in node order.
    //   word = 2 * seed id + (seed lost: parent 1 -> child 0); the last fast delta of a node ends a "segment"
    // padded to whole 512-word chunks with words that gather the always-zero slot of seed id S
    std::vector<uint32_t> dw;
    std::vector<uint32_t> endMask;             // [nDeltaChunks*32] bit j: word 16*lane + j of the chunk ends a segment
    uint64_t nFast = 0, nDeltaChunks = 0;
    uint32_t nSeg = 0;                         // nodes with at least one fast delta, in local order
    std::vector<uint32_t> chunkSeg;            // [nDeltaChunks+1] segments ending before the chunk | bit 31: chunk starts inside a segment
    std::vector<uint32_t> nodeSeg;             // [N] segment of a (local) node, NONE otherwise
    std::vector<uint32_t> boundarySegs;        // segments that span chunks: combined with atomics, zeroed per sample
    // "general" deltas (a genome count >= 2 on either side; rare): side list + DFS-interval events for their tree prefix
    std::vector<uint32_t> genSlot, genId, genPc;   // per general delta: slot of its node, seed id, parent | child << 16
    uint32_t nGenNodes = 0;
    std::vector<uint32_t> evSlot;              // events sorted by DFS position: slot | bit 31 = subtract (subtree of the node ended)
    std::vector<uint32_t> evIdx;               // [N] number of events at positions <= w (empty when there are no general deltas)
    // root's deltas (weighted-containment denominator)
    std::vector<uint32_t> rootId, rootChild;
    // K2 tiles
    std::vector<uint32_t> carrySlot, chainOff, chainNodes;
    uint32_t nK2Tiles = 0;
    // selection
    std::vector<uint32_t> bfsNodes, bfsRanks;
    uint64_t homo[4] = {0, 0, 0, 0};
};
// shard `shard` of `nShards` (contiguous DFS ranges balanced by delta count); throws std::runtime_error on bad input
void flattenIndex(const pm_index_desc& d, uint32_t shard, uint32_t nShards, FlatIndex& out);

void packBlockFirst(const uint64_t* packedOff, uint64_t nReads, uint64_t nChunks, uint32_t* out);

// reference BFS visit order (children ascending, level by level): rank of every node
void bfsRanks(const uint32_t* parent, uint64_t N, std::vector<uint32_t>& rank);

}  // namespace pm
