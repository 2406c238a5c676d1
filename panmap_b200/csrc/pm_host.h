// pm_host.h -- host-side structures shared by the .idx reader, the index flattener and the C ABI.
#pragma once
#include "../../include/panmap_b200.h"
#include <cstdint>
#include <stdexcept>
#include <functional>
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
    // carried through unchanged for consumers other than placement (mgsr / genotyping read them): LiteNode.identicalToParent,
    // LiteTree.blockRanges as (begin, end) pairs, the 4x4 substitution matrix
    std::vector<uint8_t> identicalToParent;
    std::vector<uint32_t> blockRanges;
    std::vector<double> substitutionMatrix;
};
void readIdxFile(const std::string& path, HostIndex& out);

// what a writer may add to the flat view of an index (every pointer may be null)
struct IdxExtras {
    const char* const* nodeIds = nullptr;          // [n_nodes]; "node_<i>" otherwise
    const uint8_t* identicalToParent = nullptr;    // [n_nodes]
    const uint32_t* blockRanges = nullptr; uint64_t nBlocks = 0;   // [2 * nBlocks]
    const double* substitutionMatrix = nullptr;    // [16]
};
// `.idx` container as the reference writes it (index_single_mode.cpp:1593-1636): 32-byte PMI1 header + the LiteIndex message, raw
// (zstdLevel < 0) or as independent 64 MB zstd frames.  Returns the number of bytes written.
uint64_t writeIdxFile(const std::string& path, const pm_index_desc& d, const IdxExtras& x, int zstdLevel);


// ---- `.panman` input of the index builder (pm_panman.cpp) ----
constexpr uint32_t kNoNode = 0xFFFFFFFFu;
struct PanmanNucMut { int32_t block, pos, gap; uint8_t len, type; uint32_t nucs; };   // gap == -1: main positions pos .. pos + len - 1
struct PanmanBlockMut { int32_t block; bool insertion, inversion; };
struct PanmanGapList { int32_t block; std::vector<int32_t> position, length; };
struct PanmanNode { std::string id; uint32_t parent = kNoNode; std::vector<uint32_t> children; uint32_t nucBegin = 0, nucEnd = 0, blockBegin = 0, blockEnd = 0; };
struct PanmanTree {
    std::vector<PanmanNode> nodes;            // newick pre-order == the DFS order of the index
    std::vector<PanmanNucMut> nucMuts;        // per node: [nucBegin, nucEnd)
    std::vector<PanmanBlockMut> blockMuts;    // per node: [blockBegin, blockEnd)
    std::vector<std::string> blocks;          // consensus of block b as characters
    std::vector<PanmanGapList> gaps;
};
void readPanman(const std::string& path, PanmanTree& out);
// one depth-first walk over the tree: visit(node, ungapped genome of the node) for every node, in pre-order
void walkPanmanGenomes(const PanmanTree& T, const std::function<void(uint32_t, const std::string&)>& visit);
struct PanmanFlat {   // the tree as the device pipeline of the builder takes it (pm_kernels.cuh BuildTreeView)
    std::vector<uint32_t> parent, slotBlock, blockStart, editBegin, editSlot, blockMutBegin, blockMut;
    std::string tmpl, editChar;
    std::vector<uint8_t> editSerial;
    uint32_t maxDepth = 1;
};
void flattenPanman(const PanmanTree& T, PanmanFlat& out);
void walkPanmanGenomesCoords(const PanmanTree& T, bool wantCoords, const std::function<void(uint32_t, const std::string&, const std::vector<uint32_t>&)>& visit);

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

// ---- cached image of a flattened index (pm_image.cpp) ----
struct ImageStamp {            // identifies what an image was flattened from
    uint64_t srcSize = 0, srcMtimeNs = 0;
    uint8_t srcHeader[32] = {0};   // the source's PMI1 header (seeding parameters, compression flag)
    uint32_t shard = 0, nShards = 1;
};
ImageStamp stampOfFile(const std::string& idxPath, uint32_t shard, uint32_t nShards);
uint64_t writeFlatImage(const std::string& path, FlatIndex& F, const std::vector<std::string>& nodeIds, const ImageStamp& stamp);
// false (with the reason in *why) when there is no usable image: missing, other format version, stamp mismatch, damaged
bool readFlatImage(const std::string& path, FlatIndex& F, std::vector<std::string>& nodeIds, const ImageStamp* expect, std::string* why);

void packBlockFirst(const uint64_t* packedOff, uint64_t nReads, uint64_t nChunks, uint32_t* out);

// reference BFS visit order (children ascending, level by level): rank of every node
void bfsRanks(const uint32_t* parent, uint64_t N, std::vector<uint32_t>& rank);

}  // namespace pm
