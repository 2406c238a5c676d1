// pm_kernels.cuh -- device-side data layout + kernel launch prototypes of the placement path.
// See DESIGN.md for the HBM layout and the roofline of each kernel.
#pragma once
#include "pm_logic.cuh"
#include <cuda_runtime.h>

namespace pm {

constexpr int kChunkDeltas = 512;   // K1: deltas per warp chunk (16 consecutive deltas per lane)
constexpr int kTileNodesK2 = 512;   // K2: nodes per prefix tile
constexpr int kBfsBlock = 1024;     // selection: BFS positions per block
constexpr int kLog1pLut = 1 << 16;  // log1p(count) table computed on the host with glibc (bit-identical terms)
constexpr u32 kNone = 0xFFFFFFFFu;

// per-node parent-relative sums written by K1 and read by K2: 9 u64 per node = raw, cos, wc, cont as fx128 (lo, hi) + presence
constexpr int kDeltaWords = 9;
struct __align__(16) TableSlot { u64 key; u32 count; u32 pad; };  // read seed table: one 16-byte slot = one 32-byte sector half
struct __align__(16) DictSlot { u64 key; u32 id; u32 pad; };       // index dictionary: seed hash -> dense seed id

struct Acc5 {  // exact accumulator of the 5 per-node numerators
    fx128 f[4];  // raw, cos, wc, cont
    i64 pres;
};

struct SampleAcc {  // device-side accumulators of one sample (zeroed per sample)
    u64 magSq[2], logSum[2], wcDen[2];  // fx128 as (lo, hi)
    long long kept, total, unique, multiSum, multiCount, entries, maxKeptCount, overflow, emptyKeyCount;
    unsigned touchedCount, pad0;
    unsigned recordCount[8];
    unsigned tieCount[8];
};
struct Selection {  // outcome of the tolerance chain for one metric
    double best;
    u32 bestNode;
    u32 lastRank;  // BFS rank of the last strict improvement, kNone if none
};

// ---- device view of the flattened index (immutable) ----
struct DevIndexView {
    u64 nNodes;        // global node count
    u32 nodeBegin, nodeEnd;  // shard extent (DFS indices)
    u32 nLocal;        // local nodes = ancestors of nodeBegin (ascending) ++ [nodeBegin, nodeEnd)
    u32 nAnc;
    u64 nLocalDeltas;
    u64 nSeeds;        // distinct seed hashes of the whole index
    const u32* seedId; // [nLocalDeltas] dense seed id (first-appearance order along the DFS)
    const u32* pc;     // [nLocalDeltas] parentCount (low 16) | childCount (high 16), int16 each
    const u64* lOff;   // [nLocal+1] delta offsets of local nodes
    const u32* lNode;  // [nLocal] global node id of a local node
    u64 nDeltaChunks; u64 nRealDeltas;
    const u32* chunkNode;            // [nDeltaChunks+1]
    const unsigned char* isBoundary; // [nLocal]
    const u32* boundaryNodes; u32 nBoundary;
    // tree (global arrays)
    const u32* parent;     // [nNodes]
    const double* gMag;    // [nNodes] sqrt(genomeMagnitudeSquared)
    const u32* closeOff;   // [nNodes+1] CSR: nodes whose subtree ends right before node w
    const u32* closeList;
    const u32* carrySlot;  // [nNodes] (shard-specific) chain position of the parent when it lies outside w's K2 tile
    const u32* chainOff;   // [nK2Tiles+1]
    const u32* chainNodes; // ancestors (root first) of each K2 tile's first node
    u32 nK2Tiles; u32 chainTotal;
    const unsigned char* isLeaf;  // [nNodes]
    // selection
    const u32* bfsNodes;   // [nShardNodes] shard nodes sorted by global BFS rank
    const u32* bfsRanks;   // [nShardNodes] their global BFS ranks
    u32 nShardNodes; u32 nBfsBlocks;
    // dictionary: seed hash -> seed id
    const DictSlot* dict; u64 dictMask;
    const u64* dictHash;   // [nSeeds] id -> hash
    // root's deltas (for the weighted-containment denominator): local range of global node 0
    u64 rootDBegin; u32 rootDCount; u32 hasRoot;
    const double* log1pLut;  // [kLog1pLut]
    const double* log1pSmall; // [32768] log1p(genome count)
    double ln2;            // log1p(1.0) from the host libm
};

struct WorkspaceView {
    // read table
    TableSlot* table; u64 tableMask; u64 tableCap;
    SampleAcc* acc;
    u64* synBuf; unsigned* synCount;  // per-read syncmer hashes (region of read r starts at 32*packedOff[r]) and counts
    double* ell;          // [nSeeds] log1p(read count) of seed id, 0 when absent
    u32* touched; u32 touchedCap;
    unsigned* countHist;  // [kLog1pLut] multiplicity of every read count among the kept seeds (rounding-drift model)
    u64* deltaFx;         // [nNodes][kDeltaWords]
    u64* chainA;          // [chainTotal][9]
    double* scores;       // [nNodes][5]
    double* metrics;      // [nNodes][5] or null
    double* blockMax;     // [nBfsBlocks][5]
    // records / ties per metric
    u32* recRank; u32* recNode; double* recScore; u32 recCap;  // [5][recCap]
    u32* tieNode; u32 tieCap;                                  // [5][tieCap]
    Selection* sel;       // [5]
    SampleScalars* scalars;
};

struct PlaceOpts {
    int minReadSupport;
    int forceLeaf;
    u32 skipNode;
    int wantMetrics;
};

// launches (all asynchronous on `st`)
void launchPackReads(const char* reads, const u64* off, const u64* packedOff, const u32* blockFirst, u64 nReads, u64 gBase, u64 nChunks,
                     uint4* packed, cudaStream_t st);
void launchSeedTable(const uint4* packed, const u64* off, const u64* packedOff, u64 nReads, const SeederParams& P,
                     const SeedTables* dTables, WorkspaceView W, cudaStream_t st);
void launchSeedList(const uint4* packed, const u64* off, const u64* packedOff, const u64* winOff, u64 nReads,
                    const SeederParams& P, const SeedTables* dTables, int mode, u64* synBuf, unsigned* synCount, u64* outHash,
                    unsigned char* outRev, long long* outPos, u64* outCount, cudaStream_t st);
void launchTableClear(WorkspaceView W, cudaStream_t st);
void launchTableImport(WorkspaceView W, const u64* hash, const long long* count, u64 n, cudaStream_t st);
void launchTableExport(WorkspaceView W, u64* hash, long long* count, unsigned* counter, u64 cap, cudaStream_t st);
void launchFinalize(DevIndexView I, WorkspaceView W, PlaceOpts O, const u64* homo, cudaStream_t st);
void launchDeltas(DevIndexView I, WorkspaceView W, int nSM, cudaStream_t st);
void launchPrefixScores(DevIndexView I, WorkspaceView W, PlaceOpts O, cudaStream_t st);
void launchRecords(DevIndexView I, WorkspaceView W, PlaceOpts O, cudaStream_t st);
void launchChain(WorkspaceView W, const u32* recCountOverride, cudaStream_t st);
void launchTies(DevIndexView I, WorkspaceView W, PlaceOpts O, cudaStream_t st);
void launchResetEll(WorkspaceView W, cudaStream_t st);

}  // namespace pm
