// pm_kernels.cuh -- device-side data layout + kernel launch prototypes of the placement path.
// See DESIGN.md for the HBM layout and the roofline of each kernel.
#pragma once
#include "pm_logic.cuh"
#include <cuda_runtime.h>

namespace pm {

constexpr int kChunkWords = 512;    // K1: packed delta words per warp chunk (16 consecutive words per lane)
constexpr int kBfsBlock = 1024;     // selection: BFS positions per block
constexpr int kLog1pLut = 1 << 16;  // log1p(count) table computed on the host with glibc (bit-identical terms)
constexpr u32 kNone = 0xFFFFFFFFu;
constexpr int kTieHead = 64;        // tied nodes per metric that travel with the small result block

// Per-segment (= node with 0<->1 deltas) parent-relative sums written by K1 and read by K2, 16 bytes:
//   sum = hi * 2^64 + lo, a signed 96-bit integer in units of 2^-53: the exact sum of +-log1p(readCount) over the node's seeds
//   that are in the reads (log1p(count >= 1) >= ln 2, so every addend is a multiple of 2^-53);  cnt = #gained - #lost among those.
// With genome counts 0 <-> 1 the reference's five numerators are functions of these two (placement.cpp:315-339):
//   logRaw = logContainment = sum, logCosine = sum * log1p(1), weightedContainment = presence = cnt.
struct __align__(16) SegRec { u64 lo; int hi; int cnt; };
constexpr double kEllScale = 9007199254740992.0;           // 2^53
constexpr double kEllInvScale = 1.0 / 9007199254740992.0;  // 2^-53
constexpr int kGenWords = 9;   // general-delta accumulators per node: raw, cos, wc, cont as fx128 (lo, hi) + presence
struct __align__(16) TableSlot { u64 key; u32 count; u32 pad; };  // read seed table: one 16-byte slot = one 32-byte sector half
struct __align__(16) DictSlot { u64 key; u32 id; u32 pad; };       // index dictionary: seed hash -> dense seed id

struct SampleAcc {  // device-side accumulators of one sample (zeroed per sample)
    u64 magSq[2], logSum[2], wcDen[2];  // fx128 as (lo, hi)
    long long kept, total, unique, multiSum, multiCount, entries, maxKeptCount, overflow, emptyKeyCount;
    unsigned missCount, entCount;   // missCount: entries of the miss queue (count_seeds_lane -> count_misses)
    unsigned recordCount[8];
    unsigned tieCount[8];
    long long n1NotIndex;   // sharded samples: kept seeds with read count 1 that the index does not hold (they travel as a count, not as entries)
    long long totalReads;   // sharded samples: reads of all ranks
    unsigned finDone, pad1; // root_and_scalars: blocks that finished the root pass (the last one computes the scalars)
};
// bits of SampleAcc::overflow (any non-zero value makes the host grow what was too small and redo the sample)
constexpr long long kOvfTable = 1, kOvfPair = 2, kOvfGather = 4, kOvfRecords = 8;
constexpr long long kOvfPeer = 16;   // not a capacity: a peer's data did not arrive in time (peer-memory transport); never retried
struct ScanPartial { long long multiSum, multiCount, entries, total; };          // one per table_scan block
struct FinPartial { u64 mag[2], lsum[2]; long long kept; long long maxc; };        // one per entries_finalize block
constexpr unsigned kMaxPartials = 2048;
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
    const u32* dw;         // [nDeltaChunks*512] packed fast deltas: 2 * seed id + lost; inside a chunk the 16-byte piece q of lane l
                           // (logical words 16 l + 4 q ..) is stored at uint4 index 32 q + l
    const u32* endMask;    // [nDeltaChunks*32] per lane (16 words): bit j = word j is the last fast delta of its node
    u64 nDeltaChunks;
    const u32* chunkSeg;   // [nDeltaChunks+1] segments ending before the chunk | bit 31: the chunk starts inside a segment
    const u32* nodeSeg;    // [nNodes] segment of a node, kNone when it has no fast deltas (or is not local)
    const u32* boundarySegs; u32 nBoundary; u32 nSeg;
    // general deltas (genome count >= 2 on a side)
    const u32* genSlot; const u32* genId; const u32* genPc; u32 nGenDeltas; u32 nGenNodes;
    const u32* evSlot; u32 nEvents;   // DFS-interval events of the nodes with general deltas
    const u32* evIdx;      // [nNodes] events at positions <= w (null when nGenNodes == 0)
    // tree (global arrays)
    const u32* parent;     // [nNodes]
    const double* gMag;    // [nNodes] sqrt(genomeMagnitudeSquared)
    const u32* subEnd;     // [nNodes] one past the last DFS index of the node's subtree
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
    // root's deltas (for the weighted-containment denominator)
    const u32* rootId; const u32* rootChild; u32 rootDCount; u32 hasRoot;
    const double* log1pLut;  // [kLog1pLut]
    const double* log1pSmall; // [32768] log1p(genome count)
    double ln2;            // log1p(1.0) from the host libm
};

struct WorkspaceView {
    // read table
    TableSlot* table; u64 tableMask; u64 tableCap;
    cudaTextureObject_t tableTex;   // the table as a linear uint4 texture (first probe of an insertion)
    SampleAcc* acc;
    u64* synBuf; unsigned* synCount;  // per-read syncmer hashes (region of read r starts at 32*packedOff[r]) and counts
    u64* missQ; u64 missCap;          // seeds that missed the per-SM tables of count_seeds_lane, waiting for count_misses
    // partitioned counting (tables larger than L2): seed instances scattered by table region, counted region by region (bktCount == 0: off)
    u64* bktBuf; u32* bktFill;        // [kBktWarps * bktCount][bktRegionCap] seeds and their fill counts, one region per (warp, bucket)
    u32 bktCount, bktShift, bktRegionCap;   // buckets (power of two), bucket = home slot >> bktShift
    cudaTextureObject_t ellTex;   // ell as a linear int2 texture: scattered gathers go through the TEX data path instead of the LSU's
    long long* ell;       // [nSeeds+2] log1p(read count) * 2^53 of the seed id (an exact integer), 0 when absent; slot nSeeds stays 0
    u64* entKey; u32* entCnt; u32* entId;   // [tableCap] occupied (key, count) pairs compacted by table_scan + the seed id found for them
    ScanPartial* scanPart; FinPartial* finPart;   // [kMaxPartials] per-block partials of the two finalize passes
    unsigned* countHist;  // [kLog1pLut] multiplicity of every read count among the kept seeds (rounding-drift model)
    SegRec* segRec;       // [nSeg]
    SegRec* chainA;       // [chainTotal] prefix along each K2 tile's ancestor chain
    u64* genRec;          // [nGenNodes][kGenWords]
    u64* evPrefix;        // [nEvents][kGenWords] inclusive prefix of the events
    double* scores;       // [nNodes][5]
    double* metrics;      // [nNodes][5] or null
    double* blockMax;     // [nBfsBlocks][5]
    // records / ties per metric
    u32* recRank; u32* recNode; double* recScore; u32 recCap;  // [5][recCap]
    u32* tieNode; u32 tieCap;                                  // [5][tieCap]
    u32* tieHead;                                              // [5][kTieHead] copy of the first entries (inside the result block)
    Selection* sel;       // [5]
    SampleScalars* scalars;
};

// ---- one sample over several GPUs: fixed-capacity exchange buffers with in-band counts (no host round trip between stages) ----
// seed partition exchange (all-to-all): per destination rank one segment of 1 + capPair 16-byte slots; slot 0 is the header
struct __align__(16) XHeader { u32 count; u32 flags; u32 pad[2]; };
struct __align__(16) XEntry { u64 key; u32 count; u32 pad; };
// finalized partition entries (all-gather): header + capG pairs (read count, seed id or kNone)
struct __align__(16) GHeader {
    u32 nEntries; u32 flags;            // flags: SampleAcc::overflow of the sending rank at that point
    u64 n1NotIndex;                     // seeds with count 1 that are not in the index (sent as a count only)
    long long multiSum, multiCount;     // auto min-support statistics of the partition (placement.cpp:931-955)
    long long unique, total;            // seeds / seed instances of the partition (after homopolymer removal)
    u64 nReads;                         // reads this rank seeded
    u32 maxPairCount; u32 localEntries; // sizing feedback: largest per-destination export count, unique seeds of the local table
};
static_assert(sizeof(GHeader) == 64, "GHeader is eight uint2 slots");
constexpr u32 kGHeaderSlots = sizeof(GHeader) / sizeof(uint2);
// local prefix-maximum records (all-gather): header (2 slots) + [5][recX] records of 16 bytes; recX is a run-time capacity
struct __align__(16) RHeader { u32 count[5]; u32 flags; u32 pad[2]; };
struct __align__(16) RRecord { double score; u32 rank; u32 node; };
// tie heads (all-gather): counts, flags, sizing feedback for the next sample, then [5][kTieHead] node ids
struct __align__(16) THeader { u32 tieCount[5]; u32 flags; u32 maxPairCount, localEntries, gEntries, partEntries; u32 pad[6]; };
static_assert(sizeof(THeader) == 64, "THeader is sixteen words");
constexpr u32 kTWords = sizeof(THeader) / 4 + 5 * kTieHead;   // 32-bit words per rank

struct PlaceOpts {
    int minReadSupport;
    int forceLeaf;
    u32 skipNode;
    int wantMetrics;
};

// every kernel launch of the library passes through here (pm_launch_count() reports the total: the bench's "gpu_launches")
void noteLaunch();

// launches (all asynchronous on `st`)
// endOff (optional, everywhere below): end of read r when it is shorter than off[r+1] - off[r] (homopolymer-compressed in place)
void launchPackReads(const char* reads, const u64* off, const u64* packedOff, const u32* blockFirst, u64 nReads, u64 gBase, u64 nChunks,
                     uint4* packed, cudaStream_t st, const u64* endOff = nullptr);
void launchHpcCompress(char* reads, const u64* off, u64 nReads, u64* endOff, cudaStream_t st, char* quals = nullptr);
void launchChunkOffsets(const u64* off, u64 n, u64 gBase, u64* tileSum /* [n/4096 + 1] scratch */, u64* packedOff, cudaStream_t st);
void launchSeedTable(const uint4* packed, const u64* off, const u64* packedOff, u64 nReads, const SeederParams& P,
                     const SeedTables* dTables, WorkspaceView W, cudaStream_t st, cudaEvent_t between = nullptr,
                     const unsigned char* dup = nullptr, const u64* endOff = nullptr, const char* reads = nullptr, cudaEvent_t tableReady = nullptr);
// --min-seed-quality > 0: needs the packed reads; quals = one byte per base at the reads' offsets, synPass = one byte per synBuf entry
void launchSeedTableQuality(const uint4* packed, const u64* off, const u64* packedOff, u64 nReads, const SeederParams& P,
                            const SeedTables* dTables, WorkspaceView W, cudaStream_t st, const u64* endOff, const char* quals,
                            int minSeedQuality, unsigned char* synPass);
// partitioned counting: the grid both kernels run with (one region per resident warp and bucket), whether (k, l) has a scatter kernel, and the
// pass over the regions that launchSeedTable's scatter leaves behind (call once after the last slice of a sample)
constexpr u32 kBktBlocks = 148, kBktWarps = kBktBlocks * 32;
bool bucketCountingSupports(int k, int l);
void launchCountBuckets(WorkspaceView W, cudaStream_t st);
// true when launchSeedTable can hash these parameters straight from the ASCII reads (pass `reads`, skip pack_reads)
bool seedTableReadsAscii(const SeederParams& P);
// --dedup: dup[r] = 1 when a byte-identical read holds the set already; reads [rBegin, rEnd) of the sample, `off` = all offsets
void launchDedup(const char* reads, const u64* off, u64 rBegin, u64 rEnd, unsigned long long* slots, u64 mask, unsigned char* dup, cudaStream_t st,
                 const u64* endOff = nullptr);
void launchSeedList(const uint4* packed, const u64* off, const u64* packedOff, const u64* winOff, u64 nReads,
                    const SeederParams& P, const SeedTables* dTables, int mode, u64* synBuf, unsigned* synCount, u64* outHash,
                    unsigned char* outRev, long long* outPos, u64* outCount, cudaStream_t st);
void launchSeedListsEnd(const uint4* packed, const u64* off, const u64* endOff, const u64* packedOff, const u64* winOff, u64 nSeqs, const SeederParams& P,
                        const SeedTables* dTables, u64* synBuf, unsigned* synCount, u64* outHash, u64* outCount, cudaStream_t st);
// ---- index builder (pm_build_kernels.cu) ----
constexpr u32 kBuildNoNode = 0xFFFFFFFFu;
constexpr unsigned kBuildSortCap = 16384;   // seeds per genome the shared-memory sort holds (larger genomes take the host pipeline)
struct BuildTreeView {   // the PanMAN flattened for the device: aligned template, point edits and block mutations per node
    u32 nNodes, nBlocks, nSlots, maxDepth;
    const u32* parent;            // [nNodes], kBuildNoNode for the root
    const char* tmpl;             // [nSlots] the root template in aligned order (gap slots before their main position, '-' where empty)
    const u32* slotBlock;         // [nSlots]
    const u32* blockStart;        // [nBlocks + 1]
    const u32* editBegin;         // [nNodes + 1] into editSlot / editChar
    const u32* editSlot; const char* editChar;
    const unsigned char* editSerial;   // [nNodes] 1: a slot is written twice by this node, apply in order
    const u32* blockMutBegin;     // [nNodes + 1] into blockMut
    const u32* blockMut;          // block << 2 | inversion << 1 | insertion
};
struct BuildDiffArgs {
    u32 nodeBegin, nNodes;
    const u32* parent;
    const u64* const* listPtr;    // [all nodes] sorted seed list of the node (device pointers)
    const u64* listCount;         // [all nodes]
    void* scratch;                // gridDim.x * 2 * kBuildSortCap entries of 16 bytes
    unsigned long long* cursor;   // deltas reserved so far
    unsigned long long outCap;
    u64* outHash; short* outPc; short* outCc;
    unsigned long long* nodeOff; unsigned* nodeCnt;   // [nNodes] where the node's deltas went
};
void launchGenomeMaterialize(const BuildTreeView& T, u32 nodeBegin, u32 nNodes, u32* pathScratch, unsigned char* blkScratch, char* aligned, char* genomes,
                             u64 pitch, u64* endOff, cudaStream_t st);
void launchSeedsSort(const u64* in, const u64* winOff, const u64* count, const u64* arenaOff, u64* arena, u32 nLists, cudaStream_t st);
void launchNodeDiff(const BuildDiffArgs& A, unsigned grid, cudaStream_t st);
void launchTableClear(WorkspaceView W, cudaStream_t st);
void launchSampleBegin(WorkspaceView W, cudaStream_t st);   // table clear + the sample's accumulators zeroed, one launch
void launchTableImport(WorkspaceView W, const u64* hash, const long long* count, u64 n, cudaStream_t st);
void launchTableExport(WorkspaceView W, u64* hash, long long* count, unsigned* counter, u64 cap, cudaStream_t st);
// seedMaskFraction > 0 (off by default): synchronises the stream a few dozen times to find the cut; maskScratch = two device words
void launchFinalize(DevIndexView I, WorkspaceView W, PlaceOpts O, const u64* homo, u64 expectedEntries, int nSM, cudaStream_t st,
                    double seedMaskFraction = 0.0, unsigned long long* maskScratch = nullptr, cudaStream_t stSide = nullptr, cudaEvent_t evFork = nullptr,
                    cudaEvent_t evJoin = nullptr);
void launchDeltas(DevIndexView I, WorkspaceView W, int nSM, cudaStream_t st);
void launchGeneral(DevIndexView I, WorkspaceView W, cudaStream_t st);
void launchPrefixScores(DevIndexView I, WorkspaceView W, PlaceOpts O, cudaStream_t st);
void launchRecords(DevIndexView I, WorkspaceView W, PlaceOpts O, cudaStream_t st);
void launchChain(WorkspaceView W, const u32* recCountOverride, cudaStream_t st);
void launchTies(DevIndexView I, WorkspaceView W, PlaceOpts O, cudaStream_t st);
void launchResetSample(DevIndexView I, WorkspaceView W, cudaStream_t st);

// ---- sharded samples (pm_shard_kernels.cu) ----
// local table -> per-owner segments of xSend ([nRanks][1 + capPair] slots, headers zeroed by the launch)
void launchPartitionExport(WorkspaceView W, u32 nRanks, u32 capPair, uint4* xSend, u32* exportInfo /* device, [2]: max per-destination count, total */, cudaStream_t st);
// received segments -> this rank's partition table (cleared by the caller)
void launchPartitionImport(WorkspaceView W, const uint4* xRecv, u32 nRanks, u32 capPair, cudaStream_t st);
// after table_scan on the partition: dictionary look-up of every entry, header + (count, id) pairs into gSend ([kGHeaderSlots + capG] uint2)
void launchPartitionFinalize(DevIndexView I, WorkspaceView W, PlaceOpts O, const u64* homo, int nSM, uint2* gSend, u32 capG, u64 nLocalReads,
                             const u32* maxPairCount, u32 localEntriesHint, cudaStream_t st);
// all ranks' gathered lists -> ell, exact magnitude sums, histogram, scalars (replaces entries_finalize + finish_scalars of the one-GPU path)
void launchGatheredFinalize(DevIndexView I, WorkspaceView W, PlaceOpts O, const uint2* gRecv, u32 nRanks, u32 capG, int nSM, cudaStream_t st,
                            cudaStream_t stSide = nullptr, cudaEvent_t evFork = nullptr, cudaEvent_t evJoin = nullptr);
void launchRecordsPack(WorkspaceView W, uint4* rSend, u32 recX, cudaStream_t st);
void launchChainGathered(WorkspaceView W, const uint4* rRecv, u32 nRanks, u32 recX, cudaStream_t st);
void launchTiesPack(WorkspaceView W, u32* tSend, const uint2* gSend, const u32* exportInfo /* [2] from launchPartitionExport */, cudaStream_t st);
void launchTiesFullPack(WorkspaceView W, u32* out /* [5][capT] */, u32 capT, cudaStream_t st);
// ---- peer-memory transport of the exchanges (one process per GPU, every rank's receive buffers mapped into every other rank) ----
// push_segments stores this rank's payload straight into the peers' receive buffers over NVLink and then raises, in every peer,
// the flag word (exchange, this rank) to `epoch`; wait_flags spins (bounded) until the flags of all ranks carry `epoch`.
constexpr int kPeerMax = 32;
struct PushArgs {
    const unsigned char* src;      // send buffer
    size_t srcStride;              // all-to-all: the segment for rank q starts at src + q * srcStride; all-gather: 0
    unsigned char* dst[kPeerMax];  // receive buffer of rank q as mapped here (own rank: the local one)
    size_t dstOffset;              // where this rank's segment starts inside a receive buffer
    size_t segBytes;               // capacity of a segment
    u32* flag[kPeerMax];           // flag word (exchange, this rank) inside rank q's flag block
    u32 kind;                      // 0 seed segments (XHeader), 1 pairs (GHeader), 2 / 3 fixed size
    u32 capEntries;                // kind 0: capPair, kind 1: capG
    u32 n, epoch;
};
void launchPushSegments(const PushArgs& A, WorkspaceView W, cudaStream_t st);
void launchWaitFlags(const u32* flags /* [n] of one exchange */, u32 n, u32 epoch, WorkspaceView W, cudaStream_t st);
void launchResetGathered(DevIndexView I, WorkspaceView W, const uint2* gRecv, u32 nRanks, u32 capG, cudaStream_t st);
void launchHashSeq(const char* seqs, const u64* off, u64 n, u64* fwd, u64* rev, unsigned char* status, cudaStream_t st);
// pieces of launchFinalize the sharded path runs on their own
void launchTableScan(WorkspaceView W, const u64* homo, int nSM, unsigned* nPartsOut, cudaStream_t st);
void launchRootAndScalars(DevIndexView I, WorkspaceView W, PlaceOpts O, unsigned nFinParts, cudaStream_t st);

}  // namespace pm
