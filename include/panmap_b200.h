/*
 * panmap_b200.h -- C ABI of the B200-native placement hot path (drop-in for panmap's place stage).
 *
 * panmap itself has no plugin/FFI layer: the boundary of the path is two C++ headers,
 *   placement::placeLite(...)              /root/reference/src/placement.hpp:237-244
 *   seeding::rollingSyncmers(...)          /root/reference/src/seeding.hpp:126-127
 * plus the unit-test-pinned helpers hashSeq (seeding.hpp:123), NodeMetrics::computeChildMetrics
 * (placement.hpp:151-154), resolveMinReadSupport / computeReadSeedMagnitudes (placement.hpp:99-105).
 * Every entry point below names the reference interface it replaces.  panmap_b200/host/ holds C++ shims
 * with the reference's own names/signatures that call this ABI; INTEGRATION.md shows how a panmap
 * maintainer binds them.
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Functions return PM_OK (0) or a negative
 * pm_status; pm_last_error() gives the message for the calling thread.  There is NO CPU fallback: every
 * compute entry point fails with PM_ERR_NO_DEVICE when no CUDA device is usable.
 * Handles: pm_index is immutable after creation and may be shared by any number of pm_workspace objects;
 * each pm_workspace owns one CUDA stream plus per-sample scratch, so distinct workspaces may be driven
 * concurrently from distinct host threads (the reference's batch mode, main.cpp:1574-1592).
 */
#ifndef PANMAP_B200_H
#define PANMAP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM_ABI_VERSION 5
#define PM_NONE 0xFFFFFFFFu /* "no node" (reference: UINT32_MAX, placement.hpp:159) */
#define PM_NUM_METRICS 5    /* log_raw, log_cosine, containment, weighted_containment, log_containment */

typedef enum pm_status {
    PM_OK = 0,
    PM_ERR_INVALID = -1,     /* bad argument / malformed index (reference throws std::runtime_error) */
    PM_ERR_NO_DEVICE = -2,   /* CUDA device missing or unusable: there is no CPU fallback */
    PM_ERR_CUDA = -3,        /* CUDA runtime error */
    PM_ERR_IO = -4,          /* file could not be read */
    PM_ERR_UNSUPPORTED = -5, /* option not implemented on the GPU path yet */
    PM_ERR_CAPACITY = -6     /* internal capacity exceeded after retry */
} pm_status;

typedef struct pm_host_index pm_host_index; /* a parsed .idx on the host */
typedef struct pm_index pm_index;           /* flattened index resident in HBM on one device */
typedef struct pm_workspace pm_workspace;   /* per-sample state: stream, read table, scores */
typedef struct pm_comm pm_comm;             /* one rank of a group of GPUs that place ONE sample together */

/* Seeding parameters; authoritative source is the index (placement.cpp:1094-1101). */
typedef struct pm_seed_params {
    int32_t k, s, t, l;
    int32_t open; /* 0 = closed syncmers */
    int32_t hpc;  /* homopolymer-compressed index: pm_place* collapses the reads on the device (placement.cpp:1145-1165, seeding.cpp:286-306) */
} pm_seed_params;

/* Flat view of the per-node seed-delta index (index_lite.capnp LiteIndex: seedChangeHashes /
 * seedChangeParentCounts / seedChangeChildCounts / nodeChangeOffsets + LiteTree.liteNodes[].parentIndex;
 * the reference's own flat view is built at placement.cpp:1021-1092).  Nodes are in DFS pre-order with
 * parent_index[v] < v (panmap_utils.cpp:273-286); node 0 is the root. */
typedef struct pm_index_desc {
    uint64_t n_nodes;
    uint64_t n_deltas;
    const uint64_t* delta_hash;   /* [n_deltas] */
    const int16_t* delta_parent;  /* [n_deltas] */
    const int16_t* delta_child;   /* [n_deltas] */
    const uint64_t* node_offsets; /* [n_nodes+1] */
    const uint32_t* parent_index; /* [n_nodes], entry 0 ignored */
    pm_seed_params seed;
} pm_index_desc;

/* Per-call options == placement::TraversalParams (placement.hpp:28-54), hot-path subset. */
typedef struct pm_place_params {
    int32_t trim_start;        /* trimStart */
    int32_t trim_end;          /* trimEnd */
    int32_t min_read_support;  /* minReadSupport: -1 = auto (placement.cpp:931-955) */
    int32_t dedup_reads;       /* dedupReads: every distinct read string counts once (placement.cpp:1550-1620) */
    int32_t force_leaf;        /* forceLeaf: only leaves are eligible (placement.cpp:794-795) */
    uint32_t skip_node_index;  /* leave-one-out node, PM_NONE = none (placement.hpp:91) */
    double seed_mask_fraction; /* seedMaskFraction (placement.cpp:1748-1799); CLI default 0 (main.cpp:1967); ties at the cut: smaller hash first */
    int32_t want_node_scores;  /* keep the [n_nodes][5] f64 score matrix for pm_get_node_scores */
    int32_t min_seed_quality;  /* minSeedQuality (placement.hpp:37): min average Phred over a seed's k bases; 0 = off; needs pm_place_quality */
} pm_place_params;

/* == the scalar part of placement::PlacementResult (placement.hpp:157-235) */
typedef struct pm_place_result {
    double best_score[PM_NUM_METRICS];   /* best*Score */
    uint32_t best_index[PM_NUM_METRICS]; /* best*NodeIndex after finalizeTiedIndices (lowest DFS index of the ties) */
    uint64_t tied_count[PM_NUM_METRICS]; /* tied*NodeIndices.size() after sort+unique */
    uint64_t total_reads;                /* totalReadsProcessed */
    uint64_t unique_seeds;               /* seedFreqInReads.size() after homopolymer removal / masking */
    uint64_t read_unique_seed_count;     /* readUniqueSeedCount (U') */
    int64_t total_read_seed_frequency;   /* totalReadSeedFrequency */
    int64_t min_read_support;            /* resolved value */
    double read_magnitude;               /* logReadMagnitude */
    double log_containment_denominator;
    double weighted_containment_denominator;
    float stage_ms[8];                   /* device time: 0 h2d, 1 seeding+table, 2 finalize, 3 deltas, 4 prefix+scores,
                                            5 selection, 6 d2h, 7 total (CUDA events on the workspace stream) */
} pm_place_result;

const char* pm_last_error(void);
int pm_abi_version(void);
/* number of usable CUDA devices (0 when none; never fails) */
int pm_device_count(void);
/* kernels this library has launched in this process so far (every launch site counts itself; memsets and copies are not kernels) */
uint64_t pm_launch_count(void);

/* ---- .idx container (index_single_mode.cpp:1561-1636, main.cpp:193-236): 32-byte "PMI1" header + raw
 *      Cap'n Proto LiteIndex message, raw or as independent zstd frames (libzstd.so.1 is loaded at run time when needed). ---- */
int pm_host_index_read(const char* path, pm_host_index** out);
void pm_host_index_free(pm_host_index* h);
int pm_host_index_desc(const pm_host_index* h, pm_index_desc* out); /* pointers stay owned by h */
const char* pm_host_index_node_id(const pm_host_index* h, uint64_t node); /* LiteNode.id */
/* What a LiteIndex holds besides the seed deltas; placement reads none of it except the ids, other stages do (mgsr.cpp:436-471), so a
 * read -> write round trip carries it through.  Every pointer may be NULL. */
typedef struct pm_index_extras {
    const char* const* node_ids;        /* [n_nodes] LiteNode.id; "node_<i>" when NULL */
    const uint8_t* identical_to_parent; /* [n_nodes] LiteNode.identicalToParent */
    const uint32_t* block_ranges;       /* [2 * n_blocks] LiteTree.blockRanges as (rangeBeg, rangeEnd) */
    uint64_t n_blocks;
    const double* substitution_matrix;  /* [16] LiteIndex.substitutionMatrix */
} pm_index_extras;
int pm_host_index_extras(const pm_host_index* h, pm_index_extras* out); /* pointers stay owned by h */
/* IndexBuilder::writeIndex (index_single_mode.cpp:1593-1636): PMI1 header + LiteIndex message (formatVersion 4), uncompressed when
 * zstd_level < 0, else independent 64 MB zstd frames at that level.  The file is readable by the reference's IndexReader
 * (main.cpp:193-236) and by pm_host_index_read.  One Cap'n Proto segment: PM_ERR_UNSUPPORTED beyond 4 GB of message. */
int pm_host_index_write(const char* path, const pm_index_desc* desc, const pm_index_extras* extras, int zstd_level, uint64_t* bytes_out);

/* ---- device index (replaces LiteTree::initialize + the SoA hookup, placement.cpp:1021-1092).
 *      node_begin/node_end select a shard of whole tiles for multi-GPU runs (0,0 = all nodes). ---- */
int pm_index_create(const pm_index_desc* desc, int device, pm_index** out);
int pm_index_create_shard(const pm_index_desc* desc, int device, uint32_t shard, uint32_t n_shards, pm_index** out);
void pm_index_destroy(pm_index* idx);
/* Cached image of the flattened index: the arrays pm_index_create derives from a LiteIndex (packed delta words, segment masks, seed
 * dictionary, tile chains, BFS order), stored next to the .idx so that later runs skip the parse and the flattening.  Same rule as the
 * reference's own index cache (cachedIndexUsable, main.cpp:371-396): an image is used only if it was made from exactly this file
 * (size, mtime, PMI1 header with the seeding parameters) for this shard; anything else re-flattens and rewrites it.
 * image_path NULL = idx_path + ".pmflat" (+ ".<shard>of<n>" for shards); *cache_hit reports which way it went. */
int pm_index_open_cached(const char* idx_path, const char* image_path, int device, uint32_t shard, uint32_t n_shards, pm_index** out, int* cache_hit);
int pm_index_image_write(const pm_index_desc* desc, const char* const* node_ids, uint32_t shard, uint32_t n_shards, const char* image_path, uint64_t* bytes_out);
int pm_index_create_from_image(const char* image_path, int device, pm_index** out); /* no source check: the caller vouches for the image */
const char* pm_index_node_id(const pm_index* idx, uint64_t node); /* ids travel with indexes opened from a file or an image; "" otherwise */
uint64_t pm_index_num_nodes(const pm_index* idx);
uint64_t pm_index_num_deltas(const pm_index* idx);
uint64_t pm_index_num_distinct_seeds(const pm_index* idx);
/* shard extent in DFS node indices: [begin, end) */
int pm_index_shard_range(const pm_index* idx, uint64_t* node_begin, uint64_t* node_end);
/* sample-independent per-node accumulators (NodeMetrics::genomeMagnitudeSquared / genomeUniqueSeedCount,
 * placement.cpp:292-302), precomputed at creation; out arrays are [n_nodes] */
int pm_index_genome_metrics(const pm_index* idx, double* genome_mag_sq, int64_t* genome_unique);

int pm_workspace_create(pm_index* idx, pm_workspace** out);
void pm_workspace_destroy(pm_workspace* ws);

/* ---- placement::placeLite compute part (placement.cpp:986-1950) for one sample.
 *      reads: concatenated read bytes (ASCII, any case, non-ACGT = ambiguous); read_offsets[n_reads+1].
 *      Buffers are HOST memory; the call uploads them, runs every stage on the GPU and returns the result. */
int pm_place(pm_workspace* ws, const char* reads, const uint64_t* read_offsets, uint64_t n_reads,
             const pm_place_params* params, pm_place_result* result);
/* The same for reads that already are 4-bit base codes (what a FASTQ parser can emit directly: panmap_b200/host/placement.cpp does, and
 * pm_pack_reads converts ASCII on the host): half the bytes cross PCIe and the device never sees ASCII.
 *   code: A 0, C 1, G 2, T 3 (either case), anything else 4 (ambiguous: seeding.hpp:86-112 gives every other byte hash 0)
 *   layout: 16-byte chunks of 32 bases, base j of a chunk in bits [4j, 4j+4) of the chunk read as a little-endian 128-bit word; slots past
 *   the end of a read hold 4; every read starts on a chunk boundary: read r occupies chunks [c_r, c_r + ceil(len_r / 32)), c_r = sum of the
 *   chunk counts before it.  read_offsets are the BASE offsets as for pm_place (lengths = differences).  `packed` must be 16-byte aligned.
 * Needs the raw strings and therefore returns PM_ERR_UNSUPPORTED: dedup_reads, hpc indexes, min_seed_quality. */
int pm_place_packed(pm_workspace* ws, const void* packed, const uint64_t* read_offsets, uint64_t n_reads, const pm_place_params* params,
                    pm_place_result* result);
/* ASCII -> the layout above on the host with up to `threads` threads (0 = all cores); packed_out holds pm_packed_chunks() * 16 bytes */
uint64_t pm_packed_chunks(const uint64_t* read_offsets, uint64_t n_reads);
int pm_pack_reads(const char* reads, const uint64_t* read_offsets, uint64_t n_reads, void* packed_out, int threads);
/* The same with per-base qualities for --min-seed-quality (placement.cpp:1179-1240, 1388-1533): quals has one Phred+33 byte per base at the
 * reads' offsets (a FASTQ record's quality line; for records without one the reference substitutes 'I', placement.cpp:211).
 * With params->min_seed_quality <= 0 the qualities are ignored and this is pm_place.  On this path the reference does not deduplicate
 * reads, so dedup_reads has no effect.  pm_place with min_seed_quality > 0 returns PM_ERR_INVALID. */
int pm_place_quality(pm_workspace* ws, const char* reads, const char* quals, const uint64_t* read_offsets, uint64_t n_reads,
                     const pm_place_params* params, pm_place_result* result);
/* "inputs already resident in HBM": pm_reads_upload copies + lays out the reads once (untimed by callers that
 * measure device throughput), pm_place_resident then runs every stage on them; may be called repeatedly. */
int pm_reads_upload(pm_workspace* ws, const char* reads, const uint64_t* read_offsets, uint64_t n_reads);
int pm_place_resident(pm_workspace* ws, const pm_place_params* params, pm_place_result* result);
/* the same layout step for a sample that already sits in HBM (e.g. one of many samples a batch caller keeps in a device pool): d_reads /
 * d_read_offsets are DEVICE pointers, h_read_offsets the same offsets on the host (sizing); device-to-device copies, enqueued without waiting */
int pm_reads_upload_device(pm_workspace* ws, const char* d_reads, const uint64_t* d_read_offsets, const uint64_t* h_read_offsets, uint64_t n_reads);
/* page-locked host buffers for callers that want the H2D copies of pm_place to run at full PCIe speed */
void* pm_host_alloc(uint64_t bytes);
/* Pinned landing buffers owned by the workspace, for a parser that writes a sample straight into memory the copy engine can read
 * (parallelFastqSeqs, placement.cpp:96-162, fills std::strings instead): room for read_bytes bases (+ qualities when want_quals) and
 * n_reads + 1 offsets.  Grown as needed, kept across samples, valid until the next call or pm_workspace_destroy. */
int pm_workspace_staging(pm_workspace* ws, uint64_t read_bytes, uint64_t n_reads, int want_quals, char** reads_out, uint64_t** offsets_out, char** quals_out);
void pm_host_free(void* p);

/* after pm_place: tied node lists (sorted ascending, == tied*NodeIndices), per-node scores, read seed table */
int pm_get_tied(pm_workspace* ws, int metric, uint32_t* out, uint64_t cap);
int pm_get_node_scores(pm_workspace* ws, double* out /* [n_nodes][5] */);
int pm_get_node_metrics(pm_workspace* ws, double* out /* [n_nodes][5]: logRawNum, logCosNum, presence, wcNum, logContNum */);
/* CUDA-event time of the three seeding kernels of the last pm_place_resident call, in ms: 0 pack_reads, 1 syncmers_*,
 * 2 count_seeds / seeds_from_syncmers (profiling aid for bench.py; the same events bracket nothing else) */
/* stage timers: CUDA events between the stages of every placement of this workspace (stage_ms[0..6], pm_last_kernel_ms).  Off by default --
 * like the reference's own stage timers (debug output, placement.cpp:1128,1698,1836,1926) -- because the ten stream markers cost ~25 us per
 * placement; stage_ms[7], the whole placement, is always measured.  PM_STAGE_EVENTS=1 in the environment turns them on for new workspaces. */
int pm_workspace_set_stage_timers(pm_workspace* ws, int on);
int pm_last_kernel_ms(pm_workspace* ws, float* out /* [3] */);
int pm_get_seed_table(pm_workspace* ws, uint64_t* hash, int64_t* count, uint64_t cap); /* unsorted; returns n or <0 */

/* ---- own index builder (reference: IndexBuilder::buildIndex, index_single_mode.cpp:1227-1392 / processNode :1647-2205,
 *      called from main.cpp:398-428): `.panman` -> per-node seed deltas.  Every node's ungapped genome (the reference's
 *      getStringFromReference(tree, id, aligned = false), panmap_utils.cpp:7-193) is seeded on the GPU with the read path's kernels and
 *      diffed against its parent's: the LiteIndex the reference builds with --flank-mask 0 (rsv_4K: every node; sars_20000: 39,998 of 39,999;
 *      DESIGN.md section 8).  flank_mask and sp->hpc must be 0 (PM_ERR_UNSUPPORTED otherwise: the reference's flank masking makes its index
 *      depend on the traversal history, and its hpc build is not the seeding of the collapsed genomes).
 *      The result is a host index like pm_host_index_read's: pm_host_index_desc / _extras / _write / pm_index_create apply. ---- */
int pm_index_build(const char* panman_path, const pm_seed_params* sp, int flank_mask, int device, pm_host_index** out);
/* the builder's input half alone (host only, no GPU): every node's ungapped genome concatenated in DFS pre-order (offsets[n_nodes + 1]),
 * parent indexes, the node ids joined by '\n' and, per base, its aligned (global scalar) coordinate; each buffer is malloc'ed (free with
 * pm_free); parent_index / ids_joined / coords may be null */
int pm_panman_genomes(const char* panman_path, char** bases, uint64_t** offsets, uint32_t** parent_index, char** ids_joined, uint32_t** coords,
                      uint64_t* n_nodes);

/* ---- seeding::hashSeq (seeding.hpp:123, seeding.cpp:20-30) for a batch of k-mers: forward and reverse-complement hash of each sequence;
 *      PM_ERR_INVALID ("Kmer contains non canonical base") when a sequence holds anything but ACGT/acgt, like the reference's exception ---- */
int pm_hash_seq(int device, const char* seqs, const uint64_t* seq_offsets, uint64_t n_seqs, uint64_t* out_fwd, uint64_t* out_rev);

/* ---- seeding::rollingSyncmers (seeding.cpp:47-229) for a batch of sequences, returnAll=false form:
 *      per sequence i the syncmers are written at out_*[win_offsets[i] ...] where
 *      win_offsets[i] = sum_{j<i} max(0, len_j - k + 1); out_count[i] = number of syncmers of sequence i. ---- */
int pm_rolling_syncmers(int device, const char* seqs, const uint64_t* seq_offsets, uint64_t n_seqs,
                        int k, int s, int open, int t,
                        uint64_t* out_hash, uint8_t* out_is_reverse, int64_t* out_pos, uint64_t* out_count);
/* per-read seeds as placeLite counts them (k-min-mers for l>1, placement.cpp:1598-1686); same output layout */
int pm_read_seeds(int device, const char* seqs, const uint64_t* seq_offsets, uint64_t n_seqs,
                  const pm_seed_params* sp, int trim_start, int trim_end, uint64_t* out_hash, uint64_t* out_count);

/* ---- staged entry points for multi-GPU runs: one process per GPU, each holding one shard of the node range
 *      (pm_index_create_shard) and a slice of the reads; the caller moves the small exchange buffers between
 *      ranks (NCCL / gloo all-gather).  All pointers are HOST pointers.
 *        A  pm_stage_seed            seed this rank's reads into the workspace table
 *        B  pm_stage_table_*         export (hash,count); import the concatenation of every rank's export
 *        C  pm_stage_score           filters, min-support, magnitudes, deltas, exact prefix, scores, local records
 *        D  pm_stage_records_*       local prefix-maximum records (global BFS rank, node, score) per metric
 *        E  pm_stage_select          tolerance chain over ALL ranks' records (on the device) + this shard's ties;
 *                                    the union over ranks of pm_get_tied() is the reference's tied list ---- */
int pm_stage_seed(pm_workspace* ws, const char* reads, const uint64_t* read_offsets, uint64_t n_reads,
                  const pm_place_params* params);
int pm_stage_seed_resident(pm_workspace* ws, const pm_place_params* params); /* reads laid out earlier by pm_reads_upload */
int64_t pm_stage_table_size(pm_workspace* ws);
int pm_stage_table_export(pm_workspace* ws, uint64_t* hash, int64_t* count, uint64_t cap);
int pm_stage_table_import(pm_workspace* ws, const uint64_t* hash, const int64_t* count, uint64_t n);
/* same exchange with DEVICE buffers owned by the caller (e.g. tensors handed to NCCL); export with d_hash == NULL only counts */
int pm_stage_table_export_dev(pm_workspace* ws, uint64_t* d_hash, int64_t* d_count, uint64_t cap, uint64_t* n_out);
int pm_stage_table_import_dev(pm_workspace* ws, const uint64_t* d_hash, const int64_t* d_count, uint64_t n);
/* enqueue only (no wait; an overflow surfaces in pm_stage_score); clear_first starts a new table sized for expected_total entries */
int pm_stage_table_import_dev_async(pm_workspace* ws, const uint64_t* d_hash, const int64_t* d_count, uint64_t n, int clear_first,
                                    uint64_t expected_total);
int pm_stage_score(pm_workspace* ws, const pm_place_params* params);
int64_t pm_stage_records_size(pm_workspace* ws, int metric);
int pm_stage_records_export(pm_workspace* ws, int metric, uint32_t* bfs_rank, uint32_t* node, double* score, uint64_t cap);
int pm_stage_records_export_all(pm_workspace* ws, uint32_t* counts /* [5] */, uint32_t* bfs_rank, uint32_t* node, double* score /* [5][cap] */, uint64_t cap);
int pm_stage_select(pm_workspace* ws, const uint32_t* counts /* [5] */, const uint32_t* const* bfs_rank /* [5] */,
                    const uint32_t* const* node /* [5] */, const double* const* score /* [5] */, uint64_t total_reads,
                    pm_place_result* result);

/* ---- ONE sample over several GPUs (BASELINE configs[2]: node range sharded over 1/2/4/8 GPUs; the reference's own parallel form is
 *      the level-parallel traversal + per-thread seed maps merged at the end, placement.cpp:742-913, 922-929).
 *      Rank r holds shard r of the node range (pm_index_create_shard(desc, device, r, n)) and seeds its own slice of the reads; the seed
 *      table is hash-partitioned over the ranks (all-to-all), finalized where it lives, and the per-seed log counts every shard's delta
 *      kernel needs come back with one all-gather; prefix-maximum records and tie heads take one small all-gather each.  Everything is
 *      enqueued on the workspace stream with fixed-capacity buffers and in-band counts: the host waits once, for the result.
 *      Results are bit-identical to pm_place on one GPU.  dedup_reads, seed_mask_fraction and min_seed_quality need the whole sample
 *      in one place and return PM_ERR_UNSUPPORTED here.
 *      Two transports:
 *        pm_comm_create_nccl   one process per GPU; NCCL (libnccl.so.2, resolved at run time) on the workspace stream.  The 128-byte id
 *                              comes from pm_comm_unique_id on rank 0 and reaches the other ranks through the caller (MPI, torchrun, a file).
 *        pm_comm_create_local  one process, one host thread driving n workspaces (any mix of devices, also all on one): peer copies
 *                              ordered by events.  This is the form panmap itself (a single process) would call. ---- */
#define PM_COMM_ID_BYTES 128
int pm_comm_unique_id(void* id_out /* PM_COMM_ID_BYTES */);
int pm_comm_create_nccl(pm_workspace* ws, const void* id, int rank, int n_ranks, pm_comm** out);
int pm_comm_create_local(pm_workspace* const* ws /* [n_ranks] */, int n_ranks, pm_comm** out /* [n_ranks] */);
void pm_comm_destroy(pm_comm* c);
/* NCCL transport: every rank calls with its slice of the reads (HOST buffers, offsets start at 0); every rank receives the same result.
 * _resident: the slice was laid out by pm_reads_upload on the communicator's workspace. */
int pm_place_sharded(pm_comm* c, const char* reads, const uint64_t* read_offsets, uint64_t n_reads_local, const pm_place_params* params,
                     pm_place_result* result);
int pm_place_sharded_resident(pm_comm* c, const pm_place_params* params, pm_place_result* result);
/* the rank's slice as 4-bit codes (layout and restrictions of pm_place_packed) */
int pm_place_sharded_packed(pm_comm* c, const void* packed, const uint64_t* read_offsets, uint64_t n_reads_local, const pm_place_params* params,
                            pm_place_result* result);
/* local transport: the whole sample in, cut into n contiguous slices of reads internally; the result (and pm_get_tied) is available on every
 * rank's workspace, `result` receives rank 0's. _resident: slice r was laid out by pm_reads_upload on workspace r. */
int pm_place_multi(pm_comm* const* comms, int n_ranks, const char* reads, const uint64_t* read_offsets, uint64_t n_reads,
                   const pm_place_params* params, pm_place_result* result);
int pm_place_multi_resident(pm_comm* const* comms, int n_ranks, const pm_place_params* params, pm_place_result* result);
/* bytes this rank sent / received through the transport for the last sample (collective payloads, capacities not fill levels) */
int pm_comm_last_traffic(pm_comm* c, uint64_t* bytes_sent, uint64_t* bytes_received);
/* how the exchanges of this communicator travel: "local" (in-process copies), "nccl", or "nccl bootstrap + peer-memory exchanges"
 * (every rank's receive buffers mapped into the others through CUDA IPC; the payloads are stored over NVLink, NCCL only bootstraps).
 * Decided at the first sample; PM_PEER_EXCHANGE=0 keeps NCCL. */
const char* pm_comm_transport(pm_comm* c);

/* placement::placeLite through the C++ host shim (panmap_b200/host/placement.hpp): FASTA/FASTQ(.gz) files in, result + <out_tsv> out.
 * node_ids = LiteNode ids by DFS index (for the TSV); err receives the exception text on failure. */
int pm_place_files(pm_index* idx, pm_workspace* ws, const char* const* node_ids, uint64_t n_ids, const char* reads1, const char* reads2,
                   const char* out_tsv, const pm_place_params* params, pm_place_result* result, char* err, uint64_t err_cap);

/* extractReadSequences (placement.cpp:164-197) alone: malloc'd bases/offsets of reads1 (+ reads2, pairs interleaved); free with pm_free */
int pm_read_fastx(const char* reads1, const char* reads2, char** bases, uint64_t** offsets, uint64_t* n_reads, char* err, uint64_t err_cap);
void pm_free(void* p);

/* reference BFS visit rank of every node (children ascending, level by level; placement.cpp:742-827) */
int pm_index_bfs_ranks(const pm_index* idx, uint32_t* out /* [n_nodes] */);

#ifdef __cplusplus
}
#endif
#endif /* PANMAP_B200_H */
