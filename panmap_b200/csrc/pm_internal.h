// pm_internal.h -- host-side structures shared by the translation units behind the C ABI (pm_api.cu: single-GPU pipeline and entry
// points; pm_multi.cu: one sample over several GPUs).  Not installed, not part of the ABI.
#pragma once
#include "pm_host.h"
#include "pm_kernels.cuh"

#include <algorithm>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace pm {
namespace host {

int fail(int code, const std::string& msg);   // sets the calling thread's pm_last_error() text, returns code

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            throw ::pm::host::CudaError(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

template <class F>
int guarded(F&& f) {
    try { return f(); }
    catch (const CudaError& e) { return fail(PM_ERR_CUDA, e.what()); }
    catch (const IoError& e) { return fail(PM_ERR_IO, e.what()); }
    catch (const Unsupported& e) { return fail(PM_ERR_UNSUPPORTED, e.what()); }
    catch (const std::bad_alloc&) { return fail(PM_ERR_CAPACITY, "out of host memory"); }
    catch (const std::exception& e) { return fail(PM_ERR_INVALID, e.what()); }
}
template <class T>
struct DevBuf {
    T* p = nullptr; size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    void alloc(size_t count) {
        if (p) { cudaFree(p); p = nullptr; }
        n = count;
        if (count) CK(cudaMalloc(&p, count * sizeof(T)));
    }
    void ensure(size_t count) { if (count > n) alloc(count + count / 8); }
    void upload(const std::vector<T>& v, cudaStream_t st = 0) {
        alloc(v.size());
        if (!v.empty()) { CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st)); CK(cudaStreamSynchronize(st)); }
    }
};
template <class T>
struct PinBuf {
    T* p = nullptr; size_t n = 0;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    void ensure(size_t count) {
        if (count <= n) return;
        if (p) { cudaFreeHost(p); p = nullptr; }
        n = count + count / 8;
        CK(cudaMallocHost(&p, n * sizeof(T)));
    }
};

int deviceCountNoThrow();


}  // namespace host
}  // namespace pm

using pm::u32; using pm::u64;
using pm::host::DevBuf; using pm::host::PinBuf;
struct pm_host_index { pm::HostIndex h; std::vector<const char*> idPtrs; };

struct pm_index {
    int device = 0; int nSM = 148;
    pm::FlatIndex F;  // host copy of the small arrays (tree) is kept for result assembly; big vectors are released
    DevBuf<u32> dw, endMask, chunkSeg, nodeSeg, boundarySegs, genSlot, genId, genPc, evSlot, evIdx, rootId, rootChild;
    DevBuf<u32> parent, subEnd, carrySlot, chainOff, chainNodes, bfsNodes, bfsRanks;
    DevBuf<u64> dictHash, homo;
    DevBuf<pm::DictSlot> dict;
    DevBuf<double> gMag, log1pLut, log1pSmall;
    DevBuf<unsigned char> isLeaf;
    DevBuf<pm::SeedTables> seedTables;
    pm::DevIndexView view{};
    std::vector<double> gMagSqHost; std::vector<int64_t> gUniqueHost;
    std::vector<std::string> nodeIds;   // LiteNode.id of every node when the index came from a file or an image (empty otherwise)
};

struct pm_workspace {
    pm_index* idx = nullptr;
    int device = 0;   // == idx->device (kept here: a binding's garbage collector may destroy the index handle first)
    cudaStream_t st = nullptr, stCopy = nullptr;
    cudaEvent_t ev[9]{}, evCopy[33]{}, evK[4]{}, evFork{}, evJoin{};   // evFork / evJoin: root_and_scalars beside node_deltas (stageScore)
    bool joinPending = false;
    bool stageTimers = false;   // record the events between the stages of a placement (pm_workspace_set_stage_timers)   // evK: around pack_reads / syncmers / count_seeds of the resident path
    // inputs
    DevBuf<char> reads; DevBuf<u64> off, packedOff; DevBuf<uint4> packed; DevBuf<u32> blockFirst;
    PinBuf<u64> hPackedOff; PinBuf<u32> hBlockFirst;
    PinBuf<char> hIngestReads, hIngestQuals; PinBuf<u64> hIngestOff;   // pm_workspace_staging: pinned landing buffers of a file parser
    u64 nReads = 0, nChunks = 0, totalBases = 0, totalWindows = 0, maxReadLen = 0;   // maxReadLen: longest read of the resident sample (0 = unknown)
    bool residentValid = false;  // the device copy of the reads was laid out by pm_reads_upload (not by the sliced pm_place path)
    bool uploadPending = false;  // pm_reads_upload_device enqueued copies from the pinned staging arrays and did not wait
    bool hpcDone = false;        // hpc indexes: the resident reads (and qualities) were compressed in place already, endOff is valid
    // table
    DevBuf<pm::TableSlot> table; u64 tableCap = 0; u64 lastEntries = 0; u64 tableLimit = 0; cudaTextureObject_t tableTex = 0;
    DevBuf<unsigned long long> dedupSlots; u64 dedupMask = 0; DevBuf<unsigned char> dupFlag;   // --dedup only
    DevBuf<u64> endOff;                                                                        // hpc indexes only
    DevBuf<u64> tileSum;                                                                       // scratch of the device-side chunk offsets
    DevBuf<u64> synBuf; DevBuf<unsigned> synCount;
    DevBuf<u64> bktBuf; DevBuf<u32> bktFill; u32 bktCount = 0, bktShift = 0, bktRegionCap = 0;   // partitioned counting (pm_kernels.cu), set per sample by runPlace
    DevBuf<u64> missQ;   // miss queue of the counting kernels: a quarter of the syncmer slots (overflow falls back to direct insertion)
    // the small per-sample result block lives in ONE device allocation so that it comes back with a single copy:
    // [pm::SampleAcc | pm::SampleScalars | pm::Selection x 5 | first kTieHead tied nodes of every metric]
    DevBuf<unsigned char> resultBlob;
    struct { pm::SampleAcc* p = nullptr; } acc; struct { pm::SampleScalars* p = nullptr; } scalars; struct { pm::Selection* p = nullptr; } sel;
    u32* tieHead = nullptr;
    DevBuf<long long> ell; cudaTextureObject_t ellTex = 0; DevBuf<unsigned> countHist; DevBuf<u64> entKey; DevBuf<u32> entCnt, entId;
    DevBuf<pm::ScanPartial> scanPart; DevBuf<pm::FinPartial> finPart;
    DevBuf<pm::SegRec> segRec, chainA;
    DevBuf<u64> genRec, evPrefix;
    DevBuf<double> scores, metrics, blockMaxAndBfs;
    DevBuf<u32> recRank, recNode; DevBuf<double> recScore; u32 recCap = 0;
    DevBuf<u32> tieNode; u32 tieCap = 0; DevBuf<u32> selCounts;
    DevBuf<u64> expHash; DevBuf<long long> expCount; DevBuf<unsigned> expCounter;
    DevBuf<unsigned long long> maskScratch;   // --seed-mask-fraction only
    DevBuf<char> quals; DevBuf<unsigned char> synPass; bool useQuals = false;   // --min-seed-quality only (pm_place_quality)
    // host staging (pinned)
    PinBuf<unsigned char> hStage; PinBuf<u32> hTies; PinBuf<unsigned char> hRec;
    // results of the last sample
    bool haveResult = false; bool wantMetrics = false;
    pm_place_params lastParams{};
    pm::Selection hSel[5]{}; pm::SampleAcc hAcc{}; pm::SampleScalars hScal{};
    std::vector<u32> tied[5];
    pm::WorkspaceView view{};
};


namespace pm {
namespace host {

constexpr size_t kResultBlobBytes = sizeof(SampleAcc) + sizeof(SampleScalars) + 5 * sizeof(Selection) + 5 * kTieHead * sizeof(u32);

void setDevice(int dev);
bool workspaceAlive(const pm_workspace* W);   // false once pm_workspace_destroy ran (communicators outlive their workspace in GC'd bindings)
void refreshView(pm_workspace* W);
void ensureTable(pm_workspace* W, u64 wantCap);   // grows the table in use to at least wantCap slots (never shrinks it)
PlaceOpts makeOpts(const pm_place_params& p, bool wantMetrics);
void checkParams(const pm_place_params* p);
void uploadReads(pm_workspace* W, const char* reads, const uint64_t* off, u64 n, bool fromDevice = false, const uint64_t* dOff = nullptr);
void uploadAndSeedPipelined(pm_workspace* W, const char* reads, const uint64_t* off, u64 n, const pm_place_params& prm);
void uploadAndSeedPipelinedPacked(pm_workspace* W, const uint4* hPacked, const uint64_t* off, u64 n, const pm_place_params& prm);
void stageSeed(pm_workspace* W, bool clearFirst, const pm_place_params& prm);
void stageScore(pm_workspace* W, const pm_place_params& prm);         // finalize + the three stages below
void stageDeltasScoresRecords(pm_workspace* W, const pm_place_params& prm);   // node_deltas, prefix_scores, local prefix-maximum records
void fetchSmall(pm_workspace* W);                                      // D2H of the small result block; returns with the stream idle
void enqueueSmall(pm_workspace* W);                                    // the two halves of fetchSmall around a synchronisation of the caller's
void parseSmall(pm_workspace* W);
void fillResult(pm_workspace* W, pm_place_result* r, u64 totalReads);
void finishTies(pm_workspace* W, const u32* lists, const unsigned* n);  // tie lists (concatenated per metric) -> W->tied, reference semantics
void recordStageTimes(pm_workspace* W, pm_place_result* res);

}  // namespace host
}  // namespace pm
