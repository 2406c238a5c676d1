// pm_api.cu -- the C ABI (include/panmap_b200.h): index upload, per-sample workspaces, the placement pipeline.
// All compute runs in the kernels of pm_kernels.cu; the host code here only prepares launches, moves the
// inputs/outputs and assembles the result structure.  There is no CPU compute fallback.
#include "pm_internal.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <initializer_list>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

using namespace pm;
using namespace pm::host;

namespace pm {
namespace host {

static std::atomic<unsigned long long> g_launches{0};
}  // namespace host
void noteLaunch() { host::g_launches.fetch_add(1, std::memory_order_relaxed); }
namespace host {
thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
int deviceCountNoThrow() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}


void setDevice(int dev) { CK(cudaSetDevice(dev)); }
// stage timers (cudaEventRecord between the kernels of a sample: ten stream markers, ~25 us per placement of the 1M-read sample).  Off unless
// the caller asks for them (pm_workspace_set_stage_timers; the reference's own stage timers are debug output too) or PM_STAGE_EVENTS=1 makes
// them the default; without them stage_ms[0..6] and pm_last_kernel_ms read 0 and only stage_ms[7], the whole placement, is measured.
bool stageTimersDefault() {
    static const bool v = [] { const char* e = std::getenv("PM_STAGE_EVENTS"); return e ? std::atoi(e) != 0 : false; }();
    return v;
}
static std::mutex g_wsMutex;
static std::vector<const pm_workspace*> g_wsLive;
bool workspaceAlive(const pm_workspace* W) {
    std::lock_guard<std::mutex> lk(g_wsMutex);
    return std::find(g_wsLive.begin(), g_wsLive.end(), W) != g_wsLive.end();
}
static void workspaceRegister(const pm_workspace* W, bool add) {
    std::lock_guard<std::mutex> lk(g_wsMutex);
    if (add) g_wsLive.push_back(W);
    else g_wsLive.erase(std::remove(g_wsLive.begin(), g_wsLive.end(), W), g_wsLive.end());
}

static void buildViews(pm_index* I) {
    FlatIndex& F = I->F;
    DevIndexView& V = I->view;
    V.nNodes = F.N; V.nodeBegin = F.nodeBegin; V.nodeEnd = F.nodeEnd; V.nLocal = F.nLocal; V.nAnc = F.nAnc;
    V.nLocalDeltas = F.nLocalDeltas; V.nSeeds = F.S;
    V.dw = I->dw.p; V.endMask = I->endMask.p; V.nDeltaChunks = F.nDeltaChunks; V.chunkSeg = I->chunkSeg.p; V.nodeSeg = I->nodeSeg.p;
    V.boundarySegs = I->boundarySegs.p; V.nBoundary = (u32)F.boundarySegs.size(); V.nSeg = F.nSeg;
    V.genSlot = I->genSlot.p; V.genId = I->genId.p; V.genPc = I->genPc.p; V.nGenDeltas = (u32)F.genSlot.size(); V.nGenNodes = F.nGenNodes;
    V.evSlot = I->evSlot.p; V.nEvents = (u32)F.evSlot.size(); V.evIdx = F.nGenNodes ? I->evIdx.p : nullptr;
    V.parent = I->parent.p; V.gMag = I->gMag.p; V.subEnd = I->subEnd.p;
    V.carrySlot = I->carrySlot.p; V.chainOff = I->chainOff.p; V.chainNodes = I->chainNodes.p;
    V.nK2Tiles = F.nK2Tiles; V.chainTotal = (u32)F.chainNodes.size();
    V.isLeaf = I->isLeaf.p;
    V.bfsNodes = I->bfsNodes.p; V.bfsRanks = I->bfsRanks.p; V.nShardNodes = F.nodeEnd - F.nodeBegin;
    V.nBfsBlocks = (V.nShardNodes + kBfsBlock - 1) / kBfsBlock;
    V.dict = I->dict.p; V.dictMask = F.dictMask; V.dictHash = I->dictHash.p;
    V.rootId = I->rootId.p; V.rootChild = I->rootChild.p; V.rootDCount = (u32)F.rootId.size(); V.hasRoot = 1;
    V.log1pLut = I->log1pLut.p; V.log1pSmall = I->log1pSmall.p;
    V.ln2 = std::log1p(1.0);
}

// the flattened index in I->F goes to the device (everything pm_index_create does after flattenIndex; also the whole of opening a cached image)
static void uploadIndex(pm_index* I, int device) {
    {
        I->device = device;
        setDevice(device);
        cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
        I->nSM = prop.multiProcessorCount;
        FlatIndex& F = I->F;
        I->dw.upload(F.dw); I->endMask.upload(F.endMask); I->chunkSeg.upload(F.chunkSeg); I->nodeSeg.upload(F.nodeSeg); I->boundarySegs.upload(F.boundarySegs);
        I->genSlot.upload(F.genSlot); I->genId.upload(F.genId); I->genPc.upload(F.genPc); I->evSlot.upload(F.evSlot); I->evIdx.upload(F.evIdx);
        I->rootId.upload(F.rootId); I->rootChild.upload(F.rootChild);
        I->parent.upload(F.parent); I->gMag.upload(F.gMag); I->subEnd.upload(F.subEnd);
        I->carrySlot.upload(F.carrySlot); I->chainOff.upload(F.chainOff); I->chainNodes.upload(F.chainNodes);
        I->isLeaf.upload(F.isLeaf); I->bfsNodes.upload(F.bfsNodes); I->bfsRanks.upload(F.bfsRanks);
        {
            std::vector<DictSlot> d(F.dictKeys.size());
            for (size_t i = 0; i < d.size(); ++i) { d[i].key = F.dictKeys[i]; d[i].id = F.dictVals[i]; d[i].pad = 0; }
            I->dict.upload(d);
        }
        I->dictHash.upload(F.dictHash);
        {
            std::vector<double> lut(kLog1pLut), small(32768);
            for (int c = 0; c < kLog1pLut; ++c) lut[c] = std::log1p((double)c);
            for (int c = 0; c < 32768; ++c) small[c] = std::log1p((double)c);
            I->log1pLut.upload(lut); I->log1pSmall.upload(small);
            std::vector<u64> homo(F.homo, F.homo + 4);
            I->homo.upload(homo);
            std::vector<SeedTables> st(kSeedTableElems);   // + the s = 8 rank table behind element 0 (syncmers_rank)
            buildSeedTableImage(st.data(), F.sp.k, F.sp.s);
            I->seedTables.upload(st);
        }
        I->gMagSqHost = F.gMagSq; I->gUniqueHost = F.gUnique;
        buildViews(I);
        // release the big host vectors (the device now owns them)
        std::vector<u32>().swap(F.dw); std::vector<u32>().swap(F.endMask); std::vector<u32>().swap(F.nodeSeg); std::vector<u32>().swap(F.evIdx); std::vector<u64>().swap(F.dictKeys);
        std::vector<u32>().swap(F.dictVals); std::vector<u64>().swap(F.dictHash);
        std::vector<double>().swap(F.gMagSq); std::vector<int64_t>().swap(F.gUnique);
    }
}

static int noDevice(int device) {
    return fail(PM_ERR_NO_DEVICE, "no usable CUDA device " + std::to_string(device) + " (this library has no CPU fallback)");
}

static int createIndex(const pm_index_desc* desc, int device, uint32_t shard, uint32_t nShards, pm_index** out) {
    if (!desc || !out) return fail(PM_ERR_INVALID, "null argument");
    *out = nullptr;
    if (deviceCountNoThrow() <= device || device < 0) return noDevice(device);
    return guarded([&]() -> int {
        std::unique_ptr<pm_index> I(new pm_index());
        flattenIndex(*desc, shard, nShards, I->F);
        uploadIndex(I.get(), device);
        *out = I.release();
        return PM_OK;
    });
}

void refreshView(pm_workspace* W) {
    WorkspaceView& V = W->view;
    V.table = W->table.p; V.tableTex = W->tableTex; V.tableCap = W->tableCap; V.tableMask = W->tableCap ? W->tableCap - 1 : 0;
    V.synBuf = W->synBuf.p; V.synCount = W->synCount.p; V.missQ = W->missQ.p; V.missCap = W->missQ.n;
    V.bktBuf = W->bktBuf.p; V.bktFill = W->bktFill.p; V.bktCount = W->bktCount; V.bktShift = W->bktShift; V.bktRegionCap = W->bktRegionCap;
    V.acc = W->acc.p; V.ell = W->ell.p; V.ellTex = W->ellTex; V.countHist = W->countHist.p; V.entKey = W->entKey.p; V.entCnt = W->entCnt.p; V.entId = W->entId.p;
    V.scanPart = W->scanPart.p; V.finPart = W->finPart.p;
    V.segRec = W->segRec.p; V.chainA = W->chainA.p; V.genRec = W->genRec.p; V.evPrefix = W->evPrefix.p;
    V.scores = W->scores.p; V.metrics = W->wantMetrics ? W->metrics.p : nullptr; V.blockMax = W->blockMaxAndBfs.p;
    V.recRank = W->recRank.p; V.recNode = W->recNode.p; V.recScore = W->recScore.p; V.recCap = W->recCap;
    V.tieNode = W->tieNode.p; V.tieCap = W->tieCap; V.tieHead = W->tieHead; V.sel = W->sel.p; V.scalars = W->scalars.p;
}

// table capacity is kept between 2x and 4x the unique seeds of the previous sample (tighter fits were measured: the insertions
// lose what the table passes gain)
static u64 fitLo() {   // tuning override PM_TABLE_FIT_LO (3 = tables between 1.5x and 3x the entries)
    static const u64 v = [] { const char* e = std::getenv("PM_TABLE_FIT_LO"); const u64 x = e ? std::strtoull(e, nullptr, 10) : 4; return x >= 3 && x <= 16 ? x : 4; }();
    return v;
}
static u64 fitHi() { return 2 * fitLo(); }

// grows the table in use to at least wantCap slots (never shrinks it); a prefix of a larger allocation is reused as it is.
// The table is also bound as a linear texture (first probes go through the texture path), so its size is capped by the device's
// linear-texture width (2^27 or 2^28 16-byte slots).  Estimates above the cap are clamped to it -- callers size from the number of
// k-mer windows, an upper bound that bacterial-scale samples (4e8 seed instances, ~5e7 distinct) overshoot by far; a sample whose
// distinct seeds really do not fit fails when the full-size table runs tight.
static u64 tableSlotLimit(pm_workspace* W) {
    if (!W->tableLimit) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxTexture1DLinearWidth, W->device) != cudaSuccess || v <= 0) { cudaGetLastError(); v = 1 << 27; }
        u64 lim = 1; while (lim * 2 <= (u64)v) lim <<= 1;
        if (const char* e = std::getenv("PM_TABLE_SLOT_LIMIT")) { const u64 x = std::strtoull(e, nullptr, 10); if (x >= (1u << 12) && x < lim) { lim = 1; while (lim * 2 <= x) lim <<= 1; } }   // tests
        W->tableLimit = lim;
    }
    return W->tableLimit;
}
void ensureTable(pm_workspace* W, u64 wantCap) {
    const u64 limit = tableSlotLimit(W);
    u64 cap = 1 << 12;
    while (cap < wantCap && cap < limit) cap <<= 1;
    if (cap <= W->tableCap) {
        if (wantCap > limit && W->tableCap >= limit)
            throw std::runtime_error("read seed table would exceed " + std::to_string(limit) + " slots (the device's linear-texture width)");
        return;
    }
    if (cap <= W->table.n) { W->tableCap = cap; refreshView(W); return; }   // the allocation (and its texture) already covers it
    if (W->tableTex) { cudaDestroyTextureObject(W->tableTex); W->tableTex = 0; }
    W->table.alloc(cap); W->tableCap = cap; W->entKey.alloc(cap); W->entCnt.alloc(cap); W->entId.alloc(cap);
    {
        cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = W->table.p;
        rd.res.linear.desc = cudaCreateChannelDesc<uint4>(); rd.res.linear.sizeInBytes = cap * sizeof(TableSlot);
        cudaTextureDesc td{}; td.readMode = cudaReadModeElementType;
        CK(cudaCreateTextureObject(&W->tableTex, &rd, &td, nullptr));
    }
    refreshView(W);
}

PlaceOpts makeOpts(const pm_place_params& p, bool wantMetrics) {
    PlaceOpts o; o.minReadSupport = p.min_read_support; o.forceLeaf = p.force_leaf; o.skipNode = p.skip_node_index;
    o.wantMetrics = wantMetrics ? 1 : 0; return o;
}

void checkParams(const pm_place_params* p) {
    if (!p) throw std::runtime_error("null params");
    if (p->trim_start < 0 || p->trim_end < 0) throw std::runtime_error("negative trim");
}

// host side: chunk offsets of the packed layout (ceil(len/32) 16-byte chunks per read)
static void hostPackedOffsets(pm_workspace* W, const uint64_t* off, u64 n, int k) {
    W->hPackedOff.ensure(n + 1);
    u64 acc = 0, win = 0, mx = 0;
    for (u64 i = 0; i < n; ++i) {
        W->hPackedOff.p[i] = acc;
        if (off[i + 1] < off[i]) throw std::runtime_error("read offsets not monotone");
        const u64 L = off[i + 1] - off[i];
        if (L > 0x7FFFFFF0ull) throw std::runtime_error("read longer than 2^31 bases");
        acc += (L + 31) >> 5;
        if (L >= (u64)k) win += L - (u64)k + 1;
        mx = std::max(mx, L);
    }
    W->hPackedOff.p[n] = acc;
    W->nChunks = acc; W->totalWindows = win; W->maxReadLen = mx;
    W->hBlockFirst.ensure((acc + 255) / 256 + 1);
    packBlockFirst(W->hPackedOff.p, n, acc, W->hBlockFirst.p);
}

// fromDevice: `reads` and `dOff` already live in HBM (a caller-side pool of samples): device-to-device copies, no host wait
void uploadReads(pm_workspace* W, const char* reads, const uint64_t* off, u64 n, bool fromDevice, const uint64_t* dOff) {
    pm_index* I = W->idx;
    const u64 base0 = n ? off[0] : 0;
    const u64 total = n ? off[n] - base0 : 0;
    hostPackedOffsets(W, off, n, I->F.sp.k);
    W->nReads = n; W->totalBases = total; W->hpcDone = false;
    W->synBuf.ensure(W->nChunks * 32 + 32); W->synCount.ensure(n + 1); W->missQ.ensure(W->nChunks * 8 + 65536);
    W->reads.ensure(total + 64); W->off.ensure(n + 1); W->packedOff.ensure(n + 1); W->packed.ensure(W->nChunks + 1);
    if (base0 != 0) throw std::runtime_error("read_offsets[0] must be 0");
    if (total) CK(cudaMemcpyAsync(W->reads.p, reads, total, fromDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, W->st));
    if (fromDevice) CK(cudaMemcpyAsync(W->off.p, dOff, (n + 1) * sizeof(u64), cudaMemcpyDeviceToDevice, W->st));
    else CK(cudaMemcpyAsync(W->off.p, off, (n + 1) * sizeof(u64), cudaMemcpyHostToDevice, W->st));
    CK(cudaMemcpyAsync(W->packedOff.p, W->hPackedOff.p, (n + 1) * sizeof(u64), cudaMemcpyHostToDevice, W->st));
    W->blockFirst.ensure((W->nChunks + 255) / 256 + 1);
    CK(cudaMemcpyAsync(W->blockFirst.p, W->hBlockFirst.p, ((W->nChunks + 255) / 256 + 1) * sizeof(u32), cudaMemcpyHostToDevice, W->st));
}

// --dedup: (re)size and clear the read set; returns the flag array (null when the option is off)
static unsigned char* prepareDedup(pm_workspace* W, u64 n, const pm_place_params& prm) {
    if (!prm.dedup_reads || n == 0) return nullptr;
    if (n >= 0xFFFFFFFFull) throw std::runtime_error("dedup_reads: more than 2^32 reads");
    u64 cap = 1024;
    while (cap < 2 * n) cap <<= 1;
    if (W->dedupSlots.n < cap) W->dedupSlots.alloc(cap);
    W->dedupMask = cap - 1;
    W->dupFlag.ensure(n + 1);
    CK(cudaMemsetAsync(W->dedupSlots.p, 0xFF, cap * sizeof(unsigned long long), W->st));
    return W->dupFlag.p;
}

// Host buffers -> table, pipelined: the sample is cut into slices of reads; slice i+1 is copied (copy stream) while slice i is
// packed, seeded and counted (compute stream).  Host-side chunk offsets of a slice are computed just before its copy.
// slice schedule of the host-buffer pipelines: cumulative fractions of the reads (tuning override: PM_SLICES_ASCII / PM_SLICES_PACKED =
// comma-separated cumulative cut points ending in 1)
struct SliceSchedule { int n; double cut[33]; };
// below ~1 ms of PCIe time the per-slice launches cost more than the overlap returns (PM_SLICE_MIN_BYTES overrides, read per call: tests use it
// to drive small samples through the sliced path)
static u64 sliceMinBytes() { const char* e = std::getenv("PM_SLICE_MIN_BYTES"); return e ? (u64)std::strtoull(e, nullptr, 10) : (48ull << 20); }
static SliceSchedule scheduleFromEnv(const char* var, std::initializer_list<double> dflt) {
    SliceSchedule s{0, {0.0}};
    std::vector<double> v(dflt);
    if (const char* e = std::getenv(var)) {
        std::vector<double> u; const char* p = e;
        while (*p) { char* q; const double x = std::strtod(p, &q); if (q == p) break; u.push_back(x); p = (*q == ',') ? q + 1 : q; }
        if (!u.empty() && u.size() <= 32 && u.back() == 1.0 && std::is_sorted(u.begin(), u.end()) && u.front() > 0.0) v = u;
    }
    s.n = (int)v.size();
    for (int i = 0; i < s.n; ++i) s.cut[i + 1] = v[i];
    return s;
}
void uploadAndSeedPipelined(pm_workspace* W, const char* reads, const uint64_t* off, u64 n, const pm_place_params& prm) {
    pm_index* I = W->idx;
    const int k = I->F.sp.k;
    if (n && off[0] != 0) throw std::runtime_error("read_offsets[0] must be 0");
    const u64 total = n ? off[n] : 0;
    // Slice schedule: the host prepares the chunk offsets of slice i while slice i-1 is on the wire, so the first slice is tiny
    // (its preparation is the only one nothing overlaps) and slices grow by at most 2x; they shrink again towards the end because
    // everything after the last copy (its seeding, then scoring and selection) is exposed latency.
    static const SliceSchedule sched = scheduleFromEnv("PM_SLICES_ASCII", {0.03, 0.09, 0.21, 0.40, 0.58, 0.73, 0.84, 0.91, 0.96, 1.0});
    const double* kCut = sched.cut;
    const int nSlices = total >= sliceMinBytes() ? sched.n : 1;   // small samples: one copy, one set of launches
    W->nReads = n; W->totalBases = total;
    W->hPackedOff.ensure(n + 1); W->hBlockFirst.ensure(total / 32 / 256 + n / 256 + 2 * nSlices + 32);
    W->reads.ensure(total + 64); W->off.ensure(n + 1); W->packedOff.ensure(n + 1);
    W->packed.ensure(total / 32 + n + 16); W->synBuf.ensure((total / 32 + n + 16) * 32); W->synCount.ensure(n + 1); W->missQ.ensure((total / 32 + n + 16) * 8 + 65536);
    W->blockFirst.ensure(total / 32 / 256 + n / 256 + 2 * nSlices + 32);
    if (W->tableCap == 0) ensureTable(W, std::max<u64>(1 << 16, (total > (u64)k * n ? total - (u64)(k - 1) * n : 0) / 4));
    refreshView(W);
    SeederParams P = makeSeederParams(I->F.sp.k, I->F.sp.s, I->F.sp.t, I->F.sp.l, I->F.sp.open, prm.trim_start, prm.trim_end);
    CK(cudaMemsetAsync(W->acc.p, 0, sizeof(SampleAcc), W->st));
    launchTableClear(W->view, W->st);
    unsigned char* dup = prepareDedup(W, n, prm);
    const bool ascii = seedTableReadsAscii(P);   // the default parameter sets hash straight from the bytes: no pack_reads pass
    const bool hpc = I->F.sp.hpc != 0;
    if (hpc) W->endOff.ensure(n + 1);
    if (ascii) W->tileSum.ensure(n / 4096 + 2);
    u64 chunkAcc = 0, win = 0, bfBase = 0, mxAll = 0;
    for (int sl = 0; sl < nSlices; ++sl) {
        const u64 r0 = nSlices == 1 ? 0 : (u64)((double)n * kCut[sl]), r1 = nSlices == 1 || sl + 1 == nSlices ? n : (u64)((double)n * kCut[sl + 1]);
        if (r1 == r0) continue;
        const u64 gBase = chunkAcc;
        u64 mx = 0;
        for (u64 i = r0; i < r1; ++i) {
            if (off[i + 1] < off[i]) throw std::runtime_error("read offsets not monotone");
            const u64 L = off[i + 1] - off[i];
            if (L > 0x7FFFFFF0ull) throw std::runtime_error("read longer than 2^31 bases");
            if (!ascii) W->hPackedOff.p[i] = chunkAcc;
            chunkAcc += (L + 31) >> 5;
            if (L >= (u64)k) win += L - (u64)k + 1;
            mx = std::max(mx, L);
        }
        P.maxLen = (int)std::min<u64>(mx, 0x7FFFFFFF); mxAll = std::max(mxAll, mx);   // this slice's longest read
        const u64 nCh = chunkAcc - gBase, nBlk = (nCh + 255) / 256;
        const u64 b0 = off[r0], b1 = off[r1];
        if (b1 > b0) CK(cudaMemcpyAsync(W->reads.p + b0, reads + b0, b1 - b0, cudaMemcpyHostToDevice, W->stCopy));
        CK(cudaMemcpyAsync(W->off.p + r0, off + r0, (r1 - r0 + 1) * sizeof(u64), cudaMemcpyHostToDevice, W->stCopy));
        if (!ascii) {   // pack_reads needs the chunk offsets and its block table; otherwise the offsets are rebuilt on the device
            W->hPackedOff.p[r1] = chunkAcc;
            packBlockFirst(W->hPackedOff.p + r0, r1 - r0, nCh, W->hBlockFirst.p + bfBase);
            CK(cudaMemcpyAsync(W->packedOff.p + r0, W->hPackedOff.p + r0, (r1 - r0 + 1) * sizeof(u64), cudaMemcpyHostToDevice, W->stCopy));
            CK(cudaMemcpyAsync(W->blockFirst.p + bfBase, W->hBlockFirst.p + bfBase, (nBlk + 1) * sizeof(u32), cudaMemcpyHostToDevice, W->stCopy));
        }
        CK(cudaEventRecord(W->evCopy[sl], W->stCopy));
        CK(cudaStreamWaitEvent(W->st, W->evCopy[sl], 0));
        if (ascii) launchChunkOffsets(W->off.p + r0, r1 - r0, gBase, W->tileSum.p, W->packedOff.p + r0, W->st);
        if (hpc) launchHpcCompress(W->reads.p, W->off.p + r0, r1 - r0, W->endOff.p + r0, W->st);
        if (!ascii) launchPackReads(W->reads.p, W->off.p + r0, W->packedOff.p + r0, W->blockFirst.p + bfBase, r1 - r0, gBase, nCh, W->packed.p, W->st, hpc ? W->endOff.p + r0 : nullptr);
        if (dup) launchDedup(W->reads.p, W->off.p, r0, r1, W->dedupSlots.p, W->dedupMask, dup, W->st, hpc ? W->endOff.p : nullptr);
        launchSeedTable(W->packed.p, W->off.p + r0, W->packedOff.p + r0, r1 - r0, P, I->seedTables.p, W->view, W->st, nullptr, dup ? dup + r0 : nullptr,
                        hpc ? W->endOff.p + r0 : nullptr, ascii ? W->reads.p : nullptr);
        bfBase += nBlk + 1;
    }
    launchCountBuckets(W->view, W->st);   // partitioned counting only: the slices scattered their seeds, one pass counts them
    W->nChunks = chunkAcc; W->totalWindows = win; W->maxReadLen = mxAll; W->residentValid = false; W->hpcDone = false;
}

// The same pipeline for reads that arrive as 4-bit codes (pm_place_packed): half the bytes on the wire, and the syncmer kernel takes the
// codes as they are -- no ASCII on the device at all.  hPacked: 16-byte chunks of 32 bases, read r at chunk sum_{j<r} ceil(len_j / 32).
void uploadAndSeedPipelinedPacked(pm_workspace* W, const uint4* hPacked, const uint64_t* off, u64 n, const pm_place_params& prm) {
    pm_index* I = W->idx;
    const int k = I->F.sp.k;
    if (n && off[0] != 0) throw std::runtime_error("read_offsets[0] must be 0");
    if (I->F.sp.hpc) throw Unsupported("packed reads on an hpc index: homopolymer compression works on the ASCII reads (use pm_place)");
    if (prm.dedup_reads) throw Unsupported("dedup_reads compares the raw read strings, which 4-bit codes do not preserve (use pm_place)");
    const u64 total = n ? off[n] : 0;
    static const SliceSchedule sched = scheduleFromEnv("PM_SLICES_PACKED", {0.03, 0.09, 0.21, 0.40, 0.58, 0.73, 0.84, 0.91, 0.96, 1.0});
    const double* kCut = sched.cut;
    const int nSlices = total >= sliceMinBytes() ? sched.n : 1;
    W->nReads = n; W->totalBases = total;
    W->off.ensure(n + 1); W->packedOff.ensure(n + 1);
    W->packed.ensure(total / 32 + n + 16); W->synBuf.ensure((total / 32 + n + 16) * 32); W->synCount.ensure(n + 1); W->missQ.ensure((total / 32 + n + 16) * 8 + 65536);
    W->tileSum.ensure(n / 4096 + 2);
    if (W->tableCap == 0) ensureTable(W, std::max<u64>(1 << 16, (total > (u64)k * n ? total - (u64)(k - 1) * n : 0) / 4));
    refreshView(W);
    const SeederParams P = makeSeederParams(I->F.sp.k, I->F.sp.s, I->F.sp.t, I->F.sp.l, I->F.sp.open, prm.trim_start, prm.trim_end);
    CK(cudaMemsetAsync(W->acc.p, 0, sizeof(SampleAcc), W->st));
    launchTableClear(W->view, W->st);
    u64 chunkAcc = 0, win = 0;
    for (int sl = 0; sl < nSlices; ++sl) {
        const u64 r0 = nSlices == 1 ? 0 : (u64)((double)n * kCut[sl]), r1 = nSlices == 1 || sl + 1 == nSlices ? n : (u64)((double)n * kCut[sl + 1]);
        if (r1 == r0) continue;
        const u64 gBase = chunkAcc;
        for (u64 i = r0; i < r1; ++i) {
            if (off[i + 1] < off[i]) throw std::runtime_error("read offsets not monotone");
            const u64 L = off[i + 1] - off[i];
            if (L > 0x7FFFFFF0ull) throw std::runtime_error("read longer than 2^31 bases");
            chunkAcc += (L + 31) >> 5;
            if (L >= (u64)k) win += L - (u64)k + 1;
        }
        const u64 nCh = chunkAcc - gBase;
        if (nCh) CK(cudaMemcpyAsync(W->packed.p + gBase, hPacked + gBase, nCh * sizeof(uint4), cudaMemcpyHostToDevice, W->stCopy));
        CK(cudaMemcpyAsync(W->off.p + r0, off + r0, (r1 - r0 + 1) * sizeof(u64), cudaMemcpyHostToDevice, W->stCopy));
        CK(cudaEventRecord(W->evCopy[sl], W->stCopy));
        CK(cudaStreamWaitEvent(W->st, W->evCopy[sl], 0));
        launchChunkOffsets(W->off.p + r0, r1 - r0, gBase, W->tileSum.p, W->packedOff.p + r0, W->st);
        launchSeedTable(W->packed.p, W->off.p + r0, W->packedOff.p + r0, r1 - r0, P, I->seedTables.p, W->view, W->st, nullptr, nullptr, nullptr, nullptr);
    }
    launchCountBuckets(W->view, W->st);
    W->nChunks = chunkAcc; W->totalWindows = win; W->residentValid = false; W->hpcDone = false;
}

// Partitioned counting for this sample?  Tables that do not fit L2 (PM_BUCKET_MIN_SLOTS, default 2^24 slots = 256 MB) are filled region by
// region: plan the buckets (~32 MB of table each), size one region per (resident warp, bucket) from the expected number of seed instances
// (closed syncmers of either strand: ~0.31 per k-mer window at k = 19, s = 8; 0.45 x 1.25 leaves room, a full region falls back to direct
// insertion) and clear the fill counts.  windowsUpper: an upper bound of the sample's k-mer windows.
static u64 bucketMinSlots() {   // read per call: tests drive small samples through the partitioned path with it
    const char* e = std::getenv("PM_BUCKET_MIN_SLOTS");
    return e ? (u64)std::strtoull(e, nullptr, 10) : (u64)1 << 24;
}
static void planBuckets(pm_workspace* W, u64 windowsUpper, const pm_place_params& prm) {
    const pm_seed_params& sp = W->idx->F.sp;
    W->bktCount = 0;
    const bool quality = W->useQuals && prm.min_seed_quality > 0;
    if (W->tableCap < bucketMinSlots() || quality || !bucketCountingSupports(sp.k, sp.l) || windowsUpper == 0) { refreshView(W); return; }
    u64 B = 2;
    const char* be = std::getenv("PM_BUCKET_BYTES");   // tuning override, read per sample (tools/count_probe.py sweeps it)
    const u64 bucketBytes = be ? std::max<u64>(1 << 20, std::strtoull(be, nullptr, 10)) : (u64)32 << 20;
    while (B < 256 && W->tableCap * sizeof(TableSlot) / B > bucketBytes) B <<= 1;
    unsigned shift = 0;
    while ((W->tableCap >> shift) > B) ++shift;
    const u64 expected = windowsUpper * 45 / 100;
    u64 cap = expected * 5 / 4 / ((u64)kBktWarps * B) + 256;
    cap = (cap + 3) & ~3ull;
    if (const char* e = std::getenv("PM_BUCKET_REGION_CAP")) cap = std::max<u64>(4, std::strtoull(e, nullptr, 10));   // tests: force the full-region fallback
    if (cap > 0x7FFFFFFFull) { refreshView(W); return; }
    W->bktBuf.ensure((u64)kBktWarps * B * cap); W->bktFill.ensure((u64)kBktWarps * B);
    CK(cudaMemsetAsync(W->bktFill.p, 0, (u64)kBktWarps * B * sizeof(u32), W->st));
    W->bktCount = (u32)B; W->bktShift = shift; W->bktRegionCap = (u32)cap;
    refreshView(W);
}

void stageSeed(pm_workspace* W, bool clearFirst, const pm_place_params& prm) {
    pm_index* I = W->idx;
    SeederParams P = makeSeederParams(I->F.sp.k, I->F.sp.s, I->F.sp.t, I->F.sp.l, I->F.sp.open, prm.trim_start, prm.trim_end);
    P.maxLen = (int)std::min<u64>(W->maxReadLen, 0x7FFFFFFF);
    const bool quality = W->useQuals && prm.min_seed_quality > 0;   // the reference's quality path never deduplicates (placement.cpp:1388)
    // the syncmer kernel does not touch the table: whole samples clear it on the side stream in the kernel's shadow (it leaves most of the DRAM
    // bandwidth idle) and the counting kernel waits for the event
    static const bool kSideClear = [] { const char* e = std::getenv("PM_SIDE_CLEAR"); return e ? std::atoi(e) != 0 : true; }();
    const bool sideClear = clearFirst && kSideClear && !quality && !prm.dedup_reads && !I->F.sp.hpc && W->nReads >= 100000;
    if (clearFirst) {
        if (sideClear) {   // accumulators and table in one launch on the side stream; nothing before the counting kernel's wait touches either
            CK(cudaEventRecord(W->evFork, W->st));
            CK(cudaStreamWaitEvent(W->stCopy, W->evFork, 0));
            launchSampleBegin(W->view, W->stCopy);
            CK(cudaEventRecord(W->evJoin, W->stCopy));
        } else {
            CK(cudaMemsetAsync(W->acc.p, 0, sizeof(SampleAcc), W->st));
            launchTableClear(W->view, W->st);
        }
    }
    unsigned char* dup = quality ? nullptr : prepareDedup(W, W->nReads, prm);
    const u64* endOff = nullptr;
    if (I->F.sp.hpc) {
        // in place, and NOT idempotent (a second pass would compress the compressed prefix together with the stale tail): once per
        // upload; repeated pm_place_resident calls and table-growth retries reuse the compressed bytes and endOff
        if (!W->hpcDone) {
            W->endOff.ensure(W->nReads + 1);
            launchHpcCompress(W->reads.p, W->off.p, W->nReads, W->endOff.p, W->st, W->useQuals ? W->quals.p : nullptr);
            W->hpcDone = true;
        }
        endOff = W->endOff.p;
    }
    if (dup) launchDedup(W->reads.p, W->off.p, 0, W->nReads, W->dedupSlots.p, W->dedupMask, dup, W->st, endOff);
    const bool ascii = seedTableReadsAscii(P) && !quality;
    if (W->stageTimers) CK(cudaEventRecord(W->evK[0], W->st));
    if (!ascii) launchPackReads(W->reads.p, W->off.p, W->packedOff.p, W->blockFirst.p, W->nReads, 0, W->nChunks, W->packed.p, W->st, endOff);
    if (W->stageTimers) CK(cudaEventRecord(W->evK[1], W->st));
    if (quality) {
        W->synPass.ensure(W->nChunks * 32 + 32);
        launchSeedTableQuality(W->packed.p, W->off.p, W->packedOff.p, W->nReads, P, I->seedTables.p, W->view, W->st, endOff, W->quals.p,
                               prm.min_seed_quality, W->synPass.p);
        if (W->stageTimers) CK(cudaEventRecord(W->evK[2], W->st));
    } else {
        // experiment switch (PM_RESIDENT_SLICES = n): hash and count the resident sample slice after slice, so that a slice's syncmer lists are
        // still in L2 when they are counted
        static const int kSlices = [] { const char* e = std::getenv("PM_RESIDENT_SLICES"); const int v = e ? std::atoi(e) : 1; return v < 1 ? 1 : v > 64 ? 64 : v; }();
        const int ns = (kSlices > 1 && ascii && !dup && !endOff && W->nReads >= (u64)kSlices * 1024) ? kSlices : 1;
        for (int sl = 0; sl < ns; ++sl) {
            const u64 r0 = W->nReads * (u64)sl / (u64)ns, r1 = W->nReads * (u64)(sl + 1) / (u64)ns;
            launchSeedTable(W->packed.p, W->off.p + r0, W->packedOff.p + r0, r1 - r0, P, I->seedTables.p, W->view, W->st, sl + 1 == ns && W->stageTimers ? W->evK[2] : nullptr,
                            dup ? dup + r0 : nullptr, endOff ? endOff + r0 : nullptr, ascii ? W->reads.p : nullptr, sideClear && sl == 0 ? W->evJoin : nullptr);
        }
        launchCountBuckets(W->view, W->st);   // partitioned counting only (no-op otherwise)
    }
    if (W->stageTimers) CK(cudaEventRecord(W->evK[3], W->st));
}

void stageScore(pm_workspace* W, const pm_place_params& prm) {
    pm_index* I = W->idx;
    const PlaceOpts O = makeOpts(prm, W->wantMetrics);
    if (prm.seed_mask_fraction > 0.0) W->maskScratch.ensure(2);
    static const bool kSideScalars = [] { const char* e = std::getenv("PM_SIDE_SCALARS"); return e ? std::atoi(e) != 0 : true; }();
    const bool side = kSideScalars && prm.seed_mask_fraction <= 0.0;
    launchFinalize(I->view, W->view, O, I->homo.p, W->lastEntries ? W->lastEntries : W->tableCap / 4, I->nSM, W->st, prm.seed_mask_fraction,
                   W->maskScratch.p, side ? W->stCopy : nullptr, W->evFork, W->evJoin);
    W->joinPending = side;
    stageDeltasScoresRecords(W, prm);
}
void stageDeltasScoresRecords(pm_workspace* W, const pm_place_params& prm) {
    pm_index* I = W->idx;
    const PlaceOpts O = makeOpts(prm, W->wantMetrics);
    if (W->stageTimers) CK(cudaEventRecord(W->ev[3], W->st));
    launchDeltas(I->view, W->view, I->nSM, W->st);
    launchGeneral(I->view, W->view, W->st);
    if (W->joinPending) { CK(cudaStreamWaitEvent(W->st, W->evJoin, 0)); W->joinPending = false; }   // the scalars (side stream) before the scores
    if (W->stageTimers) CK(cudaEventRecord(W->ev[4], W->st));
    launchPrefixScores(I->view, W->view, O, W->st);
    if (W->stageTimers) CK(cudaEventRecord(W->ev[5], W->st));
    launchRecords(I->view, W->view, O, W->st);
}

// D2H of the small result block: enqueue the copy, and (after the stream went idle) unpack it
void enqueueSmall(pm_workspace* W) {
    W->hStage.ensure(kResultBlobBytes);
    CK(cudaMemcpyAsync(W->hStage.p, W->resultBlob.p, kResultBlobBytes, cudaMemcpyDeviceToHost, W->st));
}
void parseSmall(pm_workspace* W) {   // runs right after a stream synchronisation
    W->uploadPending = false;
    const unsigned char* h = W->hStage.p;
    std::memcpy(&W->hAcc, h, sizeof(SampleAcc));
    std::memcpy(&W->hScal, h + sizeof(SampleAcc), sizeof(SampleScalars));
    std::memcpy(W->hSel, h + sizeof(SampleAcc) + sizeof(SampleScalars), 5 * sizeof(Selection));
}
void fetchSmall(pm_workspace* W) {
    enqueueSmall(W);
    CK(cudaStreamSynchronize(W->st));
    parseSmall(W);
}

void fillResult(pm_workspace* W, pm_place_result* r, u64 totalReads) {
    const SampleScalars& S = W->hScal;
    for (int m = 0; m < 5; ++m) {
        r->best_score[m] = W->hSel[m].best;
        r->tied_count[m] = W->tied[m].size();
        r->best_index[m] = W->tied[m].empty() ? W->hSel[m].bestNode : W->tied[m].front();
    }
    r->total_reads = totalReads;
    r->unique_seeds = (uint64_t)S.uniqueSeeds;
    r->read_unique_seed_count = (uint64_t)S.uniqueKeptInt;
    r->total_read_seed_frequency = S.totalFrequency;
    r->min_read_support = S.minSupport;
    r->read_magnitude = S.readMagnitude;
    r->log_containment_denominator = S.logContDenom;
    r->weighted_containment_denominator = S.wcDenom;
}

// tie lists -> host, finalizeTiedIndices semantics (placement.cpp:395-401): sort, unique, best = front
void finishTies(pm_workspace* W, const u32* lists, const unsigned* n) {
    size_t o = 0;
    for (int m = 0; m < 5; ++m) {
        std::vector<u32>& t = W->tied[m];
        t.assign(lists + o, lists + o + n[m]);
        o += n[m];
        const Selection& s = W->hSel[m];
        // the reference pushes bestNodeIndex (possibly UINT32_MAX before any improvement) next to every tie, and
        // an improvement leaves [node] in the list
        if (s.bestNode != kNone || !t.empty()) t.push_back(s.bestNode);
        std::sort(t.begin(), t.end());
        t.erase(std::unique(t.begin(), t.end()), t.end());
    }
}
static void fetchTies(pm_workspace* W) {
    // short lists (the usual case) came back with the result block; longer ones need one more copy
    size_t tot = 0; unsigned n[5]; bool big = false;
    for (int m = 0; m < 5; ++m) { n[m] = std::min<unsigned>(W->hAcc.tieCount[m], W->tieCap); tot += n[m]; big = big || n[m] > (unsigned)kTieHead; }
    W->hTies.ensure(tot + 1);
    size_t o = 0;
    if (big) {
        for (int m = 0; m < 5; ++m) {
            if (n[m]) CK(cudaMemcpyAsync(W->hTies.p + o, W->tieNode.p + (size_t)m * W->tieCap, n[m] * sizeof(u32), cudaMemcpyDeviceToHost, W->st));
            o += n[m];
        }
        CK(cudaStreamSynchronize(W->st));
    } else {
        const u32* head = reinterpret_cast<const u32*>(W->hStage.p + sizeof(SampleAcc) + sizeof(SampleScalars) + 5 * sizeof(Selection));
        for (int m = 0; m < 5; ++m) { std::memcpy(W->hTies.p + o, head + (size_t)m * kTieHead, n[m] * sizeof(u32)); o += n[m]; }
    }
    finishTies(W, W->hTies.p, n);
}
void recordStageTimes(pm_workspace* W, pm_place_result* res) {
    float ms = 0;
    for (int i = 0; i < 7; ++i) { ms = 0; if (W->stageTimers && cudaEventElapsedTime(&ms, W->ev[i], W->ev[i + 1]) != cudaSuccess) cudaGetLastError(); res->stage_ms[i] = ms; }
    ms = 0; if (cudaEventElapsedTime(&ms, W->ev[0], W->ev[7]) != cudaSuccess) cudaGetLastError(); res->stage_ms[7] = ms;
}

static int runPlace(pm_workspace* W, const pm_place_params* prm, pm_place_result* res, bool inputsResident, const char* reads,
             const uint64_t* off, u64 n, const uint4* packedHost = nullptr) {
    pm_index* I = W->idx;
    setDevice(I->device);
    checkParams(prm);
    if (prm->min_seed_quality > 0 && !W->useQuals) throw std::runtime_error("min_seed_quality > 0 needs the base qualities: call pm_place_quality");
    if (!res) throw std::runtime_error("null result");
    std::memset(res, 0, sizeof(*res));
    W->wantMetrics = false;
    W->lastParams = *prm;
    for (int attempt = 0; attempt < 4; ++attempt) {
        CK(cudaEventRecord(W->ev[0], W->st));
        if (W->tableCap == 0 && inputsResident) ensureTable(W, std::max<u64>(1 << 16, W->totalWindows / 4));
        else if (attempt == 0 && W->lastEntries && W->tableCap > (1u << 16) && W->tableCap > fitHi() * W->lastEntries / 2) {
            // (first attempt only: a retry has just grown the table for THIS sample and lastEntries still describes the previous one)
            // the previous sample filled under a quarter of the slots: every pass over the table is cheaper with a tighter one
            u64 cap = 1 << 16;
            while (cap < fitLo() * W->lastEntries / 2) cap <<= 1;
            if (cap < W->tableCap) { W->tableCap = cap; }   // keep the allocation, use a prefix
        }
        if (W->tableCap == 0 && !inputsResident) {   // first sample of the workspace from host buffers: size the table here, so that the plan below sees it
            const u64 total = n ? off[n] : 0, k = (u64)I->F.sp.k;
            ensureTable(W, std::max<u64>(1 << 16, (total > k * n ? total - (k - 1) * n : 0) / 4));
        }
        planBuckets(W, inputsResident ? W->totalWindows : (n ? off[n] : 0), *prm);   // also refreshes the view
        if (W->stageTimers) CK(cudaEventRecord(W->ev[1], W->st));
        if (inputsResident) stageSeed(W, true, *prm);
        else if (packedHost) uploadAndSeedPipelinedPacked(W, packedHost, off, n, *prm);
        else uploadAndSeedPipelined(W, reads, off, n, *prm);   // H2D of the slices overlaps pack + seeding of earlier slices
        W->bktCount = 0; refreshView(W);   // a per-sample decision: the staged and sharded entry points count directly
        if (W->stageTimers) CK(cudaEventRecord(W->ev[2], W->st));
        stageScore(W, *prm);
        launchChain(W->view, nullptr, W->st);
        launchTies(I->view, W->view, makeOpts(*prm, false), W->st);
        if (W->stageTimers) CK(cudaEventRecord(W->ev[6], W->st));
        // result block D2H, then the sample's reset, then ONE host synchronisation: the reset (which touches neither the block nor the tie
        // lists) no longer waits for the host to wake up in between
        // ... and the reset runs on the side stream while the copy engine delivers the block (both only depend on the ties kernel)
        CK(cudaEventRecord(W->evFork, W->st));
        CK(cudaStreamWaitEvent(W->stCopy, W->evFork, 0));
        launchResetSample(I->view, W->view, W->stCopy);
        CK(cudaEventRecord(W->evJoin, W->stCopy));
        enqueueSmall(W);
        CK(cudaStreamWaitEvent(W->st, W->evJoin, 0));
        CK(cudaEventRecord(W->ev[7], W->st));
        CK(cudaStreamSynchronize(W->st));
        parseSmall(W);
        const bool tableTight = (u64)W->hAcc.entries * 10 > W->tableCap * 7;
        if (W->hAcc.overflow || tableTight) {
            // grow and redo: the table (or a list) was too small for this sample
            CK(cudaStreamSynchronize(W->st));
            if (W->hAcc.overflow && !tableTight && attempt >= 2) throw std::runtime_error("internal capacity exceeded");
            // a tight table knows its entry count: size for it directly; an overflowed one only knows "more"
            ensureTable(W, std::max<u64>(W->tableCap * 4, W->hAcc.overflow ? 0 : fitLo() * (u64)W->hAcc.entries / 2));
            W->lastEntries = 0;
            continue;
        }
        fetchTies(W);   // long tie lists only: one more copy + synchronisation (outside the stage timers)
        W->lastEntries = (u64)W->hAcc.entries;
        fillResult(W, res, W->nReads);
        recordStageTimes(W, res);
        W->haveResult = true;
        return PM_OK;
    }
    throw std::runtime_error("read seed table kept overflowing");
}

}  // namespace host
}  // namespace pm

// =====================================================================================================
extern "C" {

const char* pm_last_error(void) { return g_err.c_str(); }
int pm_abi_version(void) { return PM_ABI_VERSION; }
int pm_device_count(void) { return deviceCountNoThrow(); }
uint64_t pm_launch_count(void) { return pm::host::g_launches.load(std::memory_order_relaxed); }

int pm_host_index_read(const char* path, pm_host_index** out) {
    if (!path || !out) return fail(PM_ERR_INVALID, "null argument");
    *out = nullptr;
    return guarded([&]() -> int {
        std::unique_ptr<pm_host_index> h(new pm_host_index());
        readIdxFile(path, h->h);
        *out = h.release();
        return PM_OK;
    });
}
void pm_host_index_free(pm_host_index* h) { delete h; }
int pm_host_index_desc(const pm_host_index* h, pm_index_desc* out) {
    if (!h || !out) return fail(PM_ERR_INVALID, "null argument");
    const HostIndex& x = h->h;
    out->n_nodes = x.parentIndex.size(); out->n_deltas = x.hash.size();
    out->delta_hash = x.hash.data(); out->delta_parent = x.parentCount.data(); out->delta_child = x.childCount.data();
    out->node_offsets = x.nodeOffsets.data(); out->parent_index = x.parentIndex.data(); out->seed = x.sp;
    return PM_OK;
}
const char* pm_host_index_node_id(const pm_host_index* h, uint64_t node) {
    if (!h || node >= h->h.nodeIds.size()) return "";
    return h->h.nodeIds[node].c_str();
}

int pm_host_index_extras(const pm_host_index* hc, pm_index_extras* out) {
    if (!hc || !out) return fail(PM_ERR_INVALID, "null argument");
    pm_host_index* h = const_cast<pm_host_index*>(hc);
    if (h->idPtrs.size() != h->h.nodeIds.size()) { h->idPtrs.clear(); for (const auto& s : h->h.nodeIds) h->idPtrs.push_back(s.c_str()); }
    out->node_ids = h->idPtrs.data();
    out->identical_to_parent = h->h.identicalToParent.empty() ? nullptr : h->h.identicalToParent.data();
    out->block_ranges = h->h.blockRanges.empty() ? nullptr : h->h.blockRanges.data(); out->n_blocks = h->h.blockRanges.size() / 2;
    out->substitution_matrix = h->h.substitutionMatrix.size() == 16 ? h->h.substitutionMatrix.data() : nullptr;
    return PM_OK;
}
int pm_host_index_write(const char* path, const pm_index_desc* desc, const pm_index_extras* extras, int zstd_level, uint64_t* bytes_out) {
    if (!path || !desc) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        IdxExtras x;
        if (extras) { x.nodeIds = extras->node_ids; x.identicalToParent = extras->identical_to_parent; x.blockRanges = extras->block_ranges;
                      x.nBlocks = extras->n_blocks; x.substitutionMatrix = extras->substitution_matrix; }
        const uint64_t n = writeIdxFile(path, *desc, x, zstd_level);
        if (bytes_out) *bytes_out = n;
        return PM_OK;
    });
}

// ---- cached image of the flattened index (pm_image.cpp) ----
int pm_index_image_write(const pm_index_desc* desc, const char* const* node_ids, uint32_t shard, uint32_t n_shards, const char* image_path, uint64_t* bytes_out) {
    if (!desc || !image_path) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        FlatIndex F;
        flattenIndex(*desc, shard, n_shards, F);
        std::vector<std::string> ids;
        if (node_ids) for (uint64_t i = 0; i < desc->n_nodes; ++i) ids.emplace_back(node_ids[i] ? node_ids[i] : "");
        ImageStamp st; st.shard = shard; st.nShards = n_shards;
        const uint64_t n = writeFlatImage(image_path, F, ids, st);
        if (bytes_out) *bytes_out = n;
        return PM_OK;
    });
}
int pm_index_create_from_image(const char* image_path, int device, pm_index** out) {
    if (!image_path || !out) return fail(PM_ERR_INVALID, "null argument");
    *out = nullptr;
    if (deviceCountNoThrow() <= device || device < 0) return noDevice(device);
    return guarded([&]() -> int {
        std::unique_ptr<pm_index> I(new pm_index());
        std::string why;
        if (!readFlatImage(image_path, I->F, I->nodeIds, nullptr, &why)) throw IoError(std::string("cannot use index image ") + image_path + ": " + why);
        uploadIndex(I.get(), device);
        *out = I.release();
        return PM_OK;
    });
}
int pm_index_open_cached(const char* idx_path, const char* image_path, int device, uint32_t shard, uint32_t n_shards, pm_index** out, int* cache_hit) {
    if (!idx_path || !out) return fail(PM_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cache_hit) *cache_hit = 0;
    if (deviceCountNoThrow() <= device || device < 0) return noDevice(device);
    return guarded([&]() -> int {
        std::string img = image_path ? image_path : std::string(idx_path) + ".pmflat";
        if (!image_path && n_shards > 1) img += "." + std::to_string(shard) + "of" + std::to_string(n_shards);
        const ImageStamp want = stampOfFile(idx_path, shard, n_shards);
        std::unique_ptr<pm_index> I(new pm_index());
        std::string why;
        if (readFlatImage(img, I->F, I->nodeIds, &want, &why)) {
            if (cache_hit) *cache_hit = 1;
        } else {
            I.reset(new pm_index());
            HostIndex h;
            readIdxFile(idx_path, h);
            pm_index_desc d{};
            d.n_nodes = h.parentIndex.size(); d.n_deltas = h.hash.size(); d.delta_hash = h.hash.data(); d.delta_parent = h.parentCount.data();
            d.delta_child = h.childCount.data(); d.node_offsets = h.nodeOffsets.data(); d.parent_index = h.parentIndex.data(); d.seed = h.sp;
            flattenIndex(d, shard, n_shards, I->F);
            I->nodeIds = h.nodeIds;
            // a cache that cannot be written (read-only directory, full disk) is not an error: the index still opens
            try { writeFlatImage(img, I->F, I->nodeIds, want); } catch (const std::exception&) {}
        }
        uploadIndex(I.get(), device);
        *out = I.release();
        return PM_OK;
    });
}
const char* pm_index_node_id(const pm_index* idx, uint64_t node) {
    if (!idx || node >= idx->nodeIds.size()) return "";
    return idx->nodeIds[node].c_str();
}

int pm_index_create(const pm_index_desc* desc, int device, pm_index** out) { return createIndex(desc, device, 0, 1, out); }
int pm_index_create_shard(const pm_index_desc* desc, int device, uint32_t shard, uint32_t n_shards, pm_index** out) {
    return createIndex(desc, device, shard, n_shards, out);
}
void pm_index_destroy(pm_index* idx) { if (idx) { cudaSetDevice(idx->device); delete idx; } }
uint64_t pm_index_num_nodes(const pm_index* idx) { return idx ? idx->F.N : 0; }
uint64_t pm_index_num_deltas(const pm_index* idx) { return idx ? idx->F.D : 0; }
uint64_t pm_index_num_distinct_seeds(const pm_index* idx) { return idx ? idx->F.S : 0; }
int pm_index_shard_range(const pm_index* idx, uint64_t* b, uint64_t* e) {
    if (!idx || !b || !e) return fail(PM_ERR_INVALID, "null argument");
    *b = idx->F.nodeBegin; *e = idx->F.nodeEnd; return PM_OK;
}
int pm_index_genome_metrics(const pm_index* idx, double* magSq, int64_t* uniq) {
    if (!idx) return fail(PM_ERR_INVALID, "null argument");
    if (magSq) std::memcpy(magSq, idx->gMagSqHost.data(), idx->gMagSqHost.size() * sizeof(double));
    if (uniq) std::memcpy(uniq, idx->gUniqueHost.data(), idx->gUniqueHost.size() * sizeof(int64_t));
    return PM_OK;
}
int pm_index_bfs_ranks(const pm_index* idx, uint32_t* out) {
    if (!idx || !out) return fail(PM_ERR_INVALID, "null argument");
    std::memcpy(out, idx->F.bfsRank.data(), idx->F.bfsRank.size() * sizeof(uint32_t));
    return PM_OK;
}

int pm_workspace_create(pm_index* idx, pm_workspace** out) {
    if (!idx || !out) return fail(PM_ERR_INVALID, "null argument");
    *out = nullptr;
    return guarded([&]() -> int {
        setDevice(idx->device);
        std::unique_ptr<pm_workspace> W(new pm_workspace());
        W->idx = idx; W->device = idx->device;
        CK(cudaStreamCreateWithFlags(&W->st, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&W->stCopy, cudaStreamNonBlocking));
        for (auto& e : W->ev) CK(cudaEventCreate(&e));
        for (auto& e : W->evCopy) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));   // one per slice of the host-buffer pipeline
        for (auto& e : W->evK) CK(cudaEventCreate(&e));
        W->stageTimers = stageTimersDefault();
        CK(cudaEventCreateWithFlags(&W->evFork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&W->evJoin, cudaEventDisableTiming));
        const FlatIndex& F = idx->F;
        const DevIndexView& V = idx->view;
        W->resultBlob.alloc(kResultBlobBytes);
        CK(cudaMemsetAsync(W->resultBlob.p, 0, kResultBlobBytes, W->st));
        W->acc.p = reinterpret_cast<SampleAcc*>(W->resultBlob.p);
        W->scalars.p = reinterpret_cast<SampleScalars*>(W->resultBlob.p + sizeof(SampleAcc));
        W->sel.p = reinterpret_cast<Selection*>(W->resultBlob.p + sizeof(SampleAcc) + sizeof(SampleScalars));
        W->tieHead = reinterpret_cast<u32*>(W->resultBlob.p + sizeof(SampleAcc) + sizeof(SampleScalars) + 5 * sizeof(Selection));
        W->ell.alloc(F.S + 2);
        {
            if (F.S + 2 >= (1ull << 27)) throw Unsupported("more than 2^27 distinct seeds: linear texture limit of the ell table");
            cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = W->ell.p;
            rd.res.linear.desc = cudaCreateChannelDesc<int2>(); rd.res.linear.sizeInBytes = (F.S + 2) * sizeof(long long);
            cudaTextureDesc td{}; td.readMode = cudaReadModeElementType;
            CK(cudaCreateTextureObject(&W->ellTex, &rd, &td, nullptr));
        } CK(cudaMemsetAsync(W->ell.p, 0, W->ell.n * sizeof(long long), W->st));   // [S] is the always-zero slot of the padding words
        W->scanPart.alloc(kMaxPartials); W->finPart.alloc(kMaxPartials);
        W->countHist.alloc(kLog1pLut); CK(cudaMemset(W->countHist.p, 0, kLog1pLut * sizeof(unsigned)));   // kept zero between samples by its consumer (root_and_scalars)
        W->segRec.alloc(F.nSeg + 1); CK(cudaMemsetAsync(W->segRec.p, 0, W->segRec.n * sizeof(SegRec), W->st));
        W->chainA.alloc(V.chainTotal ? V.chainTotal : 1);
        W->genRec.alloc((size_t)(F.nGenNodes ? F.nGenNodes : 1) * kGenWords); W->evPrefix.alloc((size_t)(V.nEvents ? V.nEvents : 1) * kGenWords);
        W->scores.alloc(F.N * 5); CK(cudaMemsetAsync(W->scores.p, 0, F.N * 5 * sizeof(double), W->st));
        W->blockMaxAndBfs.alloc((size_t)V.nBfsBlocks * 5 + (size_t)V.nShardNodes * 5 + 8);
        W->recCap = V.nShardNodes ? V.nShardNodes : 1;
        W->recRank.alloc((size_t)5 * W->recCap); W->recNode.alloc((size_t)5 * W->recCap); W->recScore.alloc((size_t)5 * W->recCap);
        W->tieCap = V.nShardNodes ? V.nShardNodes : 1;
        W->tieNode.alloc((size_t)5 * W->tieCap);
        W->expCounter.alloc(1);
        CK(cudaMemsetAsync(W->acc.p, 0, sizeof(SampleAcc), W->st));
        CK(cudaStreamSynchronize(W->st));
        refreshView(W.get());
        workspaceRegister(W.get(), true);
        *out = W.release();
        return PM_OK;
    });
}
void pm_workspace_destroy(pm_workspace* ws) {
    if (!ws) return;
    workspaceRegister(ws, false);
    cudaSetDevice(ws->device);
    if (ws->st) { cudaStreamSynchronize(ws->st); cudaStreamDestroy(ws->st); }
    if (ws->stCopy) { cudaStreamSynchronize(ws->stCopy); cudaStreamDestroy(ws->stCopy); }
    if (ws->ellTex) cudaDestroyTextureObject(ws->ellTex);
    if (ws->tableTex) cudaDestroyTextureObject(ws->tableTex);
    for (auto& e : ws->ev) if (e) cudaEventDestroy(e);
    for (auto& e : ws->evCopy) if (e) cudaEventDestroy(e);
    for (auto& e : ws->evK) if (e) cudaEventDestroy(e);
    if (ws->evFork) cudaEventDestroy(ws->evFork);
    if (ws->evJoin) cudaEventDestroy(ws->evJoin);
    delete ws;
}

int pm_place(pm_workspace* ws, const char* reads, const uint64_t* read_offsets, uint64_t n_reads, const pm_place_params* params,
             pm_place_result* result) {
    if (!ws || !read_offsets || (!reads && n_reads)) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int { return runPlace(ws, params, result, false, reads, read_offsets, n_reads); });
}
int pm_place_packed(pm_workspace* ws, const void* packed, const uint64_t* read_offsets, uint64_t n_reads, const pm_place_params* params,
                    pm_place_result* result) {
    if (!ws || !read_offsets || (!packed && n_reads)) return fail(PM_ERR_INVALID, "null argument");
    if ((reinterpret_cast<uintptr_t>(packed) & 15u) != 0) return fail(PM_ERR_INVALID, "packed reads must be 16-byte aligned");
    if (params && params->min_seed_quality > 0) return fail(PM_ERR_UNSUPPORTED, "min_seed_quality needs the ASCII reads and qualities: pm_place_quality");
    return guarded([&]() -> int { return runPlace(ws, params, result, false, nullptr, read_offsets, n_reads, static_cast<const uint4*>(packed)); });
}
int pm_place_quality(pm_workspace* ws, const char* reads, const char* quals, const uint64_t* read_offsets, uint64_t n_reads,
                     const pm_place_params* params, pm_place_result* result) {
    if (!params || params->min_seed_quality <= 0) return pm_place(ws, reads, read_offsets, n_reads, params, result);
    if (!ws || !read_offsets || ((!reads || !quals) && n_reads)) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        setDevice(ws->idx->device);
        uploadReads(ws, reads, read_offsets, n_reads);   // an optional mode: plain upload, then the resident stages
        const u64 total = n_reads ? read_offsets[n_reads] : 0;
        ws->quals.ensure(total + 64);
        if (total) CK(cudaMemcpyAsync(ws->quals.p, quals, total, cudaMemcpyHostToDevice, ws->st));
        ws->residentValid = false;   // hpc indexes compress reads and qualities in place: not reusable by pm_place_resident
        ws->useQuals = true;
        struct Reset { pm_workspace* w; ~Reset() { w->useQuals = false; } } reset{ws};
        return runPlace(ws, params, result, true, nullptr, nullptr, 0);
    });
}
int pm_reads_upload(pm_workspace* ws, const char* reads, const uint64_t* read_offsets, uint64_t n_reads) {
    if (!ws || !read_offsets || (!reads && n_reads)) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        setDevice(ws->idx->device);
        uploadReads(ws, reads, read_offsets, n_reads);
        CK(cudaStreamSynchronize(ws->st));
        ws->residentValid = true;
        return PM_OK;
    });
}
int pm_reads_upload_device(pm_workspace* ws, const char* d_reads, const uint64_t* d_read_offsets, const uint64_t* h_read_offsets, uint64_t n_reads) {
    if (!ws || !h_read_offsets || !d_read_offsets || (!d_reads && n_reads)) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        setDevice(ws->idx->device);
        if (ws->uploadPending) CK(cudaStreamSynchronize(ws->st));   // the pinned staging arrays of the previous upload may still be in flight
        uploadReads(ws, d_reads, h_read_offsets, n_reads, true, d_read_offsets);   // stream-ordered: no host wait
        ws->residentValid = true; ws->uploadPending = true;
        return PM_OK;
    });
}
int pm_place_resident(pm_workspace* ws, const pm_place_params* params, pm_place_result* result) {
    if (!ws) return fail(PM_ERR_INVALID, "null argument");
    if (!ws->residentValid) return fail(PM_ERR_INVALID, "pm_place_resident: call pm_reads_upload first");
    return guarded([&]() -> int { return runPlace(ws, params, result, true, nullptr, nullptr, 0); });
}

int pm_get_tied(pm_workspace* ws, int metric, uint32_t* out, uint64_t cap) {
    if (!ws || !ws->haveResult || metric < 0 || metric >= 5) return fail(PM_ERR_INVALID, "no result / bad metric");
    const auto& t = ws->tied[metric];
    if (out) std::memcpy(out, t.data(), std::min<uint64_t>(cap, t.size()) * sizeof(uint32_t));
    return PM_OK;
}
int pm_get_node_scores(pm_workspace* ws, double* out) {
    if (!ws || !out || !ws->haveResult) return fail(PM_ERR_INVALID, "no result");
    return guarded([&]() -> int {
        setDevice(ws->idx->device);
        CK(cudaMemcpyAsync(out, ws->scores.p, ws->idx->F.N * 5 * sizeof(double), cudaMemcpyDeviceToHost, ws->st));
        CK(cudaStreamSynchronize(ws->st));
        return PM_OK;
    });
}
// re-runs the scoring stages of the last sample with the numerator dump enabled (debug / parity tests)
int pm_get_node_metrics(pm_workspace* ws, double* out) {
    if (!ws || !out || !ws->haveResult) return fail(PM_ERR_INVALID, "no result");
    return guarded([&]() -> int {
        pm_index* I = ws->idx;
        setDevice(I->device);
        ws->metrics.ensure(I->F.N * 5);
        CK(cudaMemsetAsync(ws->metrics.p, 0, I->F.N * 5 * sizeof(double), ws->st));
        ws->wantMetrics = true; refreshView(ws);
        // ell was reset after the sample: rebuild it from the (still intact) table, then K1 + K2 only
        CK(cudaMemsetAsync(&ws->acc.p->entCount, 0, sizeof(unsigned), ws->st));
        CK(cudaMemsetAsync(ws->acc.p->magSq, 0, 6 * sizeof(u64) + 7 * sizeof(long long), ws->st));
        const PlaceOpts O = makeOpts(ws->lastParams, true);
        launchFinalize(I->view, ws->view, O, I->homo.p, ws->lastEntries ? ws->lastEntries : ws->tableCap / 4, I->nSM, ws->st);
        launchDeltas(I->view, ws->view, I->nSM, ws->st);
        launchGeneral(I->view, ws->view, ws->st);
        launchPrefixScores(I->view, ws->view, O, ws->st);
        launchResetSample(ws->idx->view, ws->view, ws->st);
        CK(cudaMemcpyAsync(out, ws->metrics.p, I->F.N * 5 * sizeof(double), cudaMemcpyDeviceToHost, ws->st));
        CK(cudaStreamSynchronize(ws->st));
        ws->wantMetrics = false; refreshView(ws);
        return PM_OK;
    });
}
int pm_get_seed_table(pm_workspace* ws, uint64_t* hash, int64_t* count, uint64_t cap) {
    if (!ws || !hash || !count) return fail(PM_ERR_INVALID, "null argument");
    int n = 0;
    const int rc = guarded([&]() -> int {
        setDevice(ws->idx->device);
        ws->expHash.ensure(cap ? cap : 1); ws->expCount.ensure(cap ? cap : 1);
        CK(cudaMemsetAsync(ws->expCounter.p, 0, sizeof(unsigned), ws->st));
        launchTableExport(ws->view, ws->expHash.p, ws->expCount.p, ws->expCounter.p, cap, ws->st);
        unsigned cnt = 0;
        CK(cudaMemcpyAsync(&cnt, ws->expCounter.p, sizeof(unsigned), cudaMemcpyDeviceToHost, ws->st));
        CK(cudaStreamSynchronize(ws->st));
        const u64 m = std::min<u64>(cnt, cap);
        if (m) {
            CK(cudaMemcpyAsync(hash, ws->expHash.p, m * sizeof(u64), cudaMemcpyDeviceToHost, ws->st));
            CK(cudaMemcpyAsync(count, ws->expCount.p, m * sizeof(long long), cudaMemcpyDeviceToHost, ws->st));
            CK(cudaStreamSynchronize(ws->st));
        }
        n = (int)cnt;
        return PM_OK;
    });
    return rc == PM_OK ? n : rc;
}

// ---- seeding::rollingSyncmers / per-read seeds on the GPU (parity + drop-in for the host shim) ----
int seedListImpl(int device, const char* seqs, const uint64_t* off, uint64_t n, const pm_seed_params* sp, int trimStart,
                        int trimEnd, int mode, uint64_t* outHash, uint8_t* outRev, int64_t* outPos, uint64_t* outCount) {
    if (!off || (!seqs && n) || !sp || !outHash || !outCount) return fail(PM_ERR_INVALID, "null argument");
    if (deviceCountNoThrow() <= device || device < 0) return fail(PM_ERR_NO_DEVICE, "no usable CUDA device (this library has no CPU fallback)");
    return guarded([&]() -> int {
        if (sp->k < 1 || sp->k > kMaxK || sp->s < 1 || sp->s > sp->k || sp->t < 0 || sp->t > sp->k - sp->s)
            throw std::runtime_error("unsupported seeding parameters (need 1 <= s <= k <= 32, 0 <= t <= k-s)");
        setDevice(device);
        cudaStream_t st; CK(cudaStreamCreate(&st));
        std::vector<u64> pOff(n + 1), wOff(n + 1);
        u64 ch = 0, win = 0;
        for (u64 i = 0; i < n; ++i) {
            const u64 L = off[i + 1] - off[i];
            pOff[i] = ch; wOff[i] = win;
            ch += (L + 31) >> 5;
            if (L >= (u64)sp->k) win += L - (u64)sp->k + 1;
        }
        pOff[n] = ch; wOff[n] = win;
        const u64 total = n ? off[n] : 0;
        DevBuf<char> dReads; DevBuf<u64> dOff, dPOff, dWOff, dHash, dCount; DevBuf<uint4> dPacked;
        DevBuf<unsigned char> dRev; DevBuf<long long> dPos; DevBuf<SeedTables> dT; DevBuf<u64> dSyn; DevBuf<unsigned> dSynCount;
        dReads.alloc(total + 64); dOff.alloc(n + 1); dPOff.alloc(n + 1); dWOff.alloc(n + 1); dPacked.alloc(ch + 1);
        dHash.alloc(win + 1); dCount.alloc(n + 1);
        if (mode == 1) { dRev.alloc(win + 1); dPos.alloc(win + 1); } else { dSyn.alloc(ch * 32 + 32); dSynCount.alloc(n + 1); }
        if (total) CK(cudaMemcpyAsync(dReads.p, seqs, total, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dOff.p, off, (n + 1) * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dPOff.p, pOff.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dWOff.p, wOff.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
        std::vector<SeedTables> T(kSeedTableElems); buildSeedTableImage(T.data(), sp->k, sp->s);
        dT.alloc(kSeedTableElems); CK(cudaMemcpyAsync(dT.p, T.data(), kSeedTableElems * sizeof(SeedTables), cudaMemcpyHostToDevice, st));
        const SeederParams P = makeSeederParams(sp->k, sp->s, sp->t, sp->l, sp->open, trimStart, trimEnd);
        std::vector<u32> bf((ch + 255) / 256 + 1);
        packBlockFirst(pOff.data(), n, ch, bf.data());
        DevBuf<u32> dBF; dBF.alloc(bf.size());
        CK(cudaMemcpyAsync(dBF.p, bf.data(), bf.size() * sizeof(u32), cudaMemcpyHostToDevice, st));
        launchPackReads(dReads.p, dOff.p, dPOff.p, dBF.p, n, 0, ch, dPacked.p, st);
        launchSeedList(dPacked.p, dOff.p, dPOff.p, dWOff.p, n, P, dT.p, mode, dSyn.p, dSynCount.p, dHash.p, dRev.p, dPos.p, dCount.p, st);
        CK(cudaGetLastError());
        if (win) CK(cudaMemcpyAsync(outHash, dHash.p, win * 8, cudaMemcpyDeviceToHost, st));
        if (mode == 1 && win) {
            if (outRev) CK(cudaMemcpyAsync(outRev, dRev.p, win, cudaMemcpyDeviceToHost, st));
            if (outPos) CK(cudaMemcpyAsync(outPos, dPos.p, win * 8, cudaMemcpyDeviceToHost, st));
        }
        if (n) CK(cudaMemcpyAsync(outCount, dCount.p, n * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        cudaStreamDestroy(st);
        return PM_OK;
    });
}
// seeding::hashSeq for a batch (seeding.cpp:20-30); a sequence with a non-ACGT base makes the call fail like the reference's std::invalid_argument
int pm_hash_seq(int device, const char* seqs, const uint64_t* seq_offsets, uint64_t n_seqs, uint64_t* out_fwd, uint64_t* out_rev) {
    if (!seq_offsets || (!seqs && n_seqs) || !out_fwd || !out_rev) return fail(PM_ERR_INVALID, "null argument");
    if (deviceCountNoThrow() <= device || device < 0) return fail(PM_ERR_NO_DEVICE, "no usable CUDA device (this library has no CPU fallback)");
    return guarded([&]() -> int {
        setDevice(device);
        const u64 total = n_seqs ? seq_offsets[n_seqs] : 0;
        DevBuf<char> dS; DevBuf<u64> dO, dF, dR; DevBuf<unsigned char> dB;
        dS.alloc(total + 1); dO.alloc(n_seqs + 1); dF.alloc(n_seqs + 1); dR.alloc(n_seqs + 1); dB.alloc(n_seqs + 1);
        if (total) CK(cudaMemcpy(dS.p, seqs, total, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dO.p, seq_offsets, (n_seqs + 1) * 8, cudaMemcpyHostToDevice));
        launchHashSeq(dS.p, dO.p, n_seqs, dF.p, dR.p, dB.p, 0);
        CK(cudaGetLastError());
        std::vector<unsigned char> bad(n_seqs + 1);
        if (n_seqs) {
            CK(cudaMemcpy(out_fwd, dF.p, n_seqs * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(out_rev, dR.p, n_seqs * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(bad.data(), dB.p, n_seqs, cudaMemcpyDeviceToHost));
        }
        for (u64 i = 0; i < n_seqs; ++i) if (bad[i]) throw std::invalid_argument("Kmer contains non canonical base");
        return PM_OK;
    });
}
int pm_rolling_syncmers(int device, const char* seqs, const uint64_t* seq_offsets, uint64_t n_seqs, int k, int s, int open, int t,
                        uint64_t* out_hash, uint8_t* out_is_reverse, int64_t* out_pos, uint64_t* out_count) {
    pm_seed_params sp{k, s, t, 1, open, 0};
    return seedListImpl(device, seqs, seq_offsets, n_seqs, &sp, 0, 0, 1, out_hash, out_is_reverse, out_pos, out_count);
}
int pm_read_seeds(int device, const char* seqs, const uint64_t* seq_offsets, uint64_t n_seqs, const pm_seed_params* sp,
                  int trim_start, int trim_end, uint64_t* out_hash, uint64_t* out_count) {
    return seedListImpl(device, seqs, seq_offsets, n_seqs, sp, trim_start, trim_end, 2, out_hash, nullptr, nullptr, out_count);
}

// ---- staged entry points (multi-GPU) ----
static int stageSeedImpl(pm_workspace* ws, const char* reads, const uint64_t* read_offsets, uint64_t n_reads, const pm_place_params* params,
                         bool resident) {
    return guarded([&]() -> int {
        setDevice(ws->idx->device);
        checkParams(params);
        ws->lastParams = *params; ws->wantMetrics = false; ws->haveResult = false;
        if (!resident) uploadReads(ws, reads, read_offsets, n_reads);
        if (ws->tableCap == 0) ensureTable(ws, std::max<u64>(1 << 16, ws->totalWindows / 4));
        refreshView(ws);
        for (int attempt = 0; attempt < 6; ++attempt) {
            stageSeed(ws, true, *params);
            SampleAcc a;
            CK(cudaMemcpyAsync(&a, ws->acc.p, sizeof(a), cudaMemcpyDeviceToHost, ws->st));
            CK(cudaStreamSynchronize(ws->st));
            if (!a.overflow) return PM_OK;
            ensureTable(ws, ws->tableCap * 4);   // grow and reseed
        }
        throw std::runtime_error("read seed table kept overflowing");
    });
}
int pm_stage_seed(pm_workspace* ws, const char* reads, const uint64_t* read_offsets, uint64_t n_reads, const pm_place_params* params) {
    if (!ws || !read_offsets || (!reads && n_reads)) return fail(PM_ERR_INVALID, "null argument");
    return stageSeedImpl(ws, reads, read_offsets, n_reads, params, false);
}
// same, on the reads previously laid out by pm_reads_upload ("inputs resident in HBM")
int pm_stage_seed_resident(pm_workspace* ws, const pm_place_params* params) {
    if (!ws) return fail(PM_ERR_INVALID, "null argument");
    if (!ws->residentValid) return fail(PM_ERR_INVALID, "pm_stage_seed_resident: call pm_reads_upload first");
    return stageSeedImpl(ws, nullptr, nullptr, 0, params, true);
}
int64_t pm_stage_table_size(pm_workspace* ws) {
    if (!ws) return fail(PM_ERR_INVALID, "null argument");
    int64_t n = 0;
    const int rc = guarded([&]() -> int {
        setDevice(ws->idx->device);
        CK(cudaMemsetAsync(ws->expCounter.p, 0, sizeof(unsigned), ws->st));
        launchTableExport(ws->view, nullptr, nullptr, ws->expCounter.p, 0, ws->st);
        unsigned cnt = 0;
        CK(cudaMemcpyAsync(&cnt, ws->expCounter.p, sizeof(unsigned), cudaMemcpyDeviceToHost, ws->st));
        CK(cudaStreamSynchronize(ws->st));
        n = cnt; return PM_OK;
    });
    return rc == PM_OK ? n : rc;
}
int pm_stage_table_export(pm_workspace* ws, uint64_t* hash, int64_t* count, uint64_t cap) {
    const int n = pm_get_seed_table(ws, hash, count, cap);
    return n < 0 ? n : PM_OK;
}
// replaces this rank's table content by the union of all ranks' exports (the caller passes the concatenation)
int pm_stage_table_import(pm_workspace* ws, const uint64_t* hash, const int64_t* count, uint64_t n) {
    if (!ws || (n && (!hash || !count))) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        setDevice(ws->idx->device);
        for (int attempt = 0; attempt < 6; ++attempt) {
            if (ws->tableCap < 2 * n + 16) ensureTable(ws, 2 * n + 16);
            refreshView(ws);
            ws->expHash.ensure(n ? n : 1); ws->expCount.ensure(n ? n : 1);
            if (n) {
                CK(cudaMemcpyAsync(ws->expHash.p, hash, n * 8, cudaMemcpyHostToDevice, ws->st));
                CK(cudaMemcpyAsync(ws->expCount.p, count, n * 8, cudaMemcpyHostToDevice, ws->st));
            }
            CK(cudaMemsetAsync(ws->acc.p, 0, sizeof(SampleAcc), ws->st));
            launchTableClear(ws->view, ws->st);
            launchTableImport(ws->view, ws->expHash.p, ws->expCount.p, n, ws->st);
            CK(cudaStreamSynchronize(ws->st));
            SampleAcc a; CK(cudaMemcpy(&a, ws->acc.p, sizeof(a), cudaMemcpyDeviceToHost));
            if (!a.overflow) return PM_OK;
            ensureTable(ws, ws->tableCap * 4);
        }
        throw std::runtime_error("table import kept overflowing");
    });
}
// device-pointer variants of the table exchange (the caller owns the buffers, e.g. torch CUDA tensors handed to NCCL)
int pm_stage_table_export_dev(pm_workspace* ws, uint64_t* d_hash, int64_t* d_count, uint64_t cap, uint64_t* n_out) {
    if (!ws || !n_out) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        setDevice(ws->idx->device);
        CK(cudaMemsetAsync(ws->expCounter.p, 0, sizeof(unsigned), ws->st));
        launchTableExport(ws->view, reinterpret_cast<u64*>(d_hash), reinterpret_cast<long long*>(d_count), ws->expCounter.p, d_hash ? cap : 0, ws->st);
        unsigned cnt = 0;
        CK(cudaMemcpyAsync(&cnt, ws->expCounter.p, sizeof(unsigned), cudaMemcpyDeviceToHost, ws->st));
        CK(cudaStreamSynchronize(ws->st));
        *n_out = cnt;
        if (d_count && cnt < cap) {   // unused entries get count 0 so that importers skip them
            CK(cudaMemsetAsync(d_count + cnt, 0, (cap - cnt) * sizeof(int64_t), ws->st));
            CK(cudaStreamSynchronize(ws->st));
        }
        return PM_OK;
    });
}
int pm_stage_table_import_dev(pm_workspace* ws, const uint64_t* d_hash, const int64_t* d_count, uint64_t n) {
    if (!ws || (n && (!d_hash || !d_count))) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        setDevice(ws->idx->device);
        for (int attempt = 0; attempt < 6; ++attempt) {
            if (ws->tableCap < 2 * n + 16) ensureTable(ws, 2 * n + 16);
            refreshView(ws);
            CK(cudaMemsetAsync(ws->acc.p, 0, sizeof(SampleAcc), ws->st));
            launchTableClear(ws->view, ws->st);
            launchTableImport(ws->view, reinterpret_cast<const u64*>(d_hash), reinterpret_cast<const long long*>(d_count), n, ws->st);
            SampleAcc a;
            CK(cudaMemcpyAsync(&a, ws->acc.p, sizeof(a), cudaMemcpyDeviceToHost, ws->st));
            CK(cudaStreamSynchronize(ws->st));
            if (!a.overflow) return PM_OK;
            ensureTable(ws, ws->tableCap * 4);
        }
        throw std::runtime_error("table import kept overflowing");
    });
}
// enqueue-only variant: clears the table and inserts the pairs without waiting; an overflow is reported by pm_stage_score
int pm_stage_table_import_dev_async(pm_workspace* ws, const uint64_t* d_hash, const int64_t* d_count, uint64_t n, int clear_first,
                                    uint64_t expected_total) {
    if (!ws || (n && (!d_hash || !d_count))) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        setDevice(ws->idx->device);
        if (clear_first) {
            if (ws->tableCap < 2 * expected_total + 16) ensureTable(ws, 2 * expected_total + 16);
            refreshView(ws);
            CK(cudaMemsetAsync(ws->acc.p, 0, sizeof(SampleAcc), ws->st));
            launchTableClear(ws->view, ws->st);
        }
        launchTableImport(ws->view, reinterpret_cast<const u64*>(d_hash), reinterpret_cast<const long long*>(d_count), n, ws->st);
        return PM_OK;
    });
}
int pm_stage_score(pm_workspace* ws, const pm_place_params* params) {
    if (!ws) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        setDevice(ws->idx->device);
        checkParams(params);
        ws->lastParams = *params;
        refreshView(ws);
        stageScore(ws, *params);
        fetchSmall(ws);
        if (ws->hAcc.overflow) throw std::runtime_error("internal capacity exceeded while scoring");
        return PM_OK;
    });
}
int64_t pm_stage_records_size(pm_workspace* ws, int metric) {
    if (!ws || metric < 0 || metric >= 5) return fail(PM_ERR_INVALID, "bad argument");
    return std::min<unsigned>(ws->hAcc.recordCount[metric], ws->recCap);
}
int pm_stage_records_export(pm_workspace* ws, int metric, uint32_t* bfs_rank, uint32_t* node, double* score, uint64_t cap) {
    if (!ws || metric < 0 || metric >= 5) return fail(PM_ERR_INVALID, "bad argument");
    return guarded([&]() -> int {
        setDevice(ws->idx->device);
        const u64 n = std::min<u64>(std::min<unsigned>(ws->hAcc.recordCount[metric], ws->recCap), cap);
        if (n) {
            CK(cudaMemcpyAsync(bfs_rank, ws->recRank.p + (size_t)metric * ws->recCap, n * 4, cudaMemcpyDeviceToHost, ws->st));
            CK(cudaMemcpyAsync(node, ws->recNode.p + (size_t)metric * ws->recCap, n * 4, cudaMemcpyDeviceToHost, ws->st));
            CK(cudaMemcpyAsync(score, ws->recScore.p + (size_t)metric * ws->recCap, n * 8, cudaMemcpyDeviceToHost, ws->st));
            CK(cudaStreamSynchronize(ws->st));
        }
        return PM_OK;
    });
}
// all five record lists with one synchronisation: arrays are [5][cap], counts[5] receives the list lengths
int pm_stage_records_export_all(pm_workspace* ws, uint32_t* counts, uint32_t* bfs_rank, uint32_t* node, double* score, uint64_t cap) {
    if (!ws || !counts) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        setDevice(ws->idx->device);
        size_t tot = 0; u64 n[5];
        for (int m = 0; m < 5; ++m) { n[m] = std::min<u64>(std::min<unsigned>(ws->hAcc.recordCount[m], ws->recCap), cap); counts[m] = (uint32_t)n[m]; tot += n[m]; }
        ws->hRec.ensure(tot * 16 + 16);
        unsigned char* h = ws->hRec.p;
        size_t o = 0;
        for (int m = 0; m < 5; ++m) {
            if (!n[m]) continue;
            CK(cudaMemcpyAsync(h + o, ws->recScore.p + (size_t)m * ws->recCap, n[m] * 8, cudaMemcpyDeviceToHost, ws->st)); o += n[m] * 8;
            CK(cudaMemcpyAsync(h + o, ws->recRank.p + (size_t)m * ws->recCap, n[m] * 4, cudaMemcpyDeviceToHost, ws->st)); o += n[m] * 4;
            CK(cudaMemcpyAsync(h + o, ws->recNode.p + (size_t)m * ws->recCap, n[m] * 4, cudaMemcpyDeviceToHost, ws->st)); o += n[m] * 4;
        }
        if (tot) CK(cudaStreamSynchronize(ws->st));
        o = 0;
        for (int m = 0; m < 5; ++m) {
            if (!n[m]) continue;
            std::memcpy(score + (size_t)m * cap, h + o, n[m] * 8); o += n[m] * 8;
            std::memcpy(bfs_rank + (size_t)m * cap, h + o, n[m] * 4); o += n[m] * 4;
            std::memcpy(node + (size_t)m * cap, h + o, n[m] * 4); o += n[m] * 4;
        }
        return PM_OK;
    });
}
// all ranks' records of every metric (concatenated per metric; counts[5]) -> chain on the device -> this shard's ties
int pm_stage_select(pm_workspace* ws, const uint32_t* counts, const uint32_t* const* bfs_rank, const uint32_t* const* node,
                    const double* const* score, uint64_t total_reads, pm_place_result* result) {
    if (!ws || !counts || !bfs_rank || !node || !score || !result) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        pm_index* I = ws->idx;
        setDevice(I->device);
        u32 maxN = 1;
        for (int m = 0; m < 5; ++m) maxN = std::max(maxN, counts[m]);
        if (maxN > ws->recCap) {
            ws->recCap = maxN;
            ws->recRank.alloc((size_t)5 * maxN); ws->recNode.alloc((size_t)5 * maxN); ws->recScore.alloc((size_t)5 * maxN);
            refreshView(ws);
        }
        size_t tot = 0;
        for (int m = 0; m < 5; ++m) tot += counts[m];
        ws->hRec.ensure(tot * 16 + 64);
        unsigned char* h = ws->hRec.p;
        std::memcpy(h, counts, 5 * sizeof(u32));
        if (!ws->selCounts.p) ws->selCounts.alloc(8);
        CK(cudaMemcpyAsync(ws->selCounts.p, h, 5 * sizeof(u32), cudaMemcpyHostToDevice, ws->st));
        size_t o = 32;
        for (int m = 0; m < 5; ++m) {
            if (!counts[m]) continue;
            const size_t n = counts[m];
            std::memcpy(h + o, score[m], n * 8);
            CK(cudaMemcpyAsync(ws->recScore.p + (size_t)m * ws->recCap, h + o, n * 8, cudaMemcpyHostToDevice, ws->st)); o += n * 8;
            std::memcpy(h + o, bfs_rank[m], n * 4);
            CK(cudaMemcpyAsync(ws->recRank.p + (size_t)m * ws->recCap, h + o, n * 4, cudaMemcpyHostToDevice, ws->st)); o += n * 4;
            std::memcpy(h + o, node[m], n * 4);
            CK(cudaMemcpyAsync(ws->recNode.p + (size_t)m * ws->recCap, h + o, n * 4, cudaMemcpyHostToDevice, ws->st)); o += n * 4;
        }
        DevBuf<u32>& dCounts = ws->selCounts;
        launchChain(ws->view, dCounts.p, ws->st);
        launchTies(I->view, ws->view, makeOpts(ws->lastParams, false), ws->st);
        fetchSmall(ws);
        launchResetSample(ws->idx->view, ws->view, ws->st);
        fetchTies(ws);  // local ties + the best node (every rank adds it; the caller unions the lists)
        std::memset(result, 0, sizeof(*result));
        fillResult(ws, result, total_reads);
        ws->haveResult = true;
        return PM_OK;
    });
}

int pm_workspace_set_stage_timers(pm_workspace* ws, int on) {
    if (!ws) return fail(PM_ERR_INVALID, "null workspace");
    ws->stageTimers = on != 0;
    return PM_OK;
}
int pm_last_kernel_ms(pm_workspace* ws, float* out) {
    if (!ws || !out) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        setDevice(ws->idx->device);
        CK(cudaStreamSynchronize(ws->st));
        for (int i = 0; i < 3; ++i) { out[i] = 0.f; if (ws->stageTimers && cudaEventElapsedTime(&out[i], ws->evK[i], ws->evK[i + 1]) != cudaSuccess) { cudaGetLastError(); out[i] = 0.f; } }
        return PM_OK;
    });
}

int pm_workspace_staging(pm_workspace* ws, uint64_t read_bytes, uint64_t n_reads, int want_quals, char** reads_out, uint64_t** offsets_out, char** quals_out) {
    if (!ws || !reads_out || !offsets_out) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        setDevice(ws->device);
        ws->hIngestReads.ensure(read_bytes + 64); ws->hIngestOff.ensure(n_reads + 1);
        if (want_quals) ws->hIngestQuals.ensure(read_bytes + 64);
        *reads_out = ws->hIngestReads.p; *offsets_out = ws->hIngestOff.p;
        if (quals_out) *quals_out = want_quals ? ws->hIngestQuals.p : nullptr;
        return PM_OK;
    });
}

void* pm_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void pm_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
