// pm_multi.cu -- ONE sample placed by several GPUs together (include/panmap_b200.h, "pm_comm").
//
// Data plane (kernels: pm_shard_kernels.cu).  Rank r owns shard r of the node range and seeds its own slice of the reads:
//   P0  seed the slice into the local table                      partition_export -> one segment per owner rank
//   X0  all-to-all of the segments                               (hash-partitioned seed table: counts of a seed meet at its owner)
//   P1  import into the partition table, table_scan, dictionary  partition_finalize -> statistics + (count, seed id) pairs
//   X1  all-gather of the pairs
//   P2  gathered_finalize (ell, magnitudes, scalars), node_deltas / prefix_scores / records on the own node range, records_pack
//   X2  all-gather of the records
//   P3  tolerance chain over all records (same on every rank), local ties, ties_pack
//   X3  all-gather of the tie heads + flags + sizing feedback
//   P4  ONE device-to-host copy, reset
// Nothing between P0 and P4 waits for the host: capacities are fixed per sample, fill counts travel in band, an overflow anywhere
// reaches every rank with X3 and all ranks then redo the sample with the sizes the counts ask for.
// Transports.  One process per GPU: the ranks meet through NCCL (libnccl.so.2 resolved with dlopen so that the library loads on machines
// without it), which carries the bootstrap -- sizing agreement and the CUDA IPC handles of every rank's receive buffers -- and, when the
// buffers cannot be mapped, the exchanges themselves.  Once mapped, X0..X3 are plain 16-byte stores into the peers' buffers over NVLink
// plus one flag word per peer (push_segments / wait_flags, pm_shard_kernels.cu): these payloads are a few megabytes at most, and four
// NCCL collectives cost ~35-50 us each at 8 ranks where a store-and-flag exchange costs a launch.  PM_PEER_EXCHANGE=0 keeps NCCL.
// One process, one host thread driving all ranks (tests, single-box tools): peer copies ordered by events.
#include "pm_internal.h"

#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

using namespace pm;
using namespace pm::host;

namespace {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};
NcclApi& ncclApi() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        // a copy that the process has loaded already (e.g. the one bundled with PyTorch) is found by its soname first
        a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) return;
#define PM_SYM(name) a.name = reinterpret_cast<decltype(a.name)>(dlsym(a.lib, "nccl" #name))
        PM_SYM(GetUniqueId); PM_SYM(CommInitRank); PM_SYM(CommDestroy); PM_SYM(AllGather); PM_SYM(Send); PM_SYM(Recv);
        PM_SYM(GroupStart); PM_SYM(GroupEnd); PM_SYM(GetErrorString); PM_SYM(GetVersion);
#undef PM_SYM
        if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.Send || !a.Recv || !a.GroupStart || !a.GroupEnd) a.lib = nullptr;
    });
    if (!a.lib) throw Unsupported("NCCL transport needs libnccl.so.2 at run time (not found / incomplete)");
    return a;
}
#define NK(call)                                                                                                   \
    do {                                                                                                           \
        ncclResult_t r_ = (call);                                                                                  \
        if (r_ != ncclSuccess)                                                                                     \
            throw CudaError(std::string(#call) + ": " + (ncclApi().GetErrorString ? ncclApi().GetErrorString(r_) : "NCCL error") + \
                            " (" __FILE__ ":" + std::to_string(__LINE__) + ")");                                   \
    } while (0)

}  // namespace

struct LocalGroup { std::vector<pm_comm*> members; };

struct pm_comm {
    int rank = 0, n = 1, device = 0;
    pm_workspace* ws = nullptr;
    ncclComm_t nccl = nullptr;                 // NCCL transport
    std::shared_ptr<LocalGroup> grp;           // in-process transport
    // exchange buffers (device): send side and the per-rank segments of the receive side
    DevBuf<uint4> xSend, xRecv;                // [n][1 + capPair] 16-byte slots
    DevBuf<uint2> gSend, gRecv;                // [kGHeaderSlots + capG] pairs, n of them on the receive side
    DevBuf<uint4> rSend, rRecv;                // [2 + 5 recX] slots
    DevBuf<u32> tSend, tRecv;                  // [kTWords]
    DevBuf<u32> tFull, tFullRecv;              // [5][capT] (only when a rank has more than kTieHead ties)
    DevBuf<u32> exportInfo;                    // [2] largest per-destination export count, local unique seeds
    DevBuf<u64> agree; PinBuf<u64> hAgree;     // sizing agreement before the first sample
    u32 capPair = 0, capG = 0, recX = 128;
    u64 localCap = 0, partCap = 0;             // table slots in use while seeding the slice / while holding the partition (one allocation)
    PinBuf<u32> hT;
    cudaEvent_t ready[5]{};
    u64 nLocalReads = 0, localBases = 0;
    u64 sent = 0, received = 0;
    // peer-memory transport (NCCL communicators whose ranks could map each other's receive buffers)
    bool peerOn = false, peerOff = false;      // peerOff: tried and not available (or switched off): stay with NCCL
    void* peerPtr[5][kPeerMax]{};              // [xRecv, gRecv, rRecv, tRecv, flags][rank]: that buffer of that rank as mapped in this process
    void* peerOpened[5][kPeerMax]{};           // what cudaIpcOpenMemHandle returned, to be closed again
    void* peerOpenedBase[5][kPeerMax]{};       // base of the opened allocation behind peerPtr (shared when two buffers sit in one allocation)
    void* mappedLocal[5]{};                    // the local buffers the mapping was made for (a reallocation invalidates it)
    DevBuf<u32> flags;                         // [4][kPeerMax]: flags[k][q] = epoch of the last exchange k whose payload from rank q is complete
    DevBuf<unsigned char> ipcStage; PinBuf<unsigned char> hIpc;
    u32 epoch = 0;
    const char* transport = "local";
    // per-call inputs
    const char* hReads = nullptr; const uint4* hPacked = nullptr; const uint64_t* hOff = nullptr; bool resident = false;
};

namespace {

size_t xSeg(const pm_comm* c) { return (size_t)c->capPair + 1; }
size_t gStride(const pm_comm* c) { return (size_t)kGHeaderSlots + c->capG; }
size_t rSlots(const pm_comm* c) { return 2 + (size_t)5 * c->recX; }

void ensureBuffers(pm_comm* c) {
    const size_t n = (size_t)c->n;
    c->xSend.ensure(n * xSeg(c)); c->xRecv.ensure(n * xSeg(c));
    c->gSend.ensure(gStride(c)); c->gRecv.ensure(n * gStride(c));
    c->rSend.ensure(rSlots(c)); c->rRecv.ensure(n * rSlots(c));
    c->tSend.ensure(kTWords); c->tRecv.ensure(n * kTWords);
    c->hT.ensure(n * kTWords);
    if (!c->exportInfo.p) c->exportInfo.alloc(2);
}

// k: 0 seed segments (all-to-all), 1 finalized pairs, 2 records, 3 tie heads, 4 full tie lists (all-gathers)
struct Xfer { const void* send; void* recv; size_t bytes; };   // per-rank payload; all-to-all: one segment
Xfer xferOf(pm_comm* c, int k, u32 capT) {
    switch (k) {
        case 0: return {c->xSend.p, c->xRecv.p, xSeg(c) * sizeof(uint4)};
        case 1: return {c->gSend.p, c->gRecv.p, gStride(c) * sizeof(uint2)};
        case 2: return {c->rSend.p, c->rRecv.p, rSlots(c) * sizeof(uint4)};
        case 3: return {c->tSend.p, c->tRecv.p, (size_t)kTWords * 4};
        default: return {c->tFull.p, c->tFullRecv.p, (size_t)5 * capT * 4};
    }
}

void exchange(std::vector<pm_comm*>& cs, int k, u32 capT = 0) {
    if (cs[0]->nccl && cs[0]->peerOn && k <= 3) {
        pm_comm* c = cs[0];
        const Xfer x = xferOf(c, k, capT);
        PushArgs A{};
        A.src = static_cast<const unsigned char*>(x.send);
        A.srcStride = k == 0 ? x.bytes : 0;
        A.dstOffset = (size_t)c->rank * x.bytes; A.segBytes = x.bytes;
        A.kind = (u32)k; A.capEntries = k == 0 ? c->capPair : c->capG; A.n = (u32)c->n; A.epoch = c->epoch;
        for (int q = 0; q < c->n; ++q) {
            A.dst[q] = static_cast<unsigned char*>(c->peerPtr[k][q]);
            A.flag[q] = static_cast<u32*>(c->peerPtr[4][q]) + (size_t)k * kPeerMax + c->rank;
        }
        launchPushSegments(A, c->ws->view, c->ws->st);
        launchWaitFlags(c->flags.p + (size_t)k * kPeerMax, (u32)c->n, c->epoch, c->ws->view, c->ws->st);
        c->sent += x.bytes * (size_t)(c->n - 1); c->received += x.bytes * (size_t)(c->n - 1);   // capacity; only the filled part travels
        return;
    }
    if (cs[0]->nccl) {
        pm_comm* c = cs[0];
        NcclApi& N = ncclApi();
        cudaStream_t st = c->ws->st;
        const Xfer x = xferOf(c, k, capT);
        if (k == 0) {
            NK(N.GroupStart());
            for (int q = 0; q < c->n; ++q) {
                if (q == c->rank) continue;
                NK(N.Send(static_cast<const char*>(x.send) + (size_t)q * x.bytes, x.bytes, ncclChar, q, c->nccl, st));
                NK(N.Recv(static_cast<char*>(x.recv) + (size_t)q * x.bytes, x.bytes, ncclChar, q, c->nccl, st));
            }
            NK(N.GroupEnd());
            CK(cudaMemcpyAsync(static_cast<char*>(x.recv) + (size_t)c->rank * x.bytes, static_cast<const char*>(x.send) + (size_t)c->rank * x.bytes, x.bytes,
                               cudaMemcpyDeviceToDevice, st));
            c->sent += x.bytes * (size_t)(c->n - 1); c->received += x.bytes * (size_t)(c->n - 1);
        } else {
            NK(N.AllGather(x.send, x.recv, x.bytes, ncclChar, c->nccl, st));
            c->sent += x.bytes * (size_t)(c->n - 1); c->received += x.bytes * (size_t)(c->n - 1);
        }
        return;
    }
    // in-process: everybody announces "my send buffer is complete", then everybody pulls from everybody
    for (pm_comm* c : cs) { setDevice(c->ws->idx->device); CK(cudaEventRecord(c->ready[k], c->ws->st)); }
    for (pm_comm* c : cs) {
        setDevice(c->ws->idx->device);
        const Xfer mine = xferOf(c, k, capT);
        for (pm_comm* q : cs) {
            const Xfer theirs = xferOf(q, k, capT);
            if (q != c) CK(cudaStreamWaitEvent(c->ws->st, q->ready[k], 0));
            const char* src = static_cast<const char*>(theirs.send) + (k == 0 ? (size_t)c->rank * theirs.bytes : 0);
            char* dst = static_cast<char*>(mine.recv) + (size_t)q->rank * mine.bytes;
            CK(cudaMemcpyPeerAsync(dst, c->ws->idx->device, src, q->ws->idx->device, mine.bytes, c->ws->st));
            if (q != c) { c->received += mine.bytes; q->sent += mine.bytes; }
        }
    }
}

// one u64 per rank, host-synchronised: used once per communicator (and after capacity trouble) to agree on the first capacities
u64 agreeMax(std::vector<pm_comm*>& cs, u64 (*value)(pm_comm*)) {
    u64 mx = 0;
    if (cs[0]->nccl) {
        pm_comm* c = cs[0];
        if (!c->agree.p) c->agree.alloc((size_t)c->n + 1);
        c->hAgree.ensure((size_t)c->n + 1);
        c->hAgree.p[0] = value(c);
        cudaStream_t st = c->ws->st;
        CK(cudaMemcpyAsync(c->agree.p, c->hAgree.p, 8, cudaMemcpyHostToDevice, st));
        NK(ncclApi().AllGather(c->agree.p, c->agree.p + 1, 8, ncclChar, c->nccl, st));
        CK(cudaMemcpyAsync(c->hAgree.p + 1, c->agree.p + 1, 8 * (size_t)c->n, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int q = 0; q < c->n; ++q) mx = std::max(mx, c->hAgree.p[1 + q]);
    } else {
        for (pm_comm* c : cs) mx = std::max(mx, value(c));
    }
    return mx;
}

// Maps every rank's receive buffers (and flag block) into this process.  Collective: all ranks call it at the same points (the
// capacities that trigger it are agreed values) and all of them end up with the same answer -- a rank that cannot map makes everybody
// stay with NCCL.
bool peerExchangeWanted() {
    static const bool v = [] { const char* e = std::getenv("PM_PEER_EXCHANGE"); return e ? std::atoi(e) != 0 : true; }();
    return v;
}
void unmapPeers(pm_comm* c) {
    for (int b = 0; b < 5; ++b)
        for (int q = 0; q < kPeerMax; ++q) {
            if (c->peerOpened[b][q]) { cudaIpcCloseMemHandle(c->peerOpened[b][q]); c->peerOpened[b][q] = nullptr; }
            c->peerPtr[b][q] = nullptr; c->peerOpenedBase[b][q] = nullptr;
        }
    c->peerOn = false;
}
void mapPeers(pm_comm* c) {
    if (!c->nccl || c->peerOff || c->n < 2) return;
    void* mine[5] = {c->xRecv.p, c->gRecv.p, c->rRecv.p, c->tRecv.p, c->flags.p};
    // the buffers only move when they grow, and they grow by the same rule from the same agreed capacities on every rank
    if (c->peerOn && std::memcmp(mine, c->mappedLocal, sizeof(mine)) == 0) return;
    if (!peerExchangeWanted()) { c->peerOff = true; return; }
    cudaStream_t st = c->ws->st;
    CK(cudaStreamSynchronize(st));   // nobody may still be writing into buffers that are about to be unmapped
    unmapPeers(c);
    if (!c->flags.p) { c->flags.alloc((size_t)4 * kPeerMax); CK(cudaMemsetAsync(c->flags.p, 0, (size_t)4 * kPeerMax * sizeof(u32), st)); }
    mine[4] = c->flags.p;
    constexpr size_t kHd = sizeof(cudaIpcMemHandle_t) + 8;        // handle of the allocation + offset of the buffer inside it
    constexpr size_t kRec = 5 * kHd + 8;                           // five of those + "I can map" word
    const size_t n = (size_t)c->n;
    c->ipcStage.ensure((n + 1) * kRec); c->hIpc.ensure((n + 1) * kRec);
    unsigned char* h = c->hIpc.p;
    std::memset(h, 0, kRec);
    bool ok = true;
    // cudaMalloc may carve small buffers out of a larger allocation; an IPC handle always names the whole allocation
    typedef int (*AddrRangeFn)(unsigned long long*, size_t*, unsigned long long);
    AddrRangeFn addrRange = nullptr;
    {
        void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) addrRange = reinterpret_cast<AddrRangeFn>(fn);
        else { cudaGetLastError(); ok = false; }
    }
    for (int b = 0; b < 5 && ok; ++b) {
        cudaIpcMemHandle_t hd;
        unsigned long long base = 0; size_t size = 0;
        if (addrRange(&base, &size, (unsigned long long)(uintptr_t)mine[b]) != 0) { ok = false; break; }
        if (cudaIpcGetMemHandle(&hd, reinterpret_cast<void*>((uintptr_t)base)) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        const u64 offset = (u64)((uintptr_t)mine[b] - (uintptr_t)base);
        std::memcpy(h + (size_t)b * kHd, &hd, sizeof(hd));
        std::memcpy(h + (size_t)b * kHd + sizeof(hd), &offset, 8);
    }
    auto gather = [&]() {   // record 0 of the host block -> records 1..n of everybody
        CK(cudaMemcpyAsync(c->ipcStage.p, h, kRec, cudaMemcpyHostToDevice, st));
        NK(ncclApi().AllGather(c->ipcStage.p, c->ipcStage.p + kRec, kRec, ncclChar, c->nccl, st));
        CK(cudaMemcpyAsync(h + kRec, c->ipcStage.p + kRec, n * kRec, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    };
    h[5 * kHd] = ok ? 1 : 0;
    gather();
    for (size_t q = 0; q < n; ++q) ok = ok && h[(1 + q) * kRec + 5 * kHd] == 1;
    if (ok) {
        for (size_t q = 0; q < n && ok; ++q)
            for (int b = 0; b < 5 && ok; ++b) {
                if ((int)q == c->rank) { c->peerPtr[b][q] = mine[b]; continue; }
                cudaIpcMemHandle_t hd; u64 offset = 0;
                std::memcpy(&hd, h + (1 + q) * kRec + (size_t)b * kHd, sizeof(hd));
                std::memcpy(&offset, h + (1 + q) * kRec + (size_t)b * kHd + sizeof(hd), 8);
                // two buffers of a rank may live in one allocation: an allocation can be opened only once per process
                void* ptr = nullptr;
                for (int pb = 0; pb < b && !ptr; ++pb)
                    if (std::memcmp(h + (1 + q) * kRec + (size_t)pb * kHd, &hd, sizeof(hd)) == 0) ptr = c->peerOpenedBase[pb][q];
                if (!ptr) {
                    if (cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
                    c->peerOpened[b][q] = ptr;
                }
                c->peerOpenedBase[b][q] = ptr;
                c->peerPtr[b][q] = static_cast<unsigned char*>(ptr) + offset;
            }
    }
    // second round: did everybody manage to open everything?
    h[5 * kHd] = ok ? 1 : 0;
    gather();
    for (size_t q = 0; q < n; ++q) ok = ok && h[(1 + q) * kRec + 5 * kHd] == 1;
    if (!ok) { unmapPeers(c); c->peerOff = true; c->transport = "nccl"; return; }
    c->peerOn = true; c->transport = "nccl bootstrap + peer-memory exchanges";
    std::memcpy(c->mappedLocal, mine, sizeof(mine));
}

void phase0(pm_comm* c, const pm_place_params& prm) {
    pm_workspace* W = c->ws;
    setDevice(W->idx->device);
    CK(cudaEventRecord(W->ev[0], W->st));
    if (c->localCap == 0) { c->localCap = 1 << 16; while (c->localCap < c->localBases / 4) c->localCap <<= 1; c->partCap = c->localCap; }
    ensureTable(W, std::max(c->localCap, c->partCap));   // the allocation covers both phases; each uses its own prefix
    W->tableCap = c->localCap;
    refreshView(W);
    if (W->stageTimers) CK(cudaEventRecord(W->ev[1], W->st));
    if (c->resident) stageSeed(W, true, prm);
    else if (c->hPacked) uploadAndSeedPipelinedPacked(W, c->hPacked, c->hOff, c->nLocalReads, prm);
    else uploadAndSeedPipelined(W, c->hReads, c->hOff, c->nLocalReads, prm);
    launchPartitionExport(W->view, (u32)c->n, c->capPair, c->xSend.p, c->exportInfo.p, W->st);
    if (W->stageTimers) CK(cudaEventRecord(W->ev[2], W->st));
}
void phase1(pm_comm* c, const pm_place_params& prm) {
    pm_workspace* W = c->ws; pm_index* I = W->idx;
    setDevice(I->device);
    W->tableCap = c->partCap;
    refreshView(W);
    launchTableClear(W->view, W->st);
    launchPartitionImport(W->view, c->xRecv.p, (u32)c->n, c->capPair, W->st);
    launchPartitionFinalize(I->view, W->view, makeOpts(prm, false), I->homo.p, I->nSM, c->gSend.p, c->capG, c->nLocalReads, c->exportInfo.p, 0, W->st);
}
void phase2(pm_comm* c, const pm_place_params& prm) {
    pm_workspace* W = c->ws; pm_index* I = W->idx;
    setDevice(I->device);
    static const bool kSideScalars = [] { const char* e = std::getenv("PM_SIDE_SCALARS"); return e ? std::atoi(e) != 0 : true; }();
    launchGatheredFinalize(I->view, W->view, makeOpts(prm, false), c->gRecv.p, (u32)c->n, c->capG, I->nSM, W->st, kSideScalars ? W->stCopy : nullptr, W->evFork, W->evJoin);
    W->joinPending = kSideScalars;
    stageDeltasScoresRecords(W, prm);
    launchRecordsPack(W->view, c->rSend.p, c->recX, W->st);
}
void phase3(pm_comm* c, const pm_place_params& prm) {
    pm_workspace* W = c->ws; pm_index* I = W->idx;
    setDevice(I->device);
    launchChainGathered(W->view, c->rRecv.p, (u32)c->n, c->recX, W->st);
    launchTies(I->view, W->view, makeOpts(prm, false), W->st);
    launchTiesPack(W->view, c->tSend.p, c->gSend.p, c->exportInfo.p, W->st);
    if (W->stageTimers) CK(cudaEventRecord(W->ev[6], W->st));
}
void phase4(pm_comm* c) {
    pm_workspace* W = c->ws; pm_index* I = W->idx;
    setDevice(I->device);
    // the reset (ell entries of the gathered pairs, boundary records) on the side stream while the copy engine delivers the two small blocks
    CK(cudaEventRecord(W->evFork, W->st));
    CK(cudaStreamWaitEvent(W->stCopy, W->evFork, 0));
    launchResetGathered(I->view, W->view, c->gRecv.p, (u32)c->n, c->capG, W->stCopy);
    CK(cudaEventRecord(W->evJoin, W->stCopy));
    enqueueSmall(W);
    CK(cudaMemcpyAsync(c->hT.p, c->tRecv.p, (size_t)c->n * kTWords * 4, cudaMemcpyDeviceToHost, W->st));
    CK(cudaStreamWaitEvent(W->st, W->evJoin, 0));
    CK(cudaEventRecord(W->ev[7], W->st));
}

const THeader* tHdr(const pm_comm* c, int q) { return reinterpret_cast<const THeader*>(c->hT.p + (size_t)q * kTWords); }

// table capacities for the next sample: 2-4x the entries each phase saw (the local table while seeding, the partition afterwards)
u64 fitCap(u64 cap, u64 entries) {
    if (entries * 10 > cap * 7 || (cap > (1u << 16) && cap > 4 * entries)) { cap = 1 << 16; while (cap < 2 * entries) cap <<= 1; }
    return cap;
}
void fitTable(pm_comm* c) {
    const THeader* h = tHdr(c, c->rank);
    c->localCap = fitCap(c->localCap, h->localEntries);
    c->partCap = fitCap(c->partCap, h->partEntries);
    c->ws->lastEntries = std::max<u64>(h->localEntries, h->partEntries);
}

void runSharded(std::vector<pm_comm*>& cs, const pm_place_params* prm, pm_place_result* res0) {
    checkParams(prm);
    for (pm_comm* c : cs) if (!workspaceAlive(c->ws)) throw std::runtime_error("communicator's workspace was destroyed");
    if (prm->dedup_reads) throw Unsupported("dedup_reads needs the whole sample on one GPU (duplicates across the ranks' read slices would go unseen): use pm_place");
    if (prm->seed_mask_fraction > 0.0) throw Unsupported("seed_mask_fraction needs the whole seed table on one GPU: use pm_place");
    if (prm->min_seed_quality > 0) throw Unsupported("min_seed_quality is not available for sharded samples: use pm_place_quality");
    for (pm_comm* c : cs) { c->ws->wantMetrics = false; c->ws->lastParams = *prm; c->ws->haveResult = false; c->sent = c->received = 0; }
    if (cs[0]->capPair == 0) {
        // first sample: nothing is known about the sample yet except its size.  Unique seeds per rank <= seed instances ~ 0.3 per base.
        const u64 maxBases = agreeMax(cs, [](pm_comm* c) { return c->localBases; });
        const u64 perRank = maxBases * 3 / 10 + 4096;
        for (pm_comm* c : cs) {
            c->capPair = (u32)std::min<u64>(perRank / (u64)c->n * 5 / 4 + 4096, 0x7FFFFFF0u);
            c->capG = (u32)std::min<u64>(perRank * 5 / 4 + 4096, 0x7FFFFFF0u);
        }
    }
    for (int attempt = 0; attempt < 6; ++attempt) {
        for (pm_comm* c : cs) { c->capG = (c->capG + 1u) & ~1u; setDevice(c->ws->idx->device); ensureBuffers(c); }   // even: per-rank segments stay 16-byte aligned
        for (pm_comm* c : cs) { mapPeers(c); ++c->epoch; }
        for (pm_comm* c : cs) phase0(c, *prm);
        exchange(cs, 0);
        for (pm_comm* c : cs) phase1(c, *prm);
        exchange(cs, 1);
        for (pm_comm* c : cs) phase2(c, *prm);
        exchange(cs, 2);
        for (pm_comm* c : cs) phase3(c, *prm);
        exchange(cs, 3);
        for (pm_comm* c : cs) phase4(c);
        for (pm_comm* c : cs) { setDevice(c->ws->idx->device); CK(cudaStreamSynchronize(c->ws->st)); parseSmall(c->ws); }
        // every rank holds the same gathered headers: the decisions below are identical everywhere
        pm_comm* c0 = cs[0];
        u32 flags = 0, maxPair = 0, maxG = 0, maxTie = 0;
        for (int q = 0; q < c0->n; ++q) {
            const THeader* h = tHdr(c0, q);
            flags |= h->flags; maxPair = std::max(maxPair, h->maxPairCount); maxG = std::max(maxG, h->gEntries);
            for (int m = 0; m < 5; ++m) maxTie = std::max(maxTie, h->tieCount[m]);
        }
        const u32 wantPair = maxPair + maxPair / 4 + 1024, wantG = maxG + maxG / 4 + 1024;
        for (pm_comm* c : cs) flags |= (u32)c->ws->hAcc.overflow & (u32)kOvfPeer;   // the last exchange itself may have timed out
        if (flags & (u32)kOvfPeer) throw std::runtime_error("sharded placement: a rank's data did not arrive (peer-memory exchange timed out)");
        if (flags) {
            if (attempt == 5) break;
            for (pm_comm* c : cs) {
                setDevice(c->ws->idx->device);
                if (flags & kOvfPair) c->capPair = std::max(c->capPair, wantPair);
                // a pair overflow starves the partitions: the gather counts of this attempt are too small to size from
                if (flags & kOvfGather) c->capG = std::max<u32>(c->capG * 2, wantG);
                if (flags & kOvfRecords) c->recX *= 4;
                if (flags & kOvfTable) { c->localCap *= 4; c->partCap *= 4; }
            }
            continue;
        }
        for (pm_comm* c : cs) { setDevice(c->ws->idx->device); c->capPair = wantPair; c->capG = wantG; fitTable(c); }
        // ties: the heads came with the result; longer lists take one more (rare) exchange
        std::vector<u32> lists; unsigned cnt[5] = {0, 0, 0, 0, 0};
        if (maxTie > (u32)kTieHead) {
            const u32 capT = maxTie;
            for (pm_comm* c : cs) {
                setDevice(c->ws->idx->device);
                c->tFull.ensure((size_t)5 * capT); c->tFullRecv.ensure((size_t)c->n * 5 * capT);
                launchTiesFullPack(c->ws->view, c->tFull.p, capT, c->ws->st);
            }
            exchange(cs, 4, capT);
            for (pm_comm* c : cs) {
                setDevice(c->ws->idx->device);
                c->hT.ensure((size_t)c->n * kTWords + (size_t)c->n * 5 * capT);
            }
            // (hT may have moved: the headers are re-read from the device copy that is still intact)
            for (pm_comm* c : cs) {
                setDevice(c->ws->idx->device);
                CK(cudaMemcpyAsync(c->hT.p, c->tRecv.p, (size_t)c->n * kTWords * 4, cudaMemcpyDeviceToHost, c->ws->st));
                CK(cudaMemcpyAsync(c->hT.p + (size_t)c->n * kTWords, c->tFullRecv.p, (size_t)c->n * 5 * capT * 4, cudaMemcpyDeviceToHost, c->ws->st));
            }
            for (pm_comm* c : cs) { setDevice(c->ws->idx->device); CK(cudaStreamSynchronize(c->ws->st)); }
            for (pm_comm* c : cs) {
                lists.clear();
                const u32* full = c->hT.p + (size_t)c->n * kTWords;
                for (int m = 0; m < 5; ++m) {
                    cnt[m] = 0;
                    for (int q = 0; q < c->n; ++q) {
                        const u32 k = std::min(tHdr(c, q)->tieCount[m], capT);
                        const u32* src = full + ((size_t)q * 5 + m) * capT;
                        lists.insert(lists.end(), src, src + k); cnt[m] += k;
                    }
                }
                finishTies(c->ws, lists.data(), cnt);
            }
        } else {
            for (pm_comm* c : cs) {
                lists.clear();
                for (int m = 0; m < 5; ++m) {
                    cnt[m] = 0;
                    for (int q = 0; q < c->n; ++q) {
                        const THeader* h = tHdr(c, q);
                        const u32* heads = c->hT.p + (size_t)q * kTWords + sizeof(THeader) / 4 + (size_t)m * kTieHead;
                        lists.insert(lists.end(), heads, heads + h->tieCount[m]); cnt[m] += h->tieCount[m];
                    }
                }
                finishTies(c->ws, lists.data(), cnt);
            }
        }
        for (pm_comm* c : cs) {
            pm_place_result tmp{};
            pm_place_result* r = (c == cs[0] && res0) ? res0 : &tmp;
            std::memset(r, 0, sizeof(*r));
            fillResult(c->ws, r, (u64)c->ws->hAcc.totalReads);
            setDevice(c->ws->idx->device);
            recordStageTimes(c->ws, r);
            c->ws->haveResult = true;
        }
        return;
    }
    throw std::runtime_error("sharded placement: internal capacities kept overflowing");
}

void checkGroup(pm_comm* const* comms, int n) {
    if (!comms || n < 1) throw std::runtime_error("null / empty communicator list");
    for (int r = 0; r < n; ++r) {
        if (!comms[r] || !comms[r]->grp || comms[r]->n != n || comms[r]->rank != r) throw std::runtime_error("pm_place_multi: pass the n communicators of pm_comm_create_local in rank order");
        if (comms[r]->grp != comms[0]->grp) throw std::runtime_error("pm_place_multi: communicators of different groups");
    }
}

}  // namespace

extern "C" {

int pm_comm_unique_id(void* id_out) {
    if (!id_out) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        static_assert(sizeof(ncclUniqueId) == PM_COMM_ID_BYTES, "ncclUniqueId size");
        ncclUniqueId id;
        NK(ncclApi().GetUniqueId(&id));
        std::memcpy(id_out, &id, sizeof(id));
        return PM_OK;
    });
}

int pm_comm_create_nccl(pm_workspace* ws, const void* id, int rank, int n_ranks, pm_comm** out) {
    if (!ws || !id || !out || n_ranks < 1 || n_ranks > 32 || rank < 0 || rank >= n_ranks) return fail(PM_ERR_INVALID, "bad argument (1 <= n_ranks <= 32)");
    *out = nullptr;
    return guarded([&]() -> int {
        setDevice(ws->idx->device);
        std::unique_ptr<pm_comm> c(new pm_comm());
        c->rank = rank; c->n = n_ranks; c->ws = ws; c->device = ws->device; c->transport = "nccl";
        ncclUniqueId uid; std::memcpy(&uid, id, sizeof(uid));
        NK(ncclApi().CommInitRank(&c->nccl, n_ranks, uid, rank));
        *out = c.release();
        return PM_OK;
    });
}

int pm_comm_create_local(pm_workspace* const* ws, int n_ranks, pm_comm** out) {
    if (!ws || !out || n_ranks < 1 || n_ranks > 32) return fail(PM_ERR_INVALID, "bad argument (1 <= n_ranks <= 32)");
    for (int r = 0; r < n_ranks; ++r) { out[r] = nullptr; if (!ws[r]) return fail(PM_ERR_INVALID, "null workspace"); }
    return guarded([&]() -> int {
        auto grp = std::make_shared<LocalGroup>();
        std::vector<std::unique_ptr<pm_comm>> cs;
        for (int r = 0; r < n_ranks; ++r) {
            setDevice(ws[r]->idx->device);
            std::unique_ptr<pm_comm> c(new pm_comm());
            c->rank = r; c->n = n_ranks; c->ws = ws[r]; c->grp = grp; c->device = ws[r]->device;
            for (auto& e : c->ready) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            cs.push_back(std::move(c));
        }
        // peer access where the ranks sit on different devices (cudaMemcpyPeerAsync stages through the host without it)
        for (int a = 0; a < n_ranks; ++a)
            for (int b = 0; b < n_ranks; ++b) {
                const int da = ws[a]->idx->device, db = ws[b]->idx->device;
                if (da == db) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, da, db) == cudaSuccess && can) { setDevice(da); if (cudaDeviceEnablePeerAccess(db, 0) != cudaSuccess) cudaGetLastError(); }
            }
        for (int r = 0; r < n_ranks; ++r) { grp->members.push_back(cs[r].get()); out[r] = cs[r].release(); }
        return PM_OK;
    });
}

void pm_comm_destroy(pm_comm* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (workspaceAlive(c->ws)) cudaStreamSynchronize(c->ws->st);   // a binding's garbage collector may have destroyed the workspace first
    else cudaDeviceSynchronize();
    unmapPeers(c);
    if (c->nccl) { try { ncclApi().CommDestroy(c->nccl); } catch (...) {} }
    for (auto& e : c->ready) if (e) cudaEventDestroy(e);
    if (c->grp) for (auto& m : c->grp->members) if (m == c) m = nullptr;
    delete c;
}

int pm_place_sharded(pm_comm* c, const char* reads, const uint64_t* read_offsets, uint64_t n_reads_local, const pm_place_params* params,
                     pm_place_result* result) {
    if (!c || !read_offsets || (!reads && n_reads_local) || !result) return fail(PM_ERR_INVALID, "null argument");
    if (!c->nccl) return fail(PM_ERR_INVALID, "pm_place_sharded needs an NCCL communicator (in-process groups: pm_place_multi)");
    return guarded([&]() -> int {
        c->hReads = reads; c->hPacked = nullptr; c->hOff = read_offsets; c->nLocalReads = n_reads_local; c->localBases = n_reads_local ? read_offsets[n_reads_local] : 0; c->resident = false;
        std::vector<pm_comm*> cs{c};
        runSharded(cs, params, result);
        return PM_OK;
    });
}
int pm_place_sharded_packed(pm_comm* c, const void* packed, const uint64_t* read_offsets, uint64_t n_reads_local, const pm_place_params* params,
                            pm_place_result* result) {
    if (!c || !read_offsets || (!packed && n_reads_local) || !result) return fail(PM_ERR_INVALID, "null argument");
    if (!c->nccl) return fail(PM_ERR_INVALID, "pm_place_sharded_packed needs an NCCL communicator");
    if ((reinterpret_cast<uintptr_t>(packed) & 15u) != 0) return fail(PM_ERR_INVALID, "packed reads must be 16-byte aligned");
    return guarded([&]() -> int {
        c->hReads = nullptr; c->hPacked = static_cast<const uint4*>(packed); c->hOff = read_offsets; c->nLocalReads = n_reads_local;
        c->localBases = n_reads_local ? read_offsets[n_reads_local] : 0; c->resident = false;
        std::vector<pm_comm*> cs{c};
        runSharded(cs, params, result);
        return PM_OK;
    });
}
int pm_place_sharded_resident(pm_comm* c, const pm_place_params* params, pm_place_result* result) {
    if (!c || !result) return fail(PM_ERR_INVALID, "null argument");
    if (!c->nccl) return fail(PM_ERR_INVALID, "pm_place_sharded_resident needs an NCCL communicator (in-process groups: pm_place_multi_resident)");
    if (!workspaceAlive(c->ws)) return fail(PM_ERR_INVALID, "communicator's workspace was destroyed");
    if (!c->ws->residentValid) return fail(PM_ERR_INVALID, "pm_place_sharded_resident: call pm_reads_upload first");
    return guarded([&]() -> int {
        c->nLocalReads = c->ws->nReads; c->localBases = c->ws->totalBases; c->resident = true;
        std::vector<pm_comm*> cs{c};
        runSharded(cs, params, result);
        return PM_OK;
    });
}

int pm_place_multi(pm_comm* const* comms, int n_ranks, const char* reads, const uint64_t* read_offsets, uint64_t n_reads,
                   const pm_place_params* params, pm_place_result* result) {
    if (!read_offsets || (!reads && n_reads) || !result) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        checkGroup(comms, n_ranks);
        if (n_reads && read_offsets[0] != 0) throw std::runtime_error("read_offsets[0] must be 0");
        std::vector<pm_comm*> cs(comms, comms + n_ranks);
        std::vector<std::vector<uint64_t>> offs((size_t)n_ranks);   // every slice starts at offset 0 of its own byte range
        for (int r = 0; r < n_ranks; ++r) {
            const uint64_t lo = n_reads * (uint64_t)r / (uint64_t)n_ranks, hi = n_reads * (uint64_t)(r + 1) / (uint64_t)n_ranks;
            offs[r].resize(hi - lo + 1);
            for (uint64_t i = lo; i <= hi; ++i) offs[r][i - lo] = read_offsets[i] - read_offsets[lo];
            cs[r]->hReads = reads + read_offsets[lo]; cs[r]->hPacked = nullptr; cs[r]->hOff = offs[r].data(); cs[r]->nLocalReads = hi - lo;
            cs[r]->localBases = read_offsets[hi] - read_offsets[lo]; cs[r]->resident = false;
        }
        runSharded(cs, params, result);
        return PM_OK;
    });
}
int pm_place_multi_resident(pm_comm* const* comms, int n_ranks, const pm_place_params* params, pm_place_result* result) {
    if (!result) return fail(PM_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        checkGroup(comms, n_ranks);
        std::vector<pm_comm*> cs(comms, comms + n_ranks);
        for (pm_comm* c : cs) {
            if (!workspaceAlive(c->ws)) throw std::runtime_error("communicator's workspace was destroyed");
            if (!c->ws->residentValid) throw std::runtime_error("pm_place_multi_resident: call pm_reads_upload on every rank's workspace first");
            c->nLocalReads = c->ws->nReads; c->localBases = c->ws->totalBases; c->resident = true;
        }
        runSharded(cs, params, result);
        return PM_OK;
    });
}

const char* pm_comm_transport(pm_comm* c) { return c ? c->transport : ""; }

int pm_comm_last_traffic(pm_comm* c, uint64_t* bytes_sent, uint64_t* bytes_received) {
    if (!c) return fail(PM_ERR_INVALID, "null argument");
    if (bytes_sent) *bytes_sent = c->sent;
    if (bytes_received) *bytes_received = c->received;
    return PM_OK;
}

}  // extern "C"
