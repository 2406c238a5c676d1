// pm_shard_kernels.cu -- kernels of the "one sample over several GPUs" path (the reference's own
// level-parallel traversal + merge is placement.cpp:742-913, its thread-local seed maps are merged by mergeSeedMaps, :922-929).
//
// Rank r seeds its slice of the reads into its local count table exactly like the one-GPU path.  Then, with no host round trip:
//   partition_export     local table -> one send segment per owner rank (owner = high bits of the mixed seed hash)
//   [all-to-all]
//   partition_import     received segments -> this rank's partition of the sample's seed table (counts of the same seed add up)
//   table_scan           (pm_kernels.cu) homopolymer removal, min-support statistics, compaction -- on 1/N of the seeds
//   partition_finalize   dictionary look-up of every partition entry; header (statistics) + (read count, seed id) pairs
//   [all-gather]
//   gathered_finalize    every rank: min-support rule from the summed statistics, log1p table, exact magnitude sums, histogram,
//                        ell[seed id] scatter -- the replicated state every shard's node_deltas gathers from
//   node_deltas / prefix_scores / bfs_* on the rank's own node range (pm_kernels.cu)
//   records_pack, [all-gather], chain_gathered      the tolerance chain replayed identically on every rank
//   collect_ties (local), ties_pack, [all-gather]   the union over ranks is the reference's tie list
// Every buffer has a fixed capacity and carries its fill count in band; a count above the capacity raises a flag that travels with
// the last all-gather, so that all ranks take the same grow-and-redo decision after the single device-to-host copy of the sample.
#include "pm_device.cuh"

namespace pm {

// ------------------------------------------------------------------------------------------------------
// partition_export: one pass over the local table.  A block takes 2048 slots at a time: owners of the occupied slots, per-owner
// counts through warp match + one shared-memory atomic per (warp, owner), ONE global atomic per (tile, owner) for the segment
// offsets, then scattered 16-byte stores into the owners' segments.
// ------------------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 32;
constexpr int kExpSlots = 2048;
__global__ void __launch_bounds__(256) partition_export(WorkspaceView W, u32 nRanks, u32 capPair, uint4* __restrict__ xSend, u32* __restrict__ maxPair) {
    __shared__ unsigned sCnt[kMaxRanks], sBase[kMaxRanks];
    const TableSlot* table = W.table; const u64 cap = W.tableCap;
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    const size_t segStride = (size_t)capPair + 1;
    if (tid < kMaxRanks) sCnt[tid] = 0;
    __syncthreads();
    const u64 nTiles = (cap + kExpSlots - 1) / kExpSlots;
    for (u64 tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
        const u64 base = tile * kExpSlots;
        uint4 v[8]; unsigned where[8];   // owner << 24 | offset inside the tile's share of that owner's segment
#pragma unroll
        for (int q = 0; q < 8; ++q) { const u64 i = base + q * 256 + tid; v[q] = i < cap ? ldSlot(table, i) : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0); }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const u64 k = slotKey(v[q]);
            const bool occ = k != kEmptyKey && v[q].z > 0;
            const unsigned owner = occ ? seedOwner(k, nRanks) : 0xFFu;
            const unsigned peers = __match_any_sync(0xffffffffu, owner);
            where[q] = 0xFFFFFFFFu;
            if (occ) {
                const int leader = __ffs(peers) - 1;
                unsigned o = 0;
                if ((int)lane == leader) o = atomicAdd(&sCnt[owner], (unsigned)__popc(peers));
                o = __shfl_sync(peers, o, leader) + __popc(peers & ((1u << lane) - 1u));
                where[q] = (owner << 24) | o;
            }
        }
        __syncthreads();
        if (tid < nRanks) {
            const unsigned c = sCnt[tid];
            sBase[tid] = c ? atomicAdd(reinterpret_cast<u32*>(xSend + (size_t)tid * segStride), c) : 0u;   // header.count of the segment
            sCnt[tid] = 0;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (where[q] == 0xFFFFFFFFu) continue;
            const unsigned owner = where[q] >> 24;
            const unsigned o = sBase[owner] + (where[q] & 0xFFFFFFu);
            if (o < capPair) xSend[(size_t)owner * segStride + 1 + o] = make_uint4(v[q].x, v[q].y, v[q].z, 0u);
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && tid == 0 && W.acc->emptyKeyCount > 0) {   // the one key that cannot live in the table travels to rank 0
        const unsigned o = atomicAdd(reinterpret_cast<u32*>(xSend), 1u);
        if (o < capPair) xSend[1 + o] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, (u32)min((long long)0xFFFFFFFFLL, W.acc->emptyKeyCount), 0u);
    }
    // after the export (the block that finishes last): largest per-destination count (sizing feedback) and the overflow flag; also
    // clears emptyKeyCount, whose share now sits in rank 0's segment
    __shared__ unsigned sLast;
    if (!lastBlockDone(&W.acc->finDone, &sLast) || tid != 0) return;
    u32 mx = 0, sum = 0;
    for (u32 p = 0; p < nRanks; ++p) {
        XHeader* h = reinterpret_cast<XHeader*>(xSend + (size_t)p * segStride);
        const u32 c = __ldcg(&h->count);
        mx = max(mx, c); sum += c;
    }
    maxPair[0] = mx; maxPair[1] = sum;   // sizing feedback: largest per-destination count, unique seeds of the local table
    if (mx > capPair) raiseFlag(W.acc, kOvfPair);
    __threadfence();
    const u32 fl = (u32)atomicOr(reinterpret_cast<unsigned long long*>(&W.acc->overflow), 0ULL);
    for (u32 p = 0; p < nRanks; ++p) reinterpret_cast<XHeader*>(xSend + (size_t)p * segStride)->flags = fl;
    W.acc->emptyKeyCount = 0;
}
void launchPartitionExport(WorkspaceView W, u32 nRanks, u32 capPair, uint4* xSend, u32* maxPairCount, cudaStream_t st) {
    // the nRanks segment headers (16 bytes at the start of every segment) with one strided memset
    cudaMemset2DAsync(xSend, ((size_t)capPair + 1) * sizeof(uint4), 0, sizeof(uint4), nRanks, st);
    const u64 nTiles = (W.tableCap + kExpSlots - 1) / kExpSlots;
    const unsigned grid = (unsigned)std::min<u64>(nTiles ? nTiles : 1, 148ull * 4);
    noteLaunch(), partition_export<<<grid, 256, 0, st>>>(W, nRanks, capPair, xSend, maxPairCount);
}

// received segments -> the partition table
__global__ void __launch_bounds__(256) partition_import(WorkspaceView W, const uint4* __restrict__ xRecv, u32 nRanks, u32 capPair) {
    __shared__ u32 sCount[kMaxRanks];
    const size_t segStride = (size_t)capPair + 1;
    if (threadIdx.x < nRanks) {
        const uint4 h = xRecv[(size_t)threadIdx.x * segStride];
        u32 c = h.x;
        if (c > capPair) { raiseFlag(W.acc, kOvfPair); c = capPair; }
        if (h.y) raiseFlag(W.acc, h.y);   // the sender's own trouble
        sCount[threadIdx.x] = c;
    }
    __syncthreads();
    const u64 total = (u64)nRanks * capPair;
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 seg = (u32)(idx / capPair), i = (u32)(idx % capPair);
        if (i >= sCount[seg]) continue;
        const uint4 e = __ldcs(xRecv + (size_t)seg * segStride + 1 + i);
        tableInsert(W.table, W.tableMask, slotKey(e), e.z, W.acc);
    }
}
void launchPartitionImport(WorkspaceView W, const uint4* xRecv, u32 nRanks, u32 capPair, cudaStream_t st) {
    noteLaunch(), partition_import<<<streamGrid((u64)nRanks * capPair, 2), 256, 0, st>>>(W, xRecv, nRanks, capPair);
}

// ------------------------------------------------------------------------------------------------------
// partition_finalize: runs after table_scan on the partition table (compacted entries in entKey / entCnt, statistics in scanPart).
// Which entries every rank needs: those the min-support rule can keep.  The automatic rule resolves to 1 or 2 (placement.cpp:
// 931-955) and is only known once all partitions' statistics are summed, so seeds with count 1 travel when the index holds them
// (their ell entry matters) and otherwise only as a count (they add log1p(1) terms to the magnitudes, nothing else).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) partition_finalize(DevIndexView I, WorkspaceView W, int configuredMinSupport, unsigned nScanParts,
                                                          uint2* __restrict__ gSend, u32 capG, u64 nLocalReads, const u32* __restrict__ maxPair,
                                                          u32 localEntries) {
    __shared__ long long sStat[4];
    __shared__ unsigned sN, sBase;
    __shared__ unsigned long long sOne;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    GHeader* hdr = reinterpret_cast<GHeader*>(gSend);
    uint2* ent = gSend + kGHeaderSlots;
    SampleAcc* acc = W.acc;
    if (tid == 0) { sN = 0; sOne = 0; }
    if (blockIdx.x == 0 && warp < 4) {
        long long t = 0;
        for (unsigned b = lane; b < nScanParts; b += 32) t += reinterpret_cast<const long long*>(&W.scanPart[b])[warp];
        t = warpSumLL(t);
        if (lane == 0) sStat[warp] = t;
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0) {
        hdr->multiSum = sStat[0]; hdr->multiCount = sStat[1]; hdr->unique = sStat[2]; hdr->total = sStat[3];
        hdr->nReads = nLocalReads; hdr->maxPairCount = *maxPair; hdr->localEntries = localEntries;
    }
    // possibly kept: count >= 1 and, with a configured minimum, >= it.  Of those only (count 1, not in the index) travel as a count.
    const u32 cfgMin = configuredMinSupport > 0 ? (u32)configuredMinSupport : 1u;
    const unsigned n = acc->entCount;
    const unsigned nIter = (n + gridDim.x * 256u - 1) / (gridDim.x * 256u);
    unsigned long long ones = 0;
    for (unsigned it = 0; it < nIter; ++it) {
        const unsigned i = (it * gridDim.x + blockIdx.x) * 256u + tid;
        u32 c = 0, id = kNone; bool emit = false;
        if (i < n) {
            c = __ldcs(&W.entCnt[i]);
            if (c >= cfgMin) {
                id = dictLookup(I, __ldcs(&W.entKey[i]));
                if (c == 1 && id == kNone) ++ones; else emit = true;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, emit);
        unsigned o = 0;
        if (m) {
            if (lane == 0) o = atomicAdd(&sN, (unsigned)__popc(m));
            o = __shfl_sync(0xffffffffu, o, 0) + __popc(m & ((1u << lane) - 1u));
        }
        __syncthreads();
        if (tid == 0) { sBase = sN ? atomicAdd(&hdr->nEntries, sN) : 0u; sN = 0; }
        __syncthreads();
        if (emit && sBase + o < capG) ent[sBase + o] = make_uint2(c, id);
        __syncthreads();
    }
    if (blockIdx.x == 0 && tid == 0 && acc->emptyKeyCount > 0) {   // rank 0 only: the key outside the table; never in the index
        const u32 c = (u32)min((long long)0xFFFFFFFFLL, acc->emptyKeyCount);
        if (c == 1 && cfgMin <= 1) ++ones;
        else if (c >= cfgMin) { const unsigned o = atomicAdd(&hdr->nEntries, 1u); if (o < capG) ent[o] = make_uint2(c, kNone); }
    }
    ones = (unsigned long long)warpSumLL((long long)ones);
    if (lane == 0 && ones) atomicAdd(&sOne, ones);
    __syncthreads();
    if (tid == 0 && sOne) atomicAdd(reinterpret_cast<unsigned long long*>(&hdr->n1NotIndex), sOne);
    // the flags go last (they must include what this very kernel raised): the block that finishes last
    __shared__ unsigned sLast;
    if (!lastBlockDone(&acc->finDone, &sLast) || tid != 0) return;
    if (__ldcg(&hdr->nEntries) > capG) raiseFlag(acc, kOvfGather);
    __threadfence();
    hdr->flags = (u32)atomicOr(reinterpret_cast<unsigned long long*>(&acc->overflow), 0ULL);
}
void launchPartitionFinalize(DevIndexView I, WorkspaceView W, PlaceOpts O, const u64* homo, int nSM, uint2* gSend, u32 capG, u64 nLocalReads,
                             const u32* maxPairCount, u32 localEntriesHint, cudaStream_t st) {
    unsigned nParts = 0;
    launchTableScan(W, homo, nSM, &nParts, st);
    cudaMemsetAsync(gSend, 0, sizeof(GHeader), st);
    noteLaunch(), partition_finalize<<<(unsigned)nSM * 2, 256, 0, st>>>(I, W, O.minReadSupport, nParts, gSend, capG, nLocalReads, maxPairCount, localEntriesHint);
}

// ------------------------------------------------------------------------------------------------------
// gathered_finalize: computeReadSeedMagnitudes (placement.cpp:957-984) over the lists of all ranks.  Integer sums, so the result is
// the same on every rank and for every rank count; block 0 also publishes the sample-wide statistics.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gathered_finalize(DevIndexView I, WorkspaceView W, int configuredMinSupport, const uint2* __restrict__ gRecv,
                                                         u32 nRanks, u32 capG) {
    __shared__ unsigned sHist[kHistSmem];
    __shared__ FinalizeShared sFin;
    __shared__ u32 sCount[kMaxRanks];
    __shared__ long long sStat[6];
    __shared__ u32 sFlags;
    const unsigned tid = threadIdx.x;
    const size_t stride = (size_t)kGHeaderSlots + capG;
    for (int i = tid; i < kHistSmem; i += blockDim.x) sHist[i] = 0;
    if (tid == 0) {
        long long ms = 0, mc = 0, un = 0, tot = 0, ones = 0, reads = 0; u32 fl = 0;
        for (u32 r = 0; r < nRanks; ++r) {
            const GHeader* h = reinterpret_cast<const GHeader*>(gRecv + (size_t)r * stride);
            ms += h->multiSum; mc += h->multiCount; un += h->unique; tot += h->total; ones += (long long)h->n1NotIndex; reads += (long long)h->nReads;
            fl |= h->flags;
            u32 c = h->nEntries;
            if (c > capG) { fl |= (u32)kOvfGather; c = capG; }
            sCount[r] = c;
        }
        sStat[0] = ms; sStat[1] = mc; sStat[2] = un; sStat[3] = tot; sStat[4] = ones; sStat[5] = reads; sFlags = fl;
    }
    __syncthreads();
    const u32 minSup = (u32)resolveMinSupport(sStat[0], sStat[1], configuredMinSupport);
    SampleAcc* acc = W.acc;
    if (blockIdx.x == 0 && tid == 0) {
        acc->multiSum = sStat[0]; acc->multiCount = sStat[1]; acc->entries = sStat[2]; acc->unique = sStat[2]; acc->total = sStat[3];
        acc->totalReads = sStat[5];
        const long long n1 = minSup <= 1 ? sStat[4] : 0;
        acc->n1NotIndex = n1;
        if (n1 > 0) atomicAdd(&W.countHist[1], (unsigned)min(n1, 0xFFFFFFFFLL));
        if (sFlags) raiseFlag(acc, sFlags);
    }
    FinalizeAcc A; A.mag = fxZero(); A.lsum = fxZero(); A.kept = 0; A.maxc = 0;
    const u64 total = (u64)nRanks * capG;
    for (u64 idx = (u64)blockIdx.x * blockDim.x + tid; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 seg = (u32)(idx / capG), i = (u32)(idx % capG);
        if (i >= sCount[seg]) continue;
        const uint2 e = __ldg(gRecv + (size_t)seg * stride + kGHeaderSlots + i);
        const u32 c = e.x;
        if (c >= minSup && c != 0) {
            const double l = finalizeSums(I, W, c, A, sHist);
            if (e.y != kNone) W.ell[e.y] = __double2ll_rn(l * kEllScale);
        }
    }
    finalizeBlockEpilogue(W, A, sHist, &sFin);
}
void launchGatheredFinalize(DevIndexView I, WorkspaceView W, PlaceOpts O, const uint2* gRecv, u32 nRanks, u32 capG, int nSM, cudaStream_t st,
                            cudaStream_t stSide, cudaEvent_t evFork, cudaEvent_t evJoin) {
    u64 g = ((u64)nRanks * capG + 255) / 256; if (g < 1) g = 1; if (g > (u64)nSM * 4) g = (u64)nSM * 4; if (g > kMaxPartials) g = kMaxPartials;
    noteLaunch(), gathered_finalize<<<(unsigned)g, 256, 0, st>>>(I, W, O.minReadSupport, gRecv, nRanks, capG);
    if (stSide && evFork && evJoin) {   // as in launchFinalize: the scalars beside node_deltas, joined before prefix_scores
        cudaEventRecord(evFork, st);
        cudaStreamWaitEvent(stSide, evFork, 0);
        launchRootAndScalars(I, W, O, (unsigned)g, stSide);
        cudaEventRecord(evJoin, stSide);
    } else launchRootAndScalars(I, W, O, (unsigned)g, st);
}
// after the sample: clear exactly the ell entries it set, and the segment records that are combined with atomics
__global__ void __launch_bounds__(256) reset_gathered(DevIndexView I, WorkspaceView W, const uint2* __restrict__ gRecv, u32 nRanks, u32 capG) {
    const size_t stride = (size_t)kGHeaderSlots + capG;
    const u64 total = (u64)nRanks * capG;
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 seg = (u32)(idx / capG), i = (u32)(idx % capG);
        const u32 cnt = min(reinterpret_cast<const GHeader*>(gRecv + (size_t)seg * stride)->nEntries, capG);
        if (i >= cnt) continue;
        const u32 id = gRecv[(size_t)seg * stride + kGHeaderSlots + i].y;
        if (id != kNone) W.ell[id] = 0;
    }
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < I.nBoundary; i += gridDim.x * blockDim.x)
        *reinterpret_cast<uint4*>(W.segRec + I.boundarySegs[i]) = make_uint4(0u, 0u, 0u, 0u);
}
void launchResetGathered(DevIndexView I, WorkspaceView W, const uint2* gRecv, u32 nRanks, u32 capG, cudaStream_t st) {
    noteLaunch(), reset_gathered<<<148 * 4, 256, 0, st>>>(I, W, gRecv, nRanks, capG);
}

// ------------------------------------------------------------------------------------------------------
// selection across ranks
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) records_pack(WorkspaceView W, uint4* __restrict__ rSend, u32 recX) {
    RHeader* hdr = reinterpret_cast<RHeader*>(rSend);
    RRecord* rec = reinterpret_cast<RRecord*>(rSend + 2);
    const int m = blockIdx.x;
    const unsigned n = W.acc->recordCount[m];
    if (threadIdx.x == 0) {
        hdr->count[m] = n;
        if (n > recX || n > W.recCap) raiseFlag(W.acc, kOvfRecords);
    }
    for (unsigned i = threadIdx.x; i < n && i < recX && i < W.recCap; i += blockDim.x) {
        RRecord r; r.score = W.recScore[(size_t)m * W.recCap + i]; r.rank = W.recRank[(size_t)m * W.recCap + i]; r.node = W.recNode[(size_t)m * W.recCap + i];
        rec[(size_t)m * recX + i] = r;
    }
    __shared__ unsigned sLast;   // the flags go last: they include what the five blocks raised
    if (!lastBlockDone(&W.acc->finDone, &sLast) || threadIdx.x != 0) return;
    hdr->flags = (u32)atomicOr(reinterpret_cast<unsigned long long*>(&W.acc->overflow), 0ULL);
}
void launchRecordsPack(WorkspaceView W, uint4* rSend, u32 recX, cudaStream_t st) {
    noteLaunch(), records_pack<<<5, 128, 0, st>>>(W, rSend, recX);
}

// the tolerance chain (placement.cpp:355-371) over the records of all ranks; block m = metric m.  Records of a rank are in no
// particular order: every step picks the lowest-rank record after the last event that beats best + tol.
__global__ void __launch_bounds__(256) chain_gathered(WorkspaceView W, const uint4* __restrict__ rRecv, u32 nRanks, u32 recX) {
    __shared__ ChainShared S;
    __shared__ u32 sCount[kMaxRanks];
    const int m = blockIdx.x;
    const size_t rSlots = 2 + (size_t)5 * recX;
    if (threadIdx.x < nRanks) {
        const RHeader* h = reinterpret_cast<const RHeader*>(rRecv + (size_t)threadIdx.x * rSlots);
        sCount[threadIdx.x] = min(h->count[m], recX);
        u32 fl = h->flags;
        for (int q = 0; q < 5; ++q) if (h->count[q] > recX) fl |= (u32)kOvfRecords;
        if (m == 0 && fl) raiseFlag(W.acc, fl);
    }
    __syncthreads();
    const Selection s = chainReplay(S, nRanks * recX, [&](unsigned idx, u32& r, double& x, u32& v) {
        if ((idx % recX) >= sCount[idx / recX]) return false;
        const RRecord* rec = reinterpret_cast<const RRecord*>(rRecv + (size_t)(idx / recX) * rSlots + 2) + (size_t)m * recX + (idx % recX);
        r = rec->rank; x = rec->score; v = rec->node;
        return true;
    });
    if (threadIdx.x == 0) W.sel[m] = s;
}
void launchChainGathered(WorkspaceView W, const uint4* rRecv, u32 nRanks, u32 recX, cudaStream_t st) { noteLaunch(), chain_gathered<<<5, 256, 0, st>>>(W, rRecv, nRanks, recX); }

__global__ void __launch_bounds__(64) ties_pack(WorkspaceView W, u32* __restrict__ tSend, const uint2* __restrict__ gSend, const u32* __restrict__ exportInfo) {
    THeader* hdr = reinterpret_cast<THeader*>(tSend);
    u32* heads = tSend + sizeof(THeader) / 4;
    if (threadIdx.x < 5) hdr->tieCount[threadIdx.x] = W.acc->tieCount[threadIdx.x];
    if (threadIdx.x == 5) {
        hdr->flags = (u32)W.acc->overflow;
        hdr->maxPairCount = exportInfo[0]; hdr->localEntries = exportInfo[1];
        hdr->gEntries = reinterpret_cast<const GHeader*>(gSend)->nEntries;
        hdr->partEntries = W.acc->entCount;
    }
    for (int m = 0; m < 5; ++m) {
        const unsigned n = min(W.acc->tieCount[m], (unsigned)kTieHead);
        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) heads[m * kTieHead + i] = W.tieHead[m * kTieHead + i];
    }
}
void launchTiesPack(WorkspaceView W, u32* tSend, const uint2* gSend, const u32* exportInfo, cudaStream_t st) { noteLaunch(), ties_pack<<<1, 64, 0, st>>>(W, tSend, gSend, exportInfo); }
// full local tie lists into a [5][capT] block (slow path: some rank has more than kTieHead ties)
__global__ void __launch_bounds__(256) ties_full_pack(WorkspaceView W, u32* __restrict__ out, u32 capT) {
    const int m = blockIdx.y;
    const unsigned n = min(min(W.acc->tieCount[m], W.tieCap), capT);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[(size_t)m * capT + i] = W.tieNode[(size_t)m * W.tieCap + i];
}
void launchTiesFullPack(WorkspaceView W, u32* out, u32 capT, cudaStream_t st) { noteLaunch(), ties_full_pack<<<dim3(64, 5), 256, 0, st>>>(W, out, capT); }


// ------------------------------------------------------------------------------------------------------
// seeding::hashSeq (seeding.cpp:20-30) for a batch of k-mers: f = XOR_i rol(c(b_i), k-1-i), r = XOR_i rol(c(comp b_{k-1-i}), k-1-i),
// rotations taken modulo 64 like the reference's rol.  One thread per sequence; status 1 = "Kmer contains non canonical base".
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hash_seq(const char* __restrict__ seqs, const u64* __restrict__ off, u64 n, u64* __restrict__ fwd, u64* __restrict__ rev,
                                                unsigned char* __restrict__ status) {
    for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (u64)gridDim.x * blockDim.x) {
        const u64 b = off[q], k = off[q + 1] - b;
        u64 f = 0, r = 0; bool bad = false;
        for (u64 i = 0; i < k; ++i) {
            const unsigned ci = baseCode((unsigned char)seqs[b + i]), cj = baseCode((unsigned char)seqs[b + k - 1 - i]);
            bad = bad || ci >= 4;
            f ^= rol64(codeHash(ci), (unsigned)((k - i - 1) & 63));
            r ^= rol64(cj < 4 ? codeHash(3 - cj) : 0ULL, (unsigned)((k - i - 1) & 63));
        }
        fwd[q] = f; rev[q] = r; status[q] = bad ? 1 : 0;
    }
}
// ------------------------------------------------------------------------------------------------------
// peer-memory exchanges (pm_multi.cu): the collectives of this path move a few megabytes at most and are bound by launch and
// protocol latency, so every rank simply WRITES its payload into the peers' buffers -- 16-byte stores over NVLink -- and raises one
// flag word per peer when the last block is through.  Only the filled part of a segment travels (its header says how much).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) push_segments(PushArgs A, WorkspaceView W) {
    __shared__ unsigned sLast;
    const u32 perPeer = gridDim.x / A.n;                 // blocks per destination
    const u32 q = blockIdx.x / perPeer, part = blockIdx.x % perPeer;
    if (q < A.n) {
        const unsigned char* src = A.src + (size_t)q * A.srcStride;
        size_t bytes = A.segBytes;
        if (A.kind == 0) bytes = ((size_t)min(reinterpret_cast<const XHeader*>(src)->count, A.capEntries) + 1) * sizeof(uint4);
        else if (A.kind == 1) bytes = (((size_t)min(reinterpret_cast<const GHeader*>(src)->nEntries, A.capEntries) + kGHeaderSlots) * sizeof(uint2) + 15) & ~(size_t)15;
        if (bytes > A.segBytes) bytes = A.segBytes;
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(A.dst[q] + A.dstOffset);
        const size_t n16 = bytes / 16;
        for (size_t i = (size_t)part * 256 + threadIdx.x; i < n16; i += (size_t)perPeer * 256) d4[i] = s4[i];
    }
    __threadfence_system();                              // this thread's remote stores are out before the block reports in
    if (!lastBlockDone(&W.acc->finDone, &sLast)) return;
    __threadfence_system();
    if (threadIdx.x < A.n) *reinterpret_cast<volatile u32*>(A.flag[threadIdx.x]) = A.epoch;
}
void launchPushSegments(const PushArgs& A, WorkspaceView W, cudaStream_t st) {
    const size_t per = (A.segBytes / 16 + 256 * 8 - 1) / (256 * 8);           // ~8 stores per thread
    const unsigned perPeer = (unsigned)std::min<size_t>(std::max<size_t>(per, 1), 64);
    noteLaunch(), push_segments<<<perPeer * A.n, 256, 0, st>>>(A, W);
}
// bounded wait (about half a minute of polling): a rank that never delivers must not hang the GPU; the sample then fails with kOvfPeer
__global__ void wait_flags(const u32* __restrict__ flags, u32 n, u32 epoch, WorkspaceView W) {
    if (threadIdx.x < n) {
        const volatile u32* f = flags + threadIdx.x;
        const long long t0 = clock64();
        while (*f != epoch) {
            __nanosleep(200);
            if (clock64() - t0 > 60000000000LL) { raiseFlag(W.acc, kOvfPeer); break; }
        }
    }
    __threadfence_system();
}
void launchWaitFlags(const u32* flags, u32 n, u32 epoch, WorkspaceView W, cudaStream_t st) { noteLaunch(), wait_flags<<<1, 32, 0, st>>>(flags, n, epoch, W); }

void launchHashSeq(const char* seqs, const u64* off, u64 n, u64* fwd, u64* rev, unsigned char* status, cudaStream_t st) {
    if (n) noteLaunch(), hash_seq<<<streamGrid(n, 1), 256, 0, st>>>(seqs, off, n, fwd, rev, status);
}

}  // namespace pm
