#!/usr/bin/env python3
"""bench.py -- placement hot path throughput on B200 (metric of BASELINE.json: placement nodes x reads scored / s).

One step = one pass of the whole place stage over one synthetic sample (BASELINE config 3 shape: 1M-node tree, 30 kb
genome, 1M x 150 bp reads; generator = tools/synth):  pack -> seed -> count table -> filters/magnitudes ->
delta kernel -> exact tree prefix + scores -> tolerance-chain selection -> 5 best nodes + tie lists on the host.

  value   : inputs (reads, offsets) already resident in HBM; CUDA-event time of K steps on the library's stream
  e2e     : the same call through the C ABI with HOST (pinned) buffers, H2D of the reads and D2H of the result inside
  roofline: dominant kernel of the step (by measured time) + the scoring kernel north_star names, algorithmic bytes/launch
  cpu_baseline / --impl reference: the reference's own placeLite (oracle/_ref, compiled unmodified) on the host cores

N > 1 (torchrun): ONE sample over the N GPUs (BASELINE configs[2]: node range sharded, reads sliced for seeding, seed table
hash-partitioned; pm_comm over NCCL) -- strong scaling, this is `value`; the sample-sharded batch mode (one replica and one sample per GPU,
no collective) is timed as well and reported under "batch_mode".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2]
    "c3": dict(name="synthetic 1M-node tree, 30 kb genome, 1M x 150 bp reads (BASELINE configs[2])", n_nodes=1_000_000, genome=30_000,
               lam=1.0, n_reads=1_000_000, read_len=150),
    "c3-small": dict(name="synthetic 100k-node tree, 30 kb genome, 100k x 150 bp reads (reduced configs[2], dev only)", n_nodes=100_000,
                     genome=30_000, lam=1.0, n_reads=100_000, read_len=150),
    # BASELINE.json configs[0]: the reference's own CPU-runnable case, real data (index built by the reference's IndexBuilder, oracle/_ref/data)
    "c1": dict(name="sars_20000_twilight_dipper.panman + isolate_R1/R2.fastq.gz, single sample (BASELINE configs[0])", real=True),
    # BASELINE.json configs[3]
    "c4": dict(name="synthetic bacterial-scale tree: 100k nodes, 5 Mb genome, 10M x 150 bp reads (BASELINE configs[3])", n_nodes=100_000,
               genome=5_000_000, lam=50.0, n_reads=10_000_000, read_len=150),
    "c4-small": dict(name="synthetic 20k nodes, 1 Mb genome, 1M x 150 bp reads (reduced configs[3], dev only)", n_nodes=20_000,
                     genome=1_000_000, lam=50.0, n_reads=1_000_000, read_len=150),
    # BASELINE.json configs[4]: sample-sharded batch; the index has config 1's shape (40k nodes, 30 kb genome, ~60 seed deltas per node)
    "c5": dict(name="batch placement of 1,024 synthetic samples (100k x 133 bp reads each, random leaves) against a sars_20000-shaped index (BASELINE configs[4])",
               n_nodes=40_000, genome=30_000, lam=3.3, n_reads=100_000, read_len=133, batch=1024),
    "c5-small": dict(name="batch of 64 synthetic samples against a sars_20000-shaped index (reduced configs[4], dev only)",
                     n_nodes=40_000, genome=30_000, lam=3.3, n_reads=100_000, read_len=133, batch=64),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """SM clock + clock-event (throttle) reasons DURING the timed region (B200_PROFILING.md's clocks line).  The timed regions here are tens of
    milliseconds, so the samples come from NVML inside a thread (one query is ~0.1 ms; a sample every millisecond) rather than from an
    `nvidia-smi -lms 100` child, which cannot deliver a single line in that time; nvidia-smi stays as the fallback when NVML is unusable."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4),
               ("hw_power_brake_slowdown", 0x80))

    def __init__(self, gpu=0):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [x for x in vis.split(",") if x.strip().isdigit()]
        self.gpu = int(ids[gpu]) if gpu < len(ids) else gpu
        self.sm, self.mask, self.mx = [], 0, None
        self.stop_flag = threading.Event()
        self.th = None
        self.source = None

    def _nvml_loop(self, nv, h):
        while True:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                pass
            if self.stop_flag.wait(0.001):
                return

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.source = "nvml"
            self.th = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.th.start()
        except Exception:
            self.source = None
            self.th = None

    def _smi_once(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
            f = [x.strip() for x in out.strip().splitlines()[0].split(",")]
            self.sm.append(float(f[0])); self.mx = float(f[1])
            for (n, bit), v in zip(self.REASONS[:4], f[2:6]):
                if v.lower().startswith("active"):
                    self.mask |= bit
            self.source = "nvidia-smi (one query right after the timed region, GPU still clocked up)"
        except Exception:
            pass

    def stop(self):
        if self.th is not None:
            self.stop_flag.set()
            self.th.join(timeout=2)
        if not self.sm:
            self._smi_once()
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}
        reasons = sorted(n for n, bit in self.REASONS if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.mx, "reasons": reasons, "samples": len(self.sm), "source": self.source}


class RealSample:
    """config 1: the reference-built sars_20000 index (.idx read by the product's own reader) + the isolate reads (the host shim's parser)"""

    def __init__(self):
        import panmap_b200 as pm
        from tests import helpers as H
        for p in (H.SARS_IDX, H.ISOLATE_R1, H.ISOLATE_R2):
            if not os.path.exists(p):
                raise SystemExit(f"workload c1 needs {p} (staged by __graft_entry__.build() from the reference's bundled data)")
        self.host = pm.HostIndex.read(H.SARS_IDX)
        h = self.host
        self.hash, self.parent, self.child, self.offsets, self.parent_index = h.hash, h.parent, h.child, h.offsets, h.parent_index
        self.k, self.s, self.t, self.l, self.open = h.k, h.s, h.t, h.l, h.open
        self.n_nodes, self.n_deltas = h.n_nodes, h.n_deltas
        self.reads, self.read_offsets = pm.read_fastx(H.ISOLATE_R1, H.ISOLATE_R2)
        self.truth = -1
        self.idx_path, self.fastq = H.SARS_IDX, (H.ISOLATE_R1, H.ISOLATE_R2)


def make_workload(name, seed=0, n_reads=None):
    from tools.synth import synth
    w = dict(WORKLOADS[name])
    t = time.time()
    if w.get("real"):
        S = RealSample()
        w["n_reads"] = int(S.read_offsets.size - 1)
    else:
        S = shared_synth(name, w, seed, n_reads or w["n_reads"])
    S.gen_seconds = time.time() - t
    return S, w


_SYNTH_ARRAYS = ("hash", "parent", "child", "offsets", "parent_index", "reads", "read_offsets")
_SYNTH_SCALARS = ("truth", "k", "s", "t", "l", "open", "n_nodes", "n_deltas")


def shared_synth(name, w, seed, n_reads):
    """the synthetic workload; under torchrun the ranks of a node share ONE generation (local rank 0 generates and leaves the arrays in
    /dev/shm, the others wait for them): a bacterial-scale sample takes minutes to generate and N copies of the generator would share the
    same host cores.  Generation is deterministic, so this only saves time."""
    from tools.synth import synth
    world, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    gen = lambda: synth.generate(w["n_nodes"], w["genome"], w["lam"], n_reads, read_len=w["read_len"], seed=seed)
    if world <= 1 or not os.path.isdir("/dev/shm"):
        return gen()
    d = f"/dev/shm/pm_bench_{name}_{seed}_{n_reads}_{os.environ.get('MASTER_PORT', '0')}"
    done = os.path.join(d, "done")
    if local == 0:
        S = gen()
        os.makedirs(d, exist_ok=True)
        for a in _SYNTH_ARRAYS:
            np.save(os.path.join(d, a + ".npy"), getattr(S, a))
        json.dump({k: int(getattr(S, k)) for k in _SYNTH_SCALARS}, open(os.path.join(d, "scalars.json"), "w"))
        open(done, "w").write("1")
        import atexit
        import shutil
        atexit.register(shutil.rmtree, d, True)       # the other ranks have loaded their copies long before rank 0 exits (barriers in between)
        return S
    t0 = time.time()
    while not os.path.exists(done):
        if time.time() - t0 > 1800:
            raise SystemExit(f"waited 30 min for local rank 0 to generate {name}")
        time.sleep(0.2)
    S = synth.Synth()
    for a in _SYNTH_ARRAYS:
        setattr(S, a, np.load(os.path.join(d, a + ".npy")))
    for k, v in json.load(open(os.path.join(d, "scalars.json"))).items():
        setattr(S, k, v)
    S.truth_genome = b""
    return S


def algorithmic_bytes(S):
    """each input read once, each output written once, reference-native widths."""
    N, D = S.n_nodes, S.n_deltas
    bases = int(S.read_offsets[-1])
    score = 12 * D + 8 * (N + 1) + 4 * N + 40 * N   # delta SoA + offsets + parent index + five f64 scores per node
    return dict(seeding=bases, scoring=score, delta_kernel=12 * D + 8 * (N + 1), total=bases + score)


def write_fastq(S, n_reads, path):
    """FASTQ of the first n_reads reads of the synthetic sample (four lines per record, constant qualities)"""
    if os.path.exists(path):
        return path
    off = S.read_offsets
    buf = S.reads.tobytes()
    with open(path, "wb") as f:
        chunk = []
        for i in range(n_reads):
            s = buf[int(off[i]):int(off[i + 1])]
            chunk.append(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)))
            if len(chunk) >= 20000:
                f.write(b"".join(chunk)); chunk = []
        f.write(b"".join(chunk))
    return path


def reference_step(S, n_reads, threads, tmpdir, cache={}):
    """the reference's own placeLite (oracle/_ref: unmodified sources) on the synthetic index (written as a real uncompressed .idx) and a
    FASTQ of the first n_reads reads.  Returns (seconds of the whole call, result incl. the reference's own stage timers)."""
    from oracle import ref
    if "idx" not in cache:
        if hasattr(S, "idx_path"):
            cache["idx"] = ref.RefIndex(S.idx_path)
        else:
            p = os.path.join(tmpdir, "synth.idx")
            ref.write_index(p, S)
            cache["idx"] = ref.RefIndex(p)
    if hasattr(S, "fastq"):     # real sample: the files themselves (gzip inflate is part of the reference's read processing)
        fq, fq2 = S.fastq
    else:
        fq, fq2 = write_fastq(S, n_reads, os.path.join(tmpdir, f"reads_{n_reads}.fastq")), ""
    t = time.perf_counter()
    r = cache["idx"].place(fq, fq2, out_tsv=os.path.join(tmpdir, "ref.tsv"), threads=threads, stage_timers=True)
    return time.perf_counter() - t, r


def reference_spans(dt, r):
    """the reference's place stage split by its own timers (placement.cpp:1128,1691,1701,1929): FASTQ parse = read processing minus seed
    extraction; buffers -> result = the whole call minus that parse (what the GPU arm's `e2e` spans: reads in host memory -> placement)"""
    st = r["stage_ms"]
    parse = max(st["read_processing"] - max(st["seeding"], 0.0) - max(st["dedup"], 0.0), 0.0)
    return {"file_to_result_ms": 1e3 * dt, "fastq_parse_ms": parse, "seeding_ms": st["seeding"], "tree_traversal_ms": st["traversal"],
            "buffers_to_result_ms": 1e3 * dt - parse}


def same_placement(a_idx, a_tied, a_score, b_idx, b_tied, b_score, rtol=1e-12):
    """best node, tie list and best score of every metric agree"""
    ok = True
    for m in range(5):
        ok = ok and int(a_idx[m]) == int(b_idx[m]) and np.array_equal(np.asarray(a_tied[m], np.uint32), np.asarray(b_tied[m], np.uint32))
        ok = ok and abs(float(a_score[m]) - float(b_score[m])) <= rtol * max(abs(float(b_score[m])), 1e-9)
    return bool(ok)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libpanmap_ref.so not built (needs /root/reference at build time)"}))
        return
    S, w = make_workload(args.workload)
    threads = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as td:
        if w.get("real") or w.get("batch"):
            # one real sample / one sample of the batch per step (the whole sample: ~0.5 s per step)
            n_s = w["n_reads"]
            reference_step(S, n_s, threads, td)
        else:
            # bounded sample: full index, reads subsampled so that K+W steps end within a few minutes
            t_small, _ = reference_step(S, min(20000, w["n_reads"]), threads, td)
            t_mid, _ = reference_step(S, min(60000, w["n_reads"]), threads, td)
            per_read = max((t_mid - t_small) / 40000.0, 1e-7)
            fixed = max(t_small - 20000 * per_read, 0.05)
            budget = 150.0 / max(args.steps + args.warmup, 1)
            n_s = int(min(w["n_reads"], max(20000, (budget - fixed) / per_read)))
        for _ in range(args.warmup):
            reference_step(S, n_s, threads, td)
        spans = []
        for _ in range(args.steps):
            dt, r = reference_step(S, n_s, threads, td)
            spans.append(reference_spans(dt, r))
    mean = {k: float(np.mean([x[k] for x in spans])) for k in spans[0]}
    # headline of this arm = the same span as the GPU arm's e2e: reads in host memory -> placement on the host (the parse is reported beside it)
    ms = mean["buffers_to_result_ms"]
    value = S.n_nodes * n_s / (ms * 1e-3)
    line = {"impl": "reference", "metric": "placement nodes x reads scored per second", "value": value, "unit": "node*reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
            "vs_baseline": None, "dtype": "u64+f64", "data": "synthetic",
            "config": {"workload": w["name"], "n_nodes": S.n_nodes, "n_deltas": S.n_deltas, "n_reads": n_s, "k": int(S.k), "s": int(S.s), "l": int(S.l)},
            "spans": mean,
            "span_note": "value / ms_per_step / e2e = buffers_to_result (the whole placeLite call minus its FASTQ parse, by the reference's own stage timers): "
                         "the span the GPU arm's e2e covers.  file_to_result_ms is the whole call (parse of the FASTQ file + TSV write included).",
            "placed": {m: int(r["best_index"][i]) for i, m in enumerate(("log_raw", "log_cosine", "containment", "weighted_containment", "log_containment"))},
            "tied_counts": [int(len(t)) for t in r["tied"]], "best_scores": [float(x) for x in r["best_score"]],
            "cpu_baseline": {"value": value, "unit": "node*reads/s", "cores": threads, "kind": "reference",
                             "sample": f"reference placeLite (oracle/_ref, unmodified sources, std-container/oneTBB stand-ins) on the full {S.n_nodes}-node index "
                                       f"with the first {n_s} of {w['n_reads']} reads per step, {threads} host threads"},
            "e2e": {"value": value, "unit": "node*reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def ncu_traffic():
    """DRAM bytes (read + write) per launch of the hot kernels, from the committed ncu --set full capture of the CURRENT round
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py --json); None when there is no such file"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return {}, None
    d = json.load(open(p))
    return d.get("kernels", {}), d.get("capture")


def pinned_copy(L, arr):
    n = int(arr.nbytes)
    p = L.pm_host_alloc(n + 64)
    C.memmove(p, arr.ctypes.data, n)
    return p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=os.environ.get("PM_BENCH_WORKLOAD", "c3"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-file-span", action="store_true", help="skip the FASTQ file -> result measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import panmap_b200 as pm
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if pm.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the placement path has no CPU fallback")
    S, w = make_workload(args.workload)
    if WORKLOADS[args.workload].get("batch"):
        return run_batch(args, pm, world, rank, local)
    host = getattr(S, "host", None) or pm.HostIndex(S.hash, S.parent, S.child, S.offsets, S.parent_index, S.k, S.s, S.t, S.l)
    params = pm.PlaceParams()
    alg = algorithmic_bytes(S)
    pk, pk_src = peaks()
    nodes_reads = S.n_nodes * w["n_reads"]

    if world > 1:
        return run_multi(args, pm, S, w, host, params, alg, pk, pk_src, world, rank, local)

    index = pm.Index(host, device=0)
    ws = pm.Workspace(index)
    # ---- value: inputs resident in HBM ----
    ws.upload(S.reads, S.read_offsets)
    for _ in range(max(args.warmup, 3)):
        ws.place_resident(params, full=False)
    sampler = ClockSampler(0)
    sampler.start()
    # the timed region: the library's stage timers are off (its default, like the reference's: ten stream markers cost ~25 us per placement); only the
    # two CUDA events around the whole placement are recorded (stage_ms[7])
    ws.stage_timers(False)
    dev_sum = 0.0
    launches0 = pm.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dev_sum += ws.place_resident(params, full=False).stage_ms[7]
    wall = time.perf_counter() - t0
    launches = pm.launch_count() - launches0
    clocks = sampler.stop()
    dev_ms = dev_sum / args.steps
    # the breakdown: the same loop again with the stage timers on (per-stage and per-kernel CUDA events); its own total is reported beside it
    ws.stage_timers(True)
    stage = np.zeros(8)
    kern = np.zeros(3)
    ws.place_resident(params, full=False)
    for _ in range(args.steps):
        r = ws.place_resident(params, full=False)
        stage += np.array(list(r.stage_ms))
        kern += np.array(ws.last_kernel_ms())
    ws.stage_timers(False)
    stage /= args.steps
    kern /= args.steps
    value = nodes_reads / (dev_ms * 1e-3)
    res = ws.place_resident(params)  # full result for the record

    # ---- e2e: host (pinned) buffers through the C ABI ----
    L = pm.lib()
    nbytes = int(S.read_offsets[-1])
    hp_reads = pinned_copy(L, S.reads)
    hp_off = pinned_copy(L, S.read_offsets)
    for _ in range(3):
        ws.place_raw(hp_reads, hp_off, w["n_reads"], params)
    t0 = time.perf_counter()
    e2e_dev = 0.0
    for _ in range(args.steps):
        r2 = ws.place_raw(hp_reads, hp_off, w["n_reads"], params)
        e2e_dev += r2.stage_ms[7]
    e2e_wall = time.perf_counter() - t0
    e2e_value = nodes_reads * args.steps / e2e_wall
    h2d = nbytes + 8 * (w["n_reads"] + 1)            # reads + offsets (the chunk offsets are rebuilt on the device)
    d2h = 432 + 4 * int(sum(res.raw.tied_count))     # accumulators/scalars/selection block + tie lists
    # the same with the reads as 4-bit codes in pinned memory (what the packing FASTQ parser of the host shim emits): half the PCIe bytes
    t0 = time.perf_counter()
    packed = pm.host_pack_reads(S.reads, S.read_offsets)
    pack_ms = 1e3 * (time.perf_counter() - t0)
    hp_packed = pinned_copy(L, packed)
    for _ in range(3):
        ws.place_packed_raw(hp_packed, hp_off, w["n_reads"], params)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r3 = ws.place_packed_raw(hp_packed, hp_off, w["n_reads"], params)
    e2ep_wall = time.perf_counter() - t0
    rp = ws.place_packed(packed, S.read_offsets, params)
    e2e_packed = {"value": nodes_reads * args.steps / e2ep_wall, "unit": "node*reads/s", "ms_per_step": 1e3 * e2ep_wall / args.steps,
                  "h2d_bytes_per_step": int(packed.nbytes) + 8 * (w["n_reads"] + 1), "d2h_bytes_per_step": d2h,
                  "host_pack_ms_not_in_span": pack_ms,
                  "same_result_as_ascii": bool(all(rp.best_index[m] == res.best_index[m] and rp.best_score[m] == res.best_score[m] and np.array_equal(rp.tied[m], res.tied[m]) for m in pm.METRICS)),
                  "span": "buffers_to_result with the reads as 4-bit codes (pm_place_packed): 16-byte chunks of 32 bases in pinned host memory -> result"}
    L.pm_host_free(hp_reads); L.pm_host_free(hp_off); L.pm_host_free(hp_packed)
    spans = {"buffers_to_result_ms": 1e3 * e2e_wall / args.steps, "packed_buffers_to_result_ms": 1e3 * e2ep_wall / args.steps}

    # between-iteration cache state (timing rule): large samples exceed L2 by themselves; small ones get an explicit flush measurement beside the hot one
    ws_bytes = nbytes + 8 * nbytes // 3 + 12 * S.n_deltas // 3 + 80 * S.n_nodes
    l2_note = (f"per-step working set ~{ws_bytes / 1e6:.0f} MB (reads + syncmer lists + count table + delta words + scores) exceeds the 126 MB L2; no explicit flush"
               if ws_bytes > 200e6 else
               f"per-step working set ~{ws_bytes / 1e6:.0f} MB fits the 126 MB L2: `value` is the hot-cache number a batch of samples sees; see cold_cache for the L2-flushed one")
    cold = None
    if ws_bytes <= 200e6:
        try:
            import torch
            junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
            ws.upload(S.reads, S.read_offsets)
            cms = 0.0
            for _ in range(args.steps):
                junk.fill_(1); torch.cuda.synchronize()
                cms += ws.place_resident(params, full=False).stage_ms[7]
            cold = {"ms_per_step": cms / args.steps, "flush": "256 MB written to HBM between steps"}
            del junk
        except Exception as e:
            cold = {"error": str(e)}
    # ---- roofline ----
    # dominant kernel of the step = the syncmer kernel (CUDA events of the library around that launch alone); it is bound by the
    # integer ALU pipe, not by HBM, so its fraction of the HBM roofline is small by construction.  The HBM-bound kernel north_star
    # names (node_deltas) is reported next to it the same way.
    peak = float(pk["hbm_gbs"])
    traffic, capture = ncu_traffic()
    # kernel names as the library picks them (pm_kernels.cu launchSyncmers / launchCount): s = 8 and >= 20 k reads -> syncmers_rank, else
    # syncmers_fast; whole samples count with one lane per read
    syn_name = (f"syncmers_rank<{S.k}>" if int(S.s) == 8 and w["n_reads"] >= 20000 else f"syncmers_fast<{S.k},{S.s}>")
    cnt_name = f"count_seeds_lane<{S.k},{S.l}>" if w["n_reads"] >= 200000 else f"seeds_from_syncmers / count_seeds<{S.k},{S.l}>"
    names = ["h2d", f"seeding+table insert ({syn_name.split('<')[0]}, {cnt_name.split('<')[0]})", "table finalize", "node_deltas", "prefix_scores", "selection", "d2h"]
    per_kernel = {"pack_reads": float(kern[0]), syn_name: float(kern[1]), cnt_name: float(kern[2]),
                  "node_deltas": float(stage[3]), "prefix_scores": float(stage[4])}
    if per_kernel["pack_reads"] < 0.01:   # the default parameter sets hash straight from the ASCII reads: no pack_reads launch
        del per_kernel["pack_reads"]
    dom_name = max(per_kernel, key=per_kernel.get)
    dom_bytes = {"pack_reads": alg["seeding"] * 3 // 2, syn_name: alg["seeding"], cnt_name: 12 * int(res.raw.unique_seeds),
                 "node_deltas": alg["delta_kernel"], "prefix_scores": 80 * S.n_nodes}[dom_name]
    ach = dom_bytes / (per_kernel[dom_name] * 1e-3) / 1e9
    sc_ach = alg["delta_kernel"] / (stage[3] * 1e-3) / 1e9
    line = {
        "metric": "placement nodes x reads scored per second", "value": value, "unit": "node*reads/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64+f64", "data": "synthetic",
        "config": {"workload": w["name"], "n_nodes": S.n_nodes, "n_deltas": S.n_deltas, "n_reads": w["n_reads"], "read_bases": nbytes, "k": S.k, "s": S.s,
                   "l": S.l, "l2": l2_note,
                   "truth_node": int(S.truth), "placed": {m: int(res.best_index[m]) for m in pm.METRICS},
                   "unique_seeds": int(res.raw.unique_seeds), "kept_seeds": int(res.raw.read_unique_seed_count),
                   "min_read_support": int(res.raw.min_read_support), "index_distinct_seeds": int(index.num_distinct_seeds)},
        "wall_ms_per_step": 1e3 * wall / args.steps,
        "stage_ms": {n: float(stage[i]) for i, n in enumerate(names)},
        "stage_ms_note": f"stage_ms / kernel_ms / roofline come from a second pass of {args.steps} steps with the library's stage timers on ({float(stage[7]):.4f} ms per step there); "
                         "value / ms_per_step are measured with them off (the library default)",
        "e2e": {"value": e2e_value, "unit": "node*reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_wall / args.steps,
                "device_ms_per_step": e2e_dev / args.steps, "span": "buffers_to_result: ASCII reads + offsets in pinned host memory -> best nodes + tie lists on the host"},
        "e2e_packed": e2e_packed,
        "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps,
        "kernel_ms": per_kernel,
        "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": traffic.get(dom_name.split("<")[0]), "traffic_capture": capture, "algorithmic_bytes_per_launch": dom_bytes, "ms": per_kernel[dom_name],
                     "peak_source": pk_src,
                     "note": "dominant kernel of the step by measured time; it is limited by the integer ALU and L1 data pipes (ncu r02e: pipe_alu 68 %, l1 data "
                             "pipe 77 %, issue 54 %), one base in per ~58 thread instructions, so the HBM fraction is small by construction -- see roofline_scoring for the "
                             "HBM-bound kernel north_star names"},
        "roofline_scoring": {"bound": "hbm", "kernel": "node_deltas (the scoring kernel north_star names)", "achieved": sc_ach, "peak": peak, "unit": "GB/s",
                             "frac": sc_ach / peak, "traffic": traffic.get("node_deltas"), "algorithmic_bytes_per_launch": alg["delta_kernel"], "ms": float(stage[3]),
                             "note": "algorithmic bytes = 12 B/delta + 8 B/node of the reference layout (SURVEY 8d); the kernel itself streams 4 B/delta"},
        "place_stage": {"algorithmic_bytes": alg["total"], "achieved": alg["total"] / (dev_ms * 1e-3) / 1e9, "frac": alg["total"] / (dev_ms * 1e-3) / 1e9 / peak},
        "clocks": clocks,
    }
    if cold is not None:
        line["cold_cache"] = cold
    with tempfile.TemporaryDirectory() as td:
        if not args.no_file_span:
            # the other span: FASTQ FILE -> result through the C++ shim (multi-threaded flat parser + the same pm_place), same file the reference arm parses
            try:
                fq, fq2 = S.fastq if hasattr(S, "fastq") else (write_fastq(S, w["n_reads"], os.path.join(td, "reads_all.fastq")), "")
                pm.place_files(ws, fq, fq2, os.path.join(td, "gpu.tsv"), params)
                t0 = time.perf_counter()
                k3 = max(2, min(args.steps, 5))
                for _ in range(k3):
                    rf = pm.place_files(ws, fq, fq2, os.path.join(td, "gpu.tsv"), params)
                spans["file_to_result_ms"] = 1e3 * (time.perf_counter() - t0) / k3
                spans["file_bytes"] = os.path.getsize(fq) + (os.path.getsize(fq2) if fq2 else 0)
                spans["file_result_same_as_buffers"] = bool(all(int(rf.best_index[i]) == int(res.best_index[m]) for i, m in enumerate(pm.METRICS)))
            except Exception as e:  # reported, never fatal
                spans["file_to_result_error"] = str(e)
        line["spans"] = spans
        if not args.no_file_span:
            # opening the index (not part of a placement): .idx -> parse + flatten + upload, against the cached flattened image
            try:
                ip = getattr(S, "idx_path", None)
                if ip is None:
                    ip = os.path.join(td, "bench.idx")
                    host.write(ip)
                else:
                    import shutil
                    ip2 = os.path.join(td, "bench.idx"); shutil.copyfile(ip, ip2); ip = ip2
                t0 = time.perf_counter(); i1 = pm.Index.open_cached(ip); t1 = time.perf_counter(); i2 = pm.Index.open_cached(ip); t2 = time.perf_counter()
                line["index_open"] = {"idx_bytes": os.path.getsize(ip), "image_bytes": os.path.getsize(ip + ".pmflat"),
                                      "parse_flatten_write_image_upload_ms": 1e3 * (t1 - t0), "cached_image_upload_ms": 1e3 * (t2 - t1),
                                      "first_was_miss": i1.cache_hit is False, "second_was_hit": i2.cache_hit is True}
                del i1, i2
            except Exception as e:  # reported, never fatal
                line["index_open"] = {"error": str(e)}
        if w.get("real") and not args.no_file_span:
            # building the index (not part of a placement): .panman -> seed deltas, genomes materialised / seeded / sorted / diffed on the
            # GPU, against the reference's own IndexBuilder (oracle/_ref) on the box's host cores; both with --flank-mask 0 (the mode the product builds)
            try:
                from tests import helpers as TH
                if os.path.exists(TH.SARS_PANMAN):
                    pm.HostIndex.build_from_panman(TH.MAMMOTH_PANMAN, k=15, s=8, t=0, l=1)
                    t0 = time.perf_counter(); bi = pm.HostIndex.build_from_panman(TH.SARS_PANMAN, k=int(S.k), s=int(S.s), t=int(S.t), l=int(S.l)); tb = time.perf_counter() - t0
                    ib = {"panman": os.path.basename(TH.SARS_PANMAN), "flank_mask": 0, "n_nodes": bi.n_nodes, "n_deltas": bi.n_deltas, "gpu_ms": 1e3 * tb}
                    from oracle import ref as _ref
                    if _ref.available():
                        for th in (1, os.cpu_count() or 1):
                            t0 = time.perf_counter(); _ref.build_index(TH.SARS_PANMAN, os.path.join(td, "ref_f0.idx"), flank_mask=0, threads=th)
                            ib[f"reference_{th}_threads_ms"] = 1e3 * (time.perf_counter() - t0)
                        rb = pm.HostIndex.read(os.path.join(td, "ref_f0.idx"))
                        ib["nodes_identical_to_reference"] = int(sum(
                            np.array_equal(bi.hash[int(bi.offsets[v]):int(bi.offsets[v + 1])], rb.hash[int(rb.offsets[v]):int(rb.offsets[v + 1])]) for v in range(bi.n_nodes)))
                    line["index_build"] = ib
            except Exception as e:  # reported, never fatal
                line["index_build"] = {"error": str(e)}
        if not args.no_cpu_baseline:
            try:
                from oracle import ref
                if ref.available():
                    threads = os.cpu_count() or 1
                    n_s = min(w["n_reads"], 200_000)
                    reference_step(S, n_s, threads, td)          # first call maps the index (seedChangesLoaded): not part of the place stage
                    dt, rr = reference_step(S, n_s, threads, td)
                    sp = reference_spans(dt, rr)
                    # the GPU on the very same sample: agreement is checked, not assumed
                    sub_off = np.ascontiguousarray(S.read_offsets[:n_s + 1])
                    g = ws.place(S.reads[:int(sub_off[-1])], sub_off, params)
                    agree = same_placement([g.best_index[m] for m in pm.METRICS], [g.tied[m] for m in pm.METRICS], [g.best_score[m] for m in pm.METRICS],
                                           rr["best_index"], rr["tied"], rr["best_score"])
                    line["cpu_baseline"] = {"value": S.n_nodes * n_s / (sp["buffers_to_result_ms"] * 1e-3), "unit": "node*reads/s", "cores": threads, "kind": "reference",
                                            "seconds": dt, "spans": sp,
                                            "sample": f"first {n_s} of {w['n_reads']} reads on the full {S.n_nodes}-node index, ONE call: reference placeLite (oracle/_ref: unmodified "
                                                      f"reference sources, std-container + std::thread oneTBB stand-ins), {threads} host threads; value = its buffers_to_result span "
                                                      f"(whole call minus its FASTQ parse); the --impl reference arm runs more reads per step",
                                            "gpu_same_sample": {"best_nodes_tie_lists_scores_agree": agree, "gpu_ms": float(g.stage_ms[7])}}
                else:
                    line["cpu_baseline"] = {"value": None, "unit": "node*reads/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
            except Exception as e:  # the baseline is reported, never fatal
                line["cpu_baseline"] = {"value": None, "unit": "node*reads/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
    print(json.dumps(line))


def batch_samples(w, rank, world, limit=None):
    """this rank's share of the batch: groups of 64 samples, every group from its own random leaf of the same synthetic tree (same seed ->
    same tree and index in every call; the leaf is chosen by truth_frac).  Returns (index arrays of the first call, list of (reads, offsets, truth))."""
    from tools.synth import synth
    B = w["batch"]
    per_rank = B // world if limit is None else min(limit, B // world)
    group = 64
    out, first = [], None
    for g in range((per_rank + group - 1) // group):
        k = min(group, per_rank - g * group)
        frac = ((rank * 131 + g * 17 + 7) % 97) / 97.0 * 0.9 + 0.05
        S = synth.generate(w["n_nodes"], w["genome"], w["lam"], k * w["n_reads"], read_len=w["read_len"], seed=0, truth_frac=frac)
        if first is None:
            first = S
        n = w["n_reads"]
        for j in range(k):
            lo, hi = int(S.read_offsets[j * n]), int(S.read_offsets[(j + 1) * n])
            out.append((S.reads[lo:hi], np.ascontiguousarray(S.read_offsets[j * n:(j + 1) * n + 1] - np.uint64(lo)), int(S.truth)))
    return first, out


def run_batch(args, pm, world, rank, local):
    """BASELINE configs[4]: many samples against one index, sample-sharded over the GPUs (reference runBatchPlacement, main.cpp:1464-1666: TBB
    workers call placeLite concurrently on one LiteTree).  Every rank holds the index and its share of the samples; `threads` host threads per
    GPU each drive their own workspace (own stream) and pull samples from a shared queue.  No collective on the data path -> weak scaling.
    value: samples resident in a device pool (device-to-device hand-over per sample inside the timed region); e2e: samples in pinned host memory."""
    import threading as th
    import torch
    w = dict(WORKLOADS[args.workload])
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl")
    dev = torch.device("cuda", local)
    S0, samples = batch_samples(w, rank, world)
    host = pm.HostIndex(S0.hash, S0.parent, S0.child, S0.offsets, S0.parent_index, S0.k, S0.s, S0.t, S0.l)
    index = pm.Index(host, device=local)
    n_threads = int(os.environ.get("PM_BATCH_THREADS", "4"))
    wss = [pm.Workspace(index) for _ in range(n_threads)]
    params = pm.PlaceParams()
    L = pm.lib()
    # device pool: every sample's bytes and offsets in HBM
    d_reads = [torch.from_numpy(r).to(dev) for r, _, _ in samples]
    d_offs = [torch.from_numpy(o.view(np.int64)).to(dev) for _, o, _ in samples]
    torch.cuda.synchronize()
    n_s = len(samples)
    placed = np.full(n_s, -1, np.int64)

    def pass_over(fn):
        nxt = [0]; lock = th.Lock()
        def work(t):
            while True:
                with lock:
                    i = nxt[0]; nxt[0] += 1
                if i >= n_s:
                    return
                placed[i] = fn(wss[t], i)
        ts = [th.Thread(target=work, args=(t,)) for t in range(n_threads)]
        for x in ts: x.start()
        for x in ts: x.join()

    def resident(ws, i):
        ws.upload_device(d_reads[i].data_ptr(), d_offs[i].data_ptr(), samples[i][1])
        return int(ws.place_resident(params, full=False).best_index[0])
    hp = [(pinned_copy(L, r), pinned_copy(L, o)) for r, o, _ in samples[:min(n_s, 128)]]
    def from_host(ws, i):
        j = i % len(hp)
        return int(ws.place_raw(hp[j][0], hp[j][1], samples[j][1].size - 1, params).best_index[0])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    def maxf(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())
    def sumf(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.SUM); return float(t.item())

    steps, warm = max(1, args.steps // 4), max(1, args.warmup // 3)      # a step = one pass over the rank's whole share of the batch
    for _ in range(warm):
        pass_over(resident)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = pm.launch_count(); t0 = time.perf_counter()
    for _ in range(steps):
        pass_over(resident)
    torch.cuda.synchronize()
    dt = maxf(time.perf_counter() - t0); launches = sumf(pm.launch_count() - l0)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    on_truth = sumf(float(sum(1 for i in range(n_s) if placed[i] == samples[i][2])))
    pass_over(from_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        pass_over(from_host)
    torch.cuda.synchronize()
    dte = maxf(time.perf_counter() - t0)
    for a, b in hp:
        L.pm_host_free(a); L.pm_host_free(b)
    total_samples = sumf(float(n_s))
    if rank == 0:
        nr = S0.n_nodes * w["n_reads"]
        bases = int(sum(int(o[-1]) for _, o, _ in samples[:1])) 
        pk, pk_src = peaks()
        alg = bases + 12 * S0.n_deltas + 8 * (S0.n_nodes + 1) + 4 * S0.n_nodes + 40 * S0.n_nodes
        ms_sample = 1e3 * dt / (steps * n_s)
        line = {"metric": "placement nodes x reads scored per second", "value": total_samples * steps * nr / dt, "unit": "node*reads/s", "n_gpus": world,
                "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64+f64",
                "data": "synthetic",
                "config": {"workload": w["name"], "n_nodes": S0.n_nodes, "n_deltas": S0.n_deltas, "samples": int(total_samples), "samples_per_gpu": n_s,
                           "reads_per_sample": w["n_reads"], "read_len": w["read_len"], "k": S0.k, "s": S0.s, "l": S0.l, "host_threads_per_gpu": n_threads,
                           "parallelism": f"sample-sharded over {world} GPU(s), {n_threads} workspaces (streams) per GPU on one shared index, no collective",
                           "l2": "a pass touches every sample of the rank's share once (>= 1.7 GB per 128 samples): nothing of a sample survives in L2 until its next turn"},
                "samples_per_second": total_samples * steps / dt, "ms_per_sample_per_gpu": ms_sample,
                "placed_on_truth_leaf": int(on_truth), "placed_total": int(total_samples),
                "e2e": {"value": total_samples * steps * nr / dte, "unit": "node*reads/s", "samples_per_second": total_samples * steps / dte,
                        "h2d_bytes_per_step": int(total_samples) * (bases + 8 * (w["n_reads"] + 1)), "d2h_bytes_per_step": int(total_samples) * 452,
                        "ms_per_step": 1e3 * dte / steps, "span": "buffers_to_result per sample: reads + offsets in pinned host memory -> best nodes + tie lists"},
                "gpu_launches": int(launches), "gpu_launches_per_sample": launches / (steps * total_samples),
                "roofline": {"bound": "hbm", "kernel": "whole place stage of one sample (latency regime: ~45 MB algorithmic per sample)", "achieved": alg / (ms_sample * 1e-3) / 1e9,
                             "peak": float(pk["hbm_gbs"]), "unit": "GB/s", "frac": alg / (ms_sample * 1e-3) / 1e9 / float(pk["hbm_gbs"]), "traffic": None, "peak_source": pk_src},
                "clocks": clocks}
        if not args.no_cpu_baseline:
            try:
                from oracle import ref
                if ref.available():
                    with tempfile.TemporaryDirectory() as td:
                        class One:   # one sample of the batch for the reference arm
                            pass
                        o = One(); o.__dict__.update(S0.__dict__); o.reads, o.read_offsets = samples[0][0], samples[0][1]
                        threads = os.cpu_count() or 1
                        reference_step(o, w["n_reads"], threads, td)
                        dtr, rr = reference_step(o, w["n_reads"], threads, td)
                        sp = reference_spans(dtr, rr)
                        g = wss[0].place(samples[0][0], samples[0][1], params)
                        line["cpu_baseline"] = {"value": nr / (sp["buffers_to_result_ms"] * 1e-3), "unit": "node*reads/s", "cores": threads, "kind": "reference", "spans": sp,
                                                "sample": f"ONE sample of the batch ({w['n_reads']} reads) through the reference placeLite with {threads} host threads; value = its buffers_to_result span",
                                                "gpu_same_sample": {"best_nodes_tie_lists_scores_agree": same_placement([g.best_index[m] for m in pm.METRICS], [g.tied[m] for m in pm.METRICS],
                                                                    [g.best_score[m] for m in pm.METRICS], rr["best_index"], rr["tied"], rr["best_score"])}}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": "node*reads/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_multi(args, pm, S, w, host, params, alg, pk, pk_src, world, rank, local):
    """N > 1: ONE sample over the N GPUs (BASELINE configs[2]): node range sharded, reads sliced for seeding, seed table hash-partitioned
    (pm_comm over NCCL, panmap_b200/csrc/pm_multi.cu) -> strong scaling; this is `value`.  The sample-sharded batch mode (one replica and
    one sample per GPU per step, no collective; reference runBatchPlacement, main.cpp:1464-1666) is timed after it and reported under
    "batch_mode"."""
    import torch
    import torch.distributed as dist
    from panmap_b200 import distributed as pmd
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    dev = torch.device("cuda", local)
    n = w["n_reads"]
    nodes_reads = S.n_nodes * n
    L = pm.lib()
    steps, warm = args.steps, max(args.warmup, 3)

    def maxf(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- one sample over N GPUs ----
    index = pm.Index(host, device=local, shard=rank, n_shards=world)
    ws = pm.Workspace(index)
    comm = pmd.make_comm(ws)
    transport = [""]
    reads, off = pmd.slice_reads(S.reads, S.read_offsets, rank, world)
    ws.upload(reads, off)
    for _ in range(warm):
        comm.place_sharded_resident(params, full=False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ws.stage_timers(False)                                     # the timed region: only the two events around the whole placement
    dev_sum = 0.0
    dist.barrier(); torch.cuda.synchronize()
    launches0 = pm.launch_count()
    t0 = time.perf_counter()
    for _ in range(steps):
        dev_sum += comm.place_sharded_resident(params, full=False).stage_ms[7]   # returns after the result is on the host (library stream synchronised)
    torch.cuda.synchronize()
    wall_local = time.perf_counter() - t0
    launches = pm.launch_count() - launches0
    dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    per = maxf(dev_sum / steps)                                # CUDA-event time of a step on the library stream, max over ranks
    wall = maxf(wall_local * 1e3 / steps)
    ws.stage_timers(True)                                      # the per-stage breakdown: a second pass with the stage timers on
    stage = np.zeros(8)
    dist.barrier()
    for _ in range(steps):
        r = comm.place_sharded_resident(params, full=False)
        stage += np.array(list(r.stage_ms))
    ws.stage_timers(False)
    stage /= steps
    res = comm.place_sharded_resident(params)
    sent, recv = comm.last_traffic()
    transport[0] = comm.transport()
    stage_max = [maxf(float(x)) for x in stage]
    # e2e: every rank's slice in pinned host memory -> result on every rank
    hp_reads = pinned_copy(L, reads); hp_off = pinned_copy(L, off)
    for _ in range(3):
        comm.place_sharded_raw(hp_reads, hp_off, off.size - 1, params)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        comm.place_sharded_raw(hp_reads, hp_off, off.size - 1, params)
    e2e_ms = maxf((time.perf_counter() - t0) * 1e3 / steps)
    packed = pm.host_pack_reads(reads, off)
    hp_packed = pinned_copy(L, packed)
    for _ in range(3):
        comm.place_sharded_packed_raw(hp_packed, hp_off, off.size - 1, params)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        comm.place_sharded_packed_raw(hp_packed, hp_off, off.size - 1, params)
    e2ep_ms = maxf((time.perf_counter() - t0) * 1e3 / steps)
    L.pm_host_free(hp_reads); L.pm_host_free(hp_off); L.pm_host_free(hp_packed)
    h2d_local = int(reads.nbytes) + 8 * int(off.size)
    t = torch.tensor([h2d_local, launches], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    h2d_all, launches_all = int(t[0].item()), int(t[1].item())
    sharded_placed = {m: int(res.best_index[m]) for m in pm.METRICS}
    sharded_tied = [np.asarray(res.tied[m]) for m in pm.METRICS]
    sharded_score = [float(res.best_score[m]) for m in pm.METRICS]
    comm.close()
    del ws, index

    # ---- batch mode: one replica per GPU, every rank places its own copy of the sample (weak scaling, no collective) ----
    batch = None
    try:
        index = pm.Index(host, device=local)
        ws = pm.Workspace(index)
        ws.upload(S.reads, S.read_offsets)
        for _ in range(3):
            ws.place_resident(params, full=False)
        dist.barrier(); torch.cuda.synchronize()
        k2 = max(3, min(steps, 10))
        dev_ms = 0.0
        for _ in range(k2):
            dev_ms += ws.place_resident(params, full=False).stage_ms[7]
        torch.cuda.synchronize(); dist.barrier()
        bper = maxf(dev_ms / k2)
        one = ws.place_resident(params)
        same = same_placement([one.best_index[m] for m in pm.METRICS], [one.tied[m] for m in pm.METRICS], [one.best_score[m] for m in pm.METRICS],
                              [sharded_placed[m] for m in pm.METRICS], sharded_tied, sharded_score, rtol=0.0)
        same = maxf(0.0 if same else 1.0) == 0.0
        batch = {"ms_per_step": bper, "value": world * nodes_reads / (bper * 1e-3), "scaling": "weak", "samples_per_step": world,
                 "parallelism": f"{world} replicas of the index, one sample per GPU per step, no collective (BASELINE configs[4] mode on configs[2] shapes)"}
    except Exception as e:  # reported, never fatal
        batch = {"error": str(e)}; same = None
    if rank == 0:
        value = nodes_reads / (per * 1e-3)
        names = ["table setup", "seeding of the read slice + partition export", "all-to-all + partition import/finalize + pairs all-gather + gathered finalize",
                 "node_deltas (own node range)", "prefix_scores", "records + records all-gather + chain + ties", "tie heads all-gather + d2h + reset"]
        line = {"metric": "placement nodes x reads scored per second", "value": value, "unit": "node*reads/s", "n_gpus": world, "steps": steps,
                "warmup": warm, "ms_per_step": per, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64+f64",
                "data": "synthetic",
                "config": {"workload": w["name"], "n_nodes": S.n_nodes, "n_deltas": S.n_deltas, "n_reads": n, "samples_per_step": 1, "k": S.k, "s": S.s, "l": S.l,
                           "parallelism": f"one sample over {world} GPUs: node range sharded (delta-balanced DFS ranges + ancestors), reads sliced for seeding, seed table "
                                          f"hash-partitioned (NCCL all-to-all), finalized (count, seed id) pairs / records / tie heads by fixed-capacity all-gathers, one D2H",
                           "l2": "working set exceeds L2; no explicit flush", "placed": sharded_placed, "truth_node": int(S.truth)},
                "wall_ms_per_step": wall,
                "stage_ms": {nm: stage_max[i] for i, nm in enumerate(names[:7])},
                "e2e": {"value": nodes_reads / (e2e_ms * 1e-3), "unit": "node*reads/s", "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": world * (432 + 4 * 336 * world),
                        "ms_per_step": e2e_ms, "span": "buffers_to_result: every rank's read slice in pinned host memory -> best nodes + tie lists on every rank"},
                "e2e_packed": {"value": nodes_reads / (e2ep_ms * 1e-3), "unit": "node*reads/s", "ms_per_step": e2ep_ms,
                               "span": "the same with every rank's slice as 4-bit codes (pm_place_sharded_packed)"},
                "gpu_launches": launches_all, "gpu_launches_per_step": launches_all / steps,
                "collective_bytes_per_step_rank0": {"sent": int(sent), "received": int(recv)}, "transport": transport[0],
                "same_result_as_one_gpu": same,
                "roofline": {"bound": "hbm", "kernel": "whole place stage (one sample, all ranks)", "achieved": alg["total"] / (per * 1e-3) / 1e9, "peak": float(pk["hbm_gbs"]) * world,
                             "unit": "GB/s", "frac": alg["total"] / (per * 1e-3) / 1e9 / (float(pk["hbm_gbs"]) * world), "traffic": None, "peak_source": pk_src},
                "batch_mode": batch,
                "clocks": clocks}
        print(json.dumps(line))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
