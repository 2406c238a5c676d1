import ctypes as C, sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import bench, panmap_b200 as pm, torch
S, w = bench.make_workload('c3')
host = pm.HostIndex(S.hash, S.parent, S.child, S.offsets, S.parent_index, S.k, S.s, S.t, S.l)
ws = pm.Workspace(pm.Index(host)); ws.stage_timers(True); params = pm.PlaceParams()
L = pm.lib(); n = w['n_reads']; nbytes = int(S.read_offsets[-1])
hp_reads = L.pm_host_alloc(nbytes + 64); hp_off = L.pm_host_alloc(8 * (n + 1))
C.memmove(hp_reads, S.reads.ctypes.data, nbytes); C.memmove(hp_off, S.read_offsets.ctypes.data, 8 * (n + 1))
for _ in range(3): ws.place_raw(hp_reads, hp_off, n, params)
st = np.zeros(8); t0 = time.perf_counter()
for _ in range(10):
    r = ws.place_raw(hp_reads, hp_off, n, params); st += np.array(list(r.stage_ms))
print('e2e wall', (time.perf_counter() - t0) * 100, 'ms; stages', (st / 10).round(3))
# raw H2D rate
a = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); d = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
torch.cuda.synchronize()
for _ in range(3): d.copy_(a, non_blocking=True)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d.copy_(a, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print('H2D 150MB ms', e0.elapsed_time(e1) / 5, 'GB/s', nbytes / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e9)
