import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch, torch.distributed as dist
import bench, panmap_b200 as pm
from panmap_b200 import distributed as pmd
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dist.init_process_group('nccl')
S, w = bench.make_workload('c3'); n = w['n_reads']
host = pm.HostIndex(S.hash, S.parent, S.child, S.offsets, S.parent_index, S.k, S.s, S.t, S.l)
index = pm.Index(host, device=local, shard=rank, n_shards=world); ws = pm.Workspace(index)
lo, hi = (n * rank) // world, (n * (rank + 1)) // world
off = S.read_offsets[lo:hi + 1] - S.read_offsets[lo]; reads = S.reads[int(S.read_offsets[lo]):int(S.read_offsets[hi])]
ws.upload(reads, off); dev = torch.device('cuda', local); p = pm.PlaceParams()
for _ in range(3): pmd.place_sharded(ws, reads, off, n, p, device=dev, resident=True)
os.environ['PM_TRACE'] = '1'
for _ in range(4): pmd.place_sharded(ws, reads, off, n, p, device=dev, resident=True)
dist.destroy_process_group()
