"""Aggregate pinned H2D ceiling of the box with N ranks copying at once (is the e2e arm's flat 4 -> 8 GPU curve
the host's limit or ours?).  Launch like bench.py:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/h2d_multi.py [bind|nobind]
Each rank: 1 GiB pinned buffer (allocated after / without binding the process to the GPU's NUMA node through
ertgpu_bind_host_thread), 8 timed 1 GiB cudaMemcpyAsync H2D, barrier on both sides; rank 0 prints one JSON line with
the per-rank GB/s and the aggregate."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from rtlamr_b200 import capi

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
mode = sys.argv[1] if len(sys.argv) > 1 else "bind"
numa = capi.bind_host_thread(local) if mode == "bind" else None
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(7)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1: dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 8
for _ in range(reps):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if world > 1: dist.barrier()
mine = torch.tensor([n * reps / dt / 1e9], dtype=torch.float64, device="cuda")
allv = [torch.zeros_like(mine) for _ in range(world)]
if world > 1:
    dist.all_gather(allv, mine)
else:
    allv = [mine]
tmax = torch.tensor([dt], dtype=torch.float64, device="cuda")
if world > 1: dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"probe": "pinned H2D, all ranks at once", "mode": mode, "n_gpus": world, "numa_rank0": numa,
                      "per_rank_gbs": [round(float(v), 1) for v in allv],
                      "aggregate_gbs": round(world * n * reps / float(tmax) / 1e9, 1)}), flush=True)
if world > 1: dist.destroy_process_group()
