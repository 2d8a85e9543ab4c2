"""scratch: raw pinned H2D bandwidth vs the e2e arm (is ertgpu_decode PCIe-bound?), with and without
binding the process to the GPU's NUMA-local cores before the pinned allocation."""
import os, sys, time
import torch

def local_cpus(idx=0):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = [i * 64 + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        return cpus
    except Exception as e:
        print("nvml affinity failed:", e)
        return []

def probe(tag):
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(7)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{tag}: raw pinned H2D 1 GiB: {n/dt/1e9:.1f} GB/s")

if len(sys.argv) > 1 and sys.argv[1] == "bind":
    cpus = local_cpus(0)
    print("local cpus", cpus[:4], "...", len(cpus))
    if cpus:
        os.sched_setaffinity(0, cpus)
    probe("bound")
else:
    probe("unbound")

def probe_chunked(mib):
    n = 1 << 30
    c = mib << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); h.fill_(7)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for rep in range(2):
        for i in range(n // c):
            d[i*c:(i+1)*c].copy_(h[i*c:(i+1)*c], non_blocking=True)
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for rep in range(5):
        for i in range(n // c):
            d[i*c:(i+1)*c].copy_(h[i*c:(i+1)*c], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"chunked {mib} MiB copies: {n/dt/1e9:.1f} GB/s")

if len(sys.argv) > 1 and sys.argv[1] == "chunk":
    for m in (8, 32, 128):
        probe_chunked(m)
