"""scratch: e2e (host pinned -> ertgpu_decode) throughput vs chunk size."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtlamr_b200 import capi, synth
nbytes = 1 << 30
h = capi.new_decoder("scm", 72, max_blocks_per_call=nbytes // 8192, max_candidates=1 << 20)
pk, _ = synth.make_packets("scm", 72, nbytes // 2, seed=1, spacing=1 << 20)
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
capi.synth_fill(0, d.data_ptr(), 0, nbytes // 2, 2, pk)
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); host.copy_(d); torch.cuda.synchronize()
for flags in (capi.DECODE_ONLY_VALID,):
    for _ in range(2):
        h.reset(); r = h.decode((host.data_ptr(), nbytes), flags, 1 << 17)
    t0 = time.perf_counter()
    for _ in range(5):
        h.reset(); r = h.decode((host.data_ptr(), nbytes), flags, 1 << 17)
    dt = (time.perf_counter() - t0) / 5
    print(f"chunk {os.environ.get('ERTGPU_CHUNK_MIB','32')} MiB: {nbytes/dt/1e9:.1f} GB/s  ({len(r)} cands)")
