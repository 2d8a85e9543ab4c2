"""scratch: e2e throughput of ertgpu_decode on PAGEABLE input (what a Go slice is) vs the number of host staging threads (ERTGPU_STAGE_THREADS)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rtlamr_b200 import capi, synth
mt, cl = "scm", 72
nbytes = 1 << 30
capi.bind_host_thread(0)
for thr in (1, 2, 4, 8, 12, 16, 24, 32):
    os.environ["ERTGPU_STAGE_THREADS"] = str(thr)
    h = capi.new_decoder(mt, cl, device=0, max_blocks_per_call=nbytes // 8192, max_candidates=1 << 20)
    pg = np.random.default_rng(1).integers(100, 156, nbytes, dtype=np.uint8)
    for _ in range(2):
        h.reset(); h.decode((pg.ctypes.data, nbytes), capi.DECODE_ONLY_VALID, 1 << 20)
    t0 = time.perf_counter()
    for _ in range(3):
        h.reset(); h.decode((pg.ctypes.data, nbytes), capi.DECODE_ONLY_VALID, 1 << 20)
    dt = (time.perf_counter() - t0) / 3
    print(f"stage threads {thr:2d}: {nbytes / 2 / dt / 1e9:.2f} Gsamples/s ({nbytes / dt / 1e9:.1f} GB/s)", flush=True)
    h.close()
