"""scratch: quick device-resident timing of the decode pipeline (not the contract bench)."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtlamr_b200 import capi, synth

mt = sys.argv[1] if len(sys.argv) > 1 else "scm"
cl = int(sys.argv[2]) if len(sys.argv) > 2 else 72
gib = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
nbytes = int(gib * (1 << 30))
h = capi.new_decoder(mt, cl, max_blocks_per_call=nbytes // 1024 + 1, max_candidates=1 << 20)
bs2 = h.cfg.block_size2
nbytes = nbytes // bs2 * bs2
nsamples = nbytes // 2
pk, truth = synth.make_packets(mt, cl, nsamples, seed=1, spacing=1 << 20)
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
t0 = time.time(); capi.synth_fill(0, d.data_ptr(), 0, nsamples, 0x5EED0001, pk); torch.cuda.synchronize(); print("synth %.3fs" % (time.time() - t0))
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
st = ts.cuda_stream
assert st != 0
for it in range(3):
    h.reset(); h.decode_device_async(d.data_ptr(), nbytes, capi.DECODE_ONLY_VALID, st); got = h.fetch(1 << 16)
print("cands/valid", h.last_counts(), "returned", len(got), "truth", len(truth))
ids = set()
for r in got:
    ids.add(bytes(r["bytes"][:12]))
times = []
h.set_stage_timing(True)
for it in range(5):
    h.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); h.decode_device_async(d.data_ptr(), nbytes, capi.DECODE_ONLY_VALID, st); e1.record(); h.fetch(1 << 16)
    times.append(e0.elapsed_time(e1))
ms = min(times)
print("stages", {k: round(v,3) for k,v in h.last_stage_ms().items()})
print(f"{mt} cl={cl} {gib} GiB: {ms:.3f} ms  -> {nsamples/ms/1e3:.1f} MS/s, {nbytes/ms/1e6:.1f} GB/s  all={['%.3f'%t for t in times]}")
