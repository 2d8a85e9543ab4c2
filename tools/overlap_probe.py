"""scratch: do two independent calls on two streams overlap (Search/extract of one under the demod of the other)?"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtlamr_b200 import capi, synth

mt, cl = "scm", 72
nbytes = 1 << 29
hs = [capi.new_decoder(mt, cl, max_blocks_per_call=nbytes // 1024 + 1, max_candidates=1 << 20) for _ in range(2)]
big = capi.new_decoder(mt, cl, max_blocks_per_call=2 * nbytes // 1024 + 1, max_candidates=1 << 20)
ns = nbytes // 2
pk, truth = synth.make_packets(mt, cl, 2 * ns, seed=1, spacing=1 << 20)
d = torch.empty(2 * nbytes, dtype=torch.uint8, device="cuda")
capi.synth_fill(0, d.data_ptr(), 0, 2 * ns, 0x5EED0001, pk)
torch.cuda.synchronize()
s = [torch.cuda.Stream(), torch.cuda.Stream()]
def two():
    for i in range(2):
        hs[i].reset(); hs[i].decode_device_async(d.data_ptr() + i * nbytes, nbytes, capi.DECODE_ONLY_VALID, s[i].cuda_stream)
    for i in range(2): hs[i].last_counts()
def one():
    big.reset(); big.decode_device_async(d.data_ptr(), 2 * nbytes, capi.DECODE_ONLY_VALID, s[0].cuda_stream); big.last_counts()
import time
for name, fn in (("one 1 GiB call", one), ("two 0.5 GiB calls on two streams", two)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(name, "min %.3f ms" % min(ts), ["%.3f" % t for t in ts])
