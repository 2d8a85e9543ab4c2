"""Full-size checks at BASELINE.json's configs[1] size (scm, ChipLength 72, 1 GiB of IQ on one GPU):
size-independent properties (every injected packet decodes, one call == many calls) plus bit-exact
comparison with the CPU oracle on random windows of the stream (SURVEY.md section 8d correctness gate)."""
import numpy as np
import pytest

import oracle
from helpers import cand_key_gpu, cand_key_oracle
from rtlamr_b200 import capi, shard, synth

pytestmark = pytest.mark.gpu

MT, CL, SEED = "scm", 72, 0x5EED0002


@pytest.fixture(scope="module")
def big(built):
    import torch
    nbytes = 1 << 30
    h = capi.new_decoder(MT, CL, max_blocks_per_call=nbytes // 8192, max_candidates=1 << 20)
    nsamples = nbytes // 2
    pk, truth = synth.make_packets(MT, CL, nsamples, seed=1, spacing=1 << 20)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    capi.synth_fill(0, d.data_ptr(), 0, nsamples, SEED, pk)
    stream = torch.cuda.Stream()
    h.decode_device_async(d.data_ptr(), nbytes, 0, stream.cuda_stream)
    got = h.fetch(1 << 18)
    yield h, d, pk, truth, got
    h.close()


def test_every_injected_packet_is_recovered(big):
    h, d, pk, truth, got = big
    valid = got[got["check_mask"] != 0]
    seen = {bytes(r["bytes"][:12]) for r in valid}
    usable = [t for t in truth if t.start_sample + h.cfg.buffer_length < (1 << 29)]
    assert len(usable) >= 511
    missing = [t for t in usable if t.data not in seen]
    assert not missing
    assert seen <= {t.data for t in truth}          # nothing CRC-valid that was not injected
    # candidates are sorted by (block, preamble, idx) and unique
    key = got["block"].astype(np.int64) * 8192 + got["idx"]
    assert np.all(np.diff(key) > 0)


def test_random_windows_match_the_oracle_bit_for_bit(big):
    h, d, pk, truth, got = big
    bs, bs2, pkl = h.cfg.block_size, h.cfg.block_size2, h.cfg.packet_length
    halo = shard.halo_blocks(bs, pkl)
    nblocks = (1 << 30) // bs2
    rng = np.random.default_rng(5)
    # windows: the start of the stream, two random places that contain a packet, the very end
    starts = [0, int(truth[137].start_sample // bs) - 100, int(truth[401].start_sample // bs) - 300, nblocks - 1024]
    for w0 in starts:
        w1 = min(nblocks, w0 + 1024)
        f0 = max(0, w0 - halo)
        iq = d[f0 * bs2:w1 * bs2].cpu().numpy()
        # the device generator and the host generator agree on this window
        assert np.array_equal(iq[:1 << 16], synth.host_fill(f0 * bs, 1 << 15, SEED, pk))
        o = oracle.Oracle(MT, CL, oracle.SEARCH_GO)
        cands, msgs = o.decode(iq, cand_cap=1 << 18)
        want = sorted((c.block + f0, c.preamble_id, c.idx, c.data[:12]) for c in cands if c.block + f0 >= w0)
        sel = got[(got["block"] >= w0) & (got["block"] < w1)]
        have = sorted((int(r["block"]), int(r["preamble_id"]), int(r["idx"]), r["bytes"][:12].tobytes()) for r in sel)
        assert have == want, (w0, len(have), len(want))
        if w0 != 0:
            assert len(want) > 0
        # Quantized of the window's last block, bit for bit
        assert np.array_equal(h.tap(capi.TAP_QUANTIZED, w1 - 1), o.quantized())


def test_one_call_equals_many_calls_at_full_size(big):
    import torch
    h, d, pk, truth, got = big
    h2 = capi.new_decoder(MT, CL, max_blocks_per_call=20000, max_candidates=1 << 20)
    bs2 = h2.cfg.block_size2
    nblocks = (1 << 30) // bs2
    stream = torch.cuda.Stream()
    parts, off = [], 0
    for n in (1, 19999, 3, 20000, 20000, 20000, 20000, 20000, 11070):
        n = min(n, nblocks - off)
        if n <= 0:
            break
        h2.decode_device_async(d.data_ptr() + off * bs2, n * bs2, 0, stream.cuda_stream)
        parts.append(h2.fetch(1 << 18))
        off += n
    assert off == nblocks
    split = np.concatenate(parts)
    assert len(split) == len(got)
    for f in ("block", "idx", "preamble_id", "check_mask", "bytes"):
        assert np.array_equal(split[f], got[f]), f
    h2.close()
