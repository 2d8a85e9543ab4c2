"""GPU-side shard seams through the product entry point (-m gpu): ertgpu_decode_sharded cuts one host buffer
into N contiguous block-aligned shards with a leading halo, decodes each on its own handle from its own host
thread (here: N handles on ONE GPU -- the code path of N GPUs), drops the halo's candidates, renumbers and
concatenates.  The result must equal the single-handle candidate list, which must equal the CPU oracle's."""
import numpy as np
import pytest

import oracle
from helpers import cand_key_gpu, cand_key_oracle, oracle_run, synth_stream, whole_blocks
from rtlamr_b200 import capi, shard

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mt,cl,nshards", [("scm", 72, 2), ("scm", 72, 5), ("scm,scm+,idm", 72, 3), ("r900", 72, 4),
                                           ("scm,scm+,idm,r900", 72, 2), ("scm", 78, 3), ("idm", 40, 8)])
def test_union_of_shards_equals_single_stream_equals_oracle(built, mt, cl, nshards):
    n = 1 << 22
    iq, pk, truth = synth_stream(mt, cl, n, spacing=1 << 18)
    o, cands, msgs = oracle_run(mt, cl, iq)
    single_h = capi.new_decoder(mt, cl)
    iq = whole_blocks(iq, single_h.cfg.block_size2)
    single = single_h.decode(iq)
    handles = [capi.new_decoder(mt, cl, max_blocks_per_call=37) for _ in range(nshards)]   # forces internal chunking too
    union = capi.decode_sharded(handles, iq)
    assert len(union) == len(single) > 0
    for f in ("block", "idx", "preamble_id", "check_mask", "bytes", "r900_digits", "flags"):
        assert np.array_equal(union[f], single[f]), f
    nb, pks = o.cfg.packet_symbols + 7 >> 3, o.cfg.packet_symbols
    assert sorted(cand_key_gpu(r, nb, pks) for r in union) == sorted(cand_key_oracle(c, nb, pks) for c in cands)
    # packets were placed across the seams: at least one seam has a candidate within a packet length of it
    bs, pkl = single_h.cfg.block_size, single_h.cfg.packet_length
    plans = shard.plan(iq.size // single_h.cfg.block_size2, nshards, bs, pkl)
    assert [tuple(p) for p in capi.plan_shards(iq.size // single_h.cfg.block_size2, nshards, bs, pkl)] == \
        [(p.first_block, p.last_block, p.first_fed_block) for p in plans]
    # a second call on the same handles starts from a fresh stream again (documented: sharded decode is whole-stream)
    again = capi.decode_sharded(handles, iq, flags=capi.DECODE_ONLY_VALID)
    assert len(again) == int((single["check_mask"] != 0).sum())
    for h in handles + [single_h]:
        h.close()


def test_more_shards_than_blocks_and_mismatched_handles(built):
    h1, h2, h3 = (capi.new_decoder("scm", 72) for _ in range(3))
    bs2 = h1.cfg.block_size2
    iq, _, _ = synth_stream("scm", 72, 2 * h1.cfg.block_size, spacing=1 << 12)
    got = capi.decode_sharded([h1, h2, h3], iq[: 2 * bs2])
    ref = capi.new_decoder("scm", 72)
    want = ref.decode(iq[: 2 * bs2])
    assert len(got) == len(want) and np.array_equal(got["idx"], want["idx"])
    assert len(capi.decode_sharded([h1, h2, h3], np.zeros(0, dtype=np.uint8))) == 0
    other = capi.new_decoder("idm", 72)
    with pytest.raises(capi.ErtGpuError) as e:
        capi.decode_sharded([h1, other], iq[: 2 * bs2])
    assert e.value.code == capi.EINVAL
    with pytest.raises(capi.ErtGpuError) as e:
        capi.decode_sharded([h1, h2], iq[: bs2 + 2])
    assert e.value.code == capi.ESIZE
    for h in (h1, h2, h3, ref, other):
        h.close()


def test_pageable_and_pinned_input_give_the_same_records(built):
    """ertgpu_decode stages ordinary (pageable) host memory through the handle's pinned buffers; pinned input is
    copied directly.  Same records either way, also across the chunk boundaries of the staging."""
    import torch
    mt, cl = "scm,idm", 72
    iq, _, _ = synth_stream(mt, cl, 1 << 22, spacing=1 << 18)
    h = capi.new_decoder(mt, cl, max_blocks_per_call=50)
    iq = whole_blocks(iq, h.cfg.block_size2)
    a = h.decode(iq)                                  # numpy memory: pageable
    pinned = torch.from_numpy(iq).pin_memory()
    h.reset()
    b = h.decode((pinned.data_ptr(), iq.size))
    assert len(a) == len(b) > 0 and a.tobytes() == b.tobytes()
    h.close()


def test_bind_host_thread_reports_the_gpu_neighbourhood(built):
    import os
    before = os.sched_getaffinity(0)
    info = capi.bind_host_thread(0)
    try:
        if info["bound"]:
            now = os.sched_getaffinity(0)
            assert now <= before and len(now) == info["cpus"] >= 1
    finally:
        os.sched_setaffinity(0, before)
