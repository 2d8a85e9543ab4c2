"""N>1 host logic on CPU: two ranks (gloo, 127.0.0.1) each decode their shard (halo + owned blocks) of one
synthetic stream and the union must equal the single-stream result.  The decode engine here is the CPU
oracle -- the point of the test is the shard plan, the halo, the block renumbering and the gather, which
bench.py and INTEGRATION.md section 4 use unchanged around the CUDA path."""
import os
import pickle
import socket
import sys
import tempfile

import numpy as np
import pytest

from rtlamr_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_properties():
    for total, n in ((131072, 8), (1000, 3), (7, 2), (5, 8)):
        plans = shard.plan(total, n, 4096, 13824)
        assert plans[0].first_block == 0 and plans[-1].last_block == total
        assert all(a.last_block == b.first_block for a, b in zip(plans, plans[1:]))
        assert all(p.first_fed_block == max(0, p.first_block - 5) for p in plans)   # ceil(13824/4096)+1 = 5
        assert sum(p.last_block - p.first_block for p in plans) == total
    assert shard.halo_blocks(8192, 105984) == 14
    p = shard.plan(100, 2, 4096, 13824)[1]
    assert p.to_global(0) == p.first_fed_block and not p.keep(0) and p.keep(p.halo_blocks)
    # candidate at stream bit g0 is reported by block floor((g0+BUF)/BS)-1
    assert shard.plan(100, 1, 4096, 13824)[0].owns_start(4096 * 10, 4096, 17920)


def test_c_abi_plan_equals_python_plan(built):
    """ertgpu_plan_shards (the product's plan, rtlamr_b200/csrc/sharded.cpp) against the Python mirror the ranks of bench.py
    use, for every stock geometry and random stream lengths / shard counts; bad requests are refused.  No GPU involved."""
    from rtlamr_b200 import capi
    rng = np.random.default_rng(11)
    geoms = [(4096, 13824), (8192, 105984), (8192, 16704), (4096, 14976), (2048, 6144)]
    for bs, pkl in geoms:
        for _ in range(40):
            total = int(rng.integers(0, 1 << 20))
            n = int(rng.integers(1, 17))
            want = [(p.first_block, p.last_block, p.first_fed_block) for p in shard.plan(total, n, bs, pkl)]
            assert [tuple(p) for p in capi.plan_shards(total, n, bs, pkl)] == want
    for bad in ((-1, 2, 4096, 13824), (10, 0, 4096, 13824), (10, 2, 0, 13824)):
        with pytest.raises(capi.ErtGpuError):
            capi.plan_shards(*bad)


def _worker(rank, world, port, path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import oracle
    from rtlamr_b200 import shard as sh, synth

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mt, cl, n = "scm,idm", 72, 1 << 21
    pk, truth = synth.make_packets(mt, cl, n, seed=3, spacing=1 << 18)
    probe = oracle.Oracle(mt, cl)
    c = probe.cfg
    total_blocks = n // c.block_size
    plan = sh.plan(total_blocks, world, c.block_size, c.packet_length)[rank]
    # every rank generates only the bytes it is fed (counter-based generator: any range, same bytes)
    first = plan.first_fed_block * c.block_size
    iq = synth.host_fill(first, (plan.last_block - plan.first_fed_block) * c.block_size, 7, pk)
    cands, msgs = oracle.Oracle(mt, cl).decode(iq)
    mine = [(plan.to_global(x.block), x.preamble_id, x.idx, x.data) for x in cands if plan.keep(x.block)]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    dist.barrier()
    if rank == 0:
        whole = synth.host_fill(0, total_blocks * c.block_size, 7, pk)
        ref, _ = oracle.Oracle(mt, cl).decode(whole)
        want = sorted((x.block, x.preamble_id, x.idx, x.data) for x in ref)
        got = sorted(t for part in gathered for t in part)
        with open(path, "wb") as f:
            pickle.dump({"equal": got == want, "n": len(want), "per_rank": [len(p) for p in gathered]}, f)
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo(built):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "out.pkl")
        mp.spawn(_worker, args=(2, port, path), nprocs=2, join=True)
        res = pickle.load(open(path, "rb"))
    assert res["equal"] and res["n"] > 100 and all(k > 0 for k in res["per_rank"])
