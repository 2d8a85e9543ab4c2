"""the streaming file receiver (rtlamr_b200/host/receiver.*, ertgpu_decode_file).
Chunked streaming must equal one-shot decoding, and the prev/next digest dedup must behave like
main.go:244-260,292 applied to the reference pipeline's messages."""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, ROOT
from helpers import synth_stream, whole_blocks

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "rtlamr_b200", "ertgpu_decode_file")


def run_cli(path, *args):
    if not os.path.exists(CLI):
        subprocess.run(["make", "-C", os.path.join(ROOT, "rtlamr_b200", "host")], check=True, stdout=subprocess.DEVNULL)
    out = subprocess.run([CLI, *args, path], capture_output=True, text=True, timeout=300, check=True)
    return [l for l in out.stdout.splitlines() if l.startswith("{")], out.stderr


def test_streaming_in_small_calls_equals_one_shot(built):
    sample = os.path.join(GOLDEN, "sample_cl78.bin")
    a, _ = run_cli(sample, "-msgtype=scm", "-symbollength=78", "-blocks=4096")
    b, _ = run_cli(sample, "-msgtype=scm", "-symbollength=78", "-blocks=7")
    c, err = run_cli(sample, "-msgtype=scm", "-symbollength=78", "-blocks=1")
    assert a == b == c and len(a) == 14
    assert "69 blocks" in err
    assert a[0] == "{Block:4 Idx:1031 SCM:{ID:17580293 Type: 8 Tamper:{Phy:01 Enc:01} Consumption:  111414 CRC:0xD005}}"


def test_batch_of_files_equals_each_file_alone(built, tmp_path):
    """Several files on one command line are independent streams through one decoder: zeroed history, block
    numbers from 0 and an empty dedup set at every file boundary."""
    sample = os.path.join(GOLDEN, "sample_cl78.bin")
    other = tmp_path / "second.bin"
    raw = np.fromfile(sample, dtype=np.uint8)
    raw[16384 * 3:].tofile(other)                      # the same capture entered three blocks late
    a, _ = run_cli(sample, "-msgtype=scm", "-symbollength=78")
    b, _ = run_cli(str(other), "-msgtype=scm", "-symbollength=78")
    out = subprocess.run([CLI, "-msgtype=scm", "-symbollength=78", "-blocks=5", sample, str(other), sample],
                         capture_output=True, text=True, timeout=300, check=True)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines == a + b + a
    assert len(b) >= 10 and b != a
    assert out.stderr.count("messages (") == 3 and "second.bin: " in out.stderr


def test_cross_block_dedup_matches_main_go_rule(built, tmp_path):
    from rtlamr_b200 import synth
    mt, cl = "scm", 72
    o = oracle.Oracle(mt, cl)
    bs, buf, sl = o.cfg.block_size, o.cfg.buffer_length, o.cfg.symbol_length
    n = 1 << 21
    pk, truth = synth.make_packets(mt, cl, n, seed=7, spacing=1 << 18)
    # move every other packet so that its ~69 matching sample phases straddle a block boundary:
    # the preamble starts at stream bit S + SL, reported with Idx = (S + SL + BUF) mod BS
    for i in range(1, len(pk), 2):
        s0 = int(pk["start_sample"][i])
        idx = (s0 + sl + buf) % bs
        pk["start_sample"][i] = s0 + ((bs - 2) - idx)          # Idx of the true start = BS - 2
    iq = synth.host_fill(0, n, 0x5EED0001, pk)
    iq = whole_blocks(iq, o.cfg.block_size2)
    path = tmp_path / "stream.bin"
    iq.tofile(path)
    _, msgs = o.decode(iq)
    # main.go:244-260,292 on the reference pipeline's messages
    by_block = {}
    for m in msgs:
        by_block.setdefault(m.block, []).append((m.proto, m.meter_type, m.meter_id, m.checksum))
    kept, prev, prev_block = 0, set(), -2
    for b in sorted(by_block):
        if b != prev_block + 1:
            prev = set()
        nxt = set()
        for d in by_block[b]:
            nxt.add(d)
            if d not in prev:
                kept += 1
        prev, prev_block = nxt, b
    uniq, _ = run_cli(str(path), f"-msgtype={mt}", f"-symbollength={cl}", "-blocks=64")
    alln, _ = run_cli(str(path), f"-msgtype={mt}", f"-symbollength={cl}", "-blocks=64", "-blockdedup=false")
    assert len(alln) == len(msgs)
    assert len(uniq) == kept
    assert kept < len(msgs), "the stream should contain at least one packet that spans two blocks"
    ids = {int(re.search(r"(?:ID|ERTSerialNumber): *(\d+)", l).group(1)) for l in uniq}
    assert ids == {m.meter_id for m in msgs}
