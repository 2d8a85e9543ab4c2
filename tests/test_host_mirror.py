"""The C++ mirror of protocol.Decoder + parsers (rtlamr_b200/host) above the C ABI: messages must equal
the ones the reference pipeline (oracle: Decode + every parser's Parse) emits on the same bytes."""
import json
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from helpers import oracle_run, synth_stream, whole_blocks

NAMES = {"SCM": oracle.SCM, "SCM+": oracle.SCMPLUS, "IDM": oracle.IDM, "NetIDM": oracle.NETIDM, "R900": oracle.R900,
         "R900BCD": oracle.R900BCD}
CONS_FIELD = {"SCM": 4, "SCM+": 4, "IDM": 12, "NetIDM": 11, "R900": 4, "R900BCD": 4}


def test_host_library_loads_and_rejects_bad_msgtype(built):
    from rtlamr_b200 import host
    host.lib()
    with pytest.raises(RuntimeError) as e:   # parse.go:49 "invalid message type"
        host.Receiver("bogus", 72)
    assert "invalid message type" in str(e.value)


def _cand_records(cands):
    from rtlamr_b200 import capi
    rec = np.zeros(len(cands), dtype=capi.CAND_DTYPE)
    for i, c in enumerate(sorted(cands, key=lambda c: (c.block, c.preamble_id, c.idx))):
        rec[i]["block"], rec[i]["idx"], rec[i]["preamble_id"] = c.block, c.idx, c.preamble_id
        rec[i]["bytes"][:len(c.data)] = np.frombuffer(c.data, dtype=np.uint8)
    return rec


@pytest.mark.parametrize("mt,cl", [("scm", 72), ("scm+", 72), ("idm", 72), ("netidm", 48), ("scm,scm+,idm", 72), ("idm,netidm", 72)])
def test_parsers_alone_match_the_oracle_parsers(built, mt, cl):
    """No device: the C++ parsers (CRC re-check, `seen` bookkeeping, field extraction, String()/Record()) fed with the
    ORACLE's candidate list must emit the oracle's messages -- the byte-level half of the host mirror on the CPU."""
    from rtlamr_b200 import host
    iq, pk, truth = synth_stream(mt, cl, 1 << 21, spacing=1 << 18)
    o, cands, msgs = oracle_run(mt, cl, iq)
    p = host.Parsers(mt, cl)
    got = p.parse(_cand_records(cands))
    a = sorted((m.block, m.idx, NAMES[m.msgtype], m.meter_id, m.meter_type, int(m.record[CONS_FIELD[m.msgtype]]), m.checksum)
               for m in got)
    b = sorted((m.block, m.idx, m.proto, m.meter_id, m.meter_type, m.consumption, m.checksum) for m in msgs)
    assert a == b and len(a) >= 4
    ids = {(m.msgtype.lower(), m.meter_id) for m in got}
    for t in truth[:-1]:
        assert (t.msgtype, t.meter_id) in ids
    p.close()


@pytest.mark.parametrize("mt,cl", [("r900", 72), ("r900bcd", 32)])
def test_r900_parser_alone_with_oracle_digits(built, mt, cl):
    """The r900 parser's byte/digit half (r900.go:195-245: digits -> 5-bit symbols -> RS syndrome -> fields) without a
    device: candidates AND their 42 payload digits come from the oracle (its r900 `quantized` buffer, block by block)."""
    from rtlamr_b200 import capi, host
    iq, pk, truth = synth_stream(mt, cl, 1 << 20, spacing=1 << 18)
    o = oracle.Oracle(mt, cl)
    bs2 = o.cfg.block_size2
    iq = whole_blocks(iq, bs2)
    recs, want = [], []
    for b in range(iq.size // bs2):
        cands, msgs = o.decode(iq[b * bs2:(b + 1) * bs2])
        q = o.r900_quantized()
        want += msgs
        rec = _cand_records(cands)
        for r in rec:
            payload = int(r["idx"]) + o.cfg.preamble_length - o.cfg.symbol_length
            r["r900_digits"][:] = [int(q[payload + k * 4 * cl]) for k in range(capi.R900_DIGITS)]
            r["flags"] = capi.CAND_HAS_R900
        recs.append(rec)
    p = host.Parsers(mt, cl)
    got = p.parse(np.concatenate(recs))
    a = sorted((m.block, m.idx, NAMES[m.msgtype], m.meter_id, m.meter_type, int(m.record[CONS_FIELD[m.msgtype]]), m.checksum)
               for m in got)
    b_ = sorted((m.block, m.idx, m.proto, m.meter_id, m.meter_type, m.consumption, m.checksum) for m in want)
    assert a == b_ and len(a) >= 2
    with pytest.raises(RuntimeError):   # an r900 candidate without digits is a programming error, not a silent miss
        bad = np.concatenate(recs)[:1].copy()
        bad["flags"] = 0
        p.parse(bad)
    p.close()


def test_cross_block_dedup_rule_on_the_cpu(built):
    """main.go:244-260,292 (prev/next digest maps) in the C++ mirror, no device: parse the oracle's candidates of a
    stream whose packets straddle block boundaries, with and without the cross-block dedup, against the rule applied in Python to
    the oracle's messages."""
    from rtlamr_b200 import host, synth
    mt, cl = "scm", 72
    o = oracle.Oracle(mt, cl)
    bs, buf, sl = o.cfg.block_size, o.cfg.buffer_length, o.cfg.symbol_length
    n = 1 << 21
    pk, truth = synth.make_packets(mt, cl, n, seed=7, spacing=1 << 18)
    for i in range(1, len(pk), 2):                               # Idx of the true start = BS - 2: phases straddle
        s0 = int(pk["start_sample"][i])
        pk["start_sample"][i] = s0 + ((bs - 2) - (s0 + sl + buf) % bs)
    iq = whole_blocks(synth.host_fill(0, n, 0x5EED0001, pk), o.cfg.block_size2)
    cands, msgs = o.decode(iq)
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
    p = host.Parsers(mt, cl)
    rec = _cand_records(cands)
    uniq, dropped = p.parse_dedup(rec, block_dedup=True)
    alln, dropped_all = p.parse_dedup(rec, block_dedup=False)
    assert len(alln) == len(msgs) and dropped_all == 0
    assert len(uniq) == kept and dropped == len(msgs) - kept
    assert kept < len(msgs), "the stream should contain at least one packet that spans two blocks"
    assert {m.meter_id for m in uniq} == {m.meter_id for m in msgs}
    # a gap in the block numbers empties the memory: the same candidates two blocks later are all reported again
    later = rec.copy()
    later["block"] += int(rec["block"].max()) + 3
    twice, dropped2 = p.parse_dedup(np.concatenate([rec, later]), block_dedup=True)
    assert len(twice) == 2 * kept and dropped2 == 2 * dropped
    p.close()


def test_parsers_alone_on_the_golden_capture(built, sample_iq):
    """sample.bin at chip length 78, exact Search: the 853 oracle candidates give the 14 golden messages and rtlamr's
    plain formatting (scm.go:139-143)."""
    from rtlamr_b200 import host
    gold = json.load(open(os.path.join(GOLDEN, "sample_cl78_scm.json")))["exact"]["messages"]
    o, cands, msgs = oracle_run("scm", 78, sample_iq, oracle.SEARCH_EXACT)
    assert len(cands) == 853
    p = host.Parsers("scm", 78)
    got = p.parse(_cand_records(cands))
    assert [[m.block, m.idx, m.meter_id, m.meter_type, int(m.record[4])] for m in got] == [[g[0], g[1], g[3], g[4], g[5]] for g in gold]
    assert got[0].text == "{ID:17580293 Type: 8 Tamper:{Phy:01 Enc:01} Consumption:  111414 CRC:0xD005}"
    with pytest.raises(RuntimeError):   # a candidate of a preamble nobody registered
        bad = _cand_records(cands[:1])
        bad["preamble_id"] = 3
        p.parse(bad)
    p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mt,cl", [("scm", 72), ("scm+", 72), ("idm", 72), ("netidm", 48), ("r900", 72), ("r900bcd", 32),
                                   ("scm,scm+,idm,r900", 72), ("idm,netidm", 72)])
def test_messages_match_reference_pipeline(built, mt, cl):
    from rtlamr_b200 import host
    iq, pk, truth = synth_stream(mt, cl, 1 << 21, spacing=1 << 18)
    o, cands, msgs = oracle_run(mt, cl, iq)
    r = host.Receiver(mt, cl)
    assert r.cfg["BlockSize"] == o.cfg.block_size and r.cfg["BufferLength"] == o.cfg.buffer_length
    assert r.cfg["CenterFreq"] == o.cfg.center_freq and r.cfg["SampleRate"] == o.cfg.sample_rate
    got = r.decode(whole_blocks(iq, r.cfg["BlockSize2"]))
    a = sorted((m.block, m.idx, NAMES[m.msgtype], m.meter_id, m.meter_type, int(m.record[CONS_FIELD[m.msgtype]]), m.checksum)
               for m in got)
    b = sorted((m.block, m.idx, m.proto, m.meter_id, m.meter_type, m.consumption, m.checksum) for m in msgs)
    assert a == b and len(a) >= 4
    # every injected packet that fits the stream came out
    ids = {(m.msgtype.lower(), m.meter_id) for m in got}
    for t in truth[:-1]:
        assert (t.msgtype, t.meter_id) in ids
    r.close()


@pytest.mark.gpu
def test_sample_bin_messages_and_strings(built, sample_iq):
    from rtlamr_b200 import host
    gold = json.load(open(os.path.join(GOLDEN, "sample_cl78_scm.json")))["exact"]["messages"]
    r = host.Receiver("scm", 78)
    got = r.decode(whole_blocks(sample_iq, r.cfg["BlockSize2"]))
    assert [[m.block, m.idx, m.meter_id, m.meter_type, int(m.record[4])] for m in got] == [[g[0], g[1], g[3], g[4], g[5]] for g in gold]
    # rtlamr's plain formatting (scm.go:139-143)
    assert got[0].text == "{ID:17580293 Type: 8 Tamper:{Phy:01 Enc:01} Consumption:  111414 CRC:0xD005}"
    assert "ChipLength: 78" in r.log() and "Protocols: scm" in r.log()
    with pytest.raises(ValueError):   # short / ragged input: the reference panics (decode.go:222)
        r.decode(sample_iq[:1000])
    r.close()
