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


def test_filter_chain_on_the_cpu(built):
    """flags.go:226-259 + parse.go:126-149 + main.go:236-260 in the C++ mirror, no device: -filterid, -filtertype and -unique
    on the oracle's candidates of a synthetic stream, against the same rules applied in Python to the unfiltered messages
    (chain order filterid, filtertype, unique; a rejected message leaves no digest behind)."""
    from rtlamr_b200 import host, synth
    mt, cl = "scm,scm+,idm", 72
    o = oracle.Oracle(mt, cl)
    n = 1 << 22
    pk, truth = synth.make_packets(mt, cl, n, seed=5, spacing=1 << 17)
    iq = whole_blocks(synth.host_fill(0, n, 0x5EED0001, pk), o.cfg.block_size2)
    cands, msgs = o.decode(iq)
    one = _cand_records(cands)
    # the same transmissions three times, far apart in block numbers: every meter repeats its checksum (what -unique drops)
    gap = int(one["block"].max()) + 3
    rec = np.concatenate([one, one, one])
    rec["block"][len(one):2 * len(one)] += gap
    rec["block"][2 * len(one):] += 2 * gap
    p = host.Parsers(mt, cl)
    allm, _ = p.parse_dedup(rec, block_dedup=False)
    assert len(allm) == 3 * len(msgs) > 60
    ids = sorted({m.meter_id for m in allm})
    types = sorted({m.meter_type for m in allm})
    want_ids, want_types = ids[::2], types[:1]

    def model(filterid, filtertype, unique, block_dedup):
        last, out, rejected, dropped = {}, [], 0, 0
        prev, prev_block, i = set(), -2, 0
        while i < len(allm):
            b = allm[i].block
            if b != prev_block + 1:
                prev = set()
            nxt = set()
            while i < len(allm) and allm[i].block == b:
                m = allm[i]
                i += 1
                ok = (not filterid or m.meter_id in filterid) and (not filtertype or m.meter_type in filtertype)
                if ok and unique:
                    if last.get(m.meter_id) == m.checksum:
                        ok = False
                    else:
                        last[m.meter_id] = m.checksum
                if not ok:
                    rejected += 1
                    continue
                dg = (m.msgtype, m.meter_type, m.meter_id, m.checksum)
                nxt.add(dg)
                if block_dedup and dg in prev:
                    dropped += 1
                    continue
                out.append(m)
            prev, prev_block = nxt, b
        return out, dropped, rejected

    for fid, ftype, uniq, dd in [(want_ids, [], False, True), ([], want_types, False, True), (want_ids, want_types, True, True),
                                 ([], [], True, False), ([], [], True, True), (ids, types, False, False)]:
        got, dropped, rejected = p.parse_filtered(rec, ",".join(map(str, fid)), ",".join(map(str, ftype)), uniq, dd)
        want, wdropped, wrejected = model(set(fid), set(ftype), uniq, dd)
        assert [(m.block, m.idx, m.text) for m in got] == [(m.block, m.idx, m.text) for m in want]
        assert (dropped, rejected) == (wdropped, wrejected)
    # -unique alone: every meter is reported once per distinct checksum run
    got, _, rejected = p.parse_filtered(rec, unique=True, block_dedup=False)
    assert rejected > 0 and len(got) + rejected == len(allm)
    with pytest.raises(RuntimeError):
        p.parse_filtered(rec, filterid="12,x")      # strconv.ParseUint error (flags.go:213-217)
    p.close()


def test_plain_and_csv_encoders_on_the_cpu(built, sample_iq):
    """parse.go:103-129 (LogMessage), flags.go:261-272 (PlainEncoder), csv/csv.go:27-38 over encoding/csv's rules: the lines
    the C++ mirror prints for the golden capture's messages against lines built here from each message's String() / Record()."""
    import csv
    import io
    from rtlamr_b200 import host
    mt, cl = "scm", 78
    o = oracle.Oracle(mt, cl, oracle.SEARCH_EXACT)
    cands, msgs = o.decode(whole_blocks(sample_iq, o.cfg.block_size2))
    rec = _cand_records(cands)
    p = host.Parsers(mt, cl)
    allm, _ = p.parse_dedup(rec, block_dedup=False)
    assert len(allm) == len(msgs) >= 14      # the parser's own per-block dedup leaves the 14 golden messages
    t, ns, off, ln = 1700000000, 123456000, 4096, 8192          # 2023-11-14T22:13:20.123456Z
    plain = p.encode(rec, "plain", t, ns, off, ln, True).splitlines()
    plain_off = p.encode(rec, "plain", t, ns, off, ln, False).splitlines()
    assert plain == ["{Time:2023-11-14T22:13:20.123 SCM:%s}" % m.text for m in allm]
    assert plain_off == ["{Time:2023-11-14T22:13:20.123 Offset:4096 Length:8192 SCM:%s}" % m.text for m in allm]
    buf = io.StringIO()
    w = csv.writer(buf, lineterminator="\n", quoting=csv.QUOTE_MINIMAL)
    for m in allm:
        w.writerow(["2023-11-14T22:13:20.123456Z", "4096", "8192", *m.record])
    assert p.encode(rec, "csv", t, ns, off, ln) == buf.getvalue()
    # whole seconds: RFC3339Nano drops the fraction
    assert p.encode(rec, "csv", t, 0, 0, 0).startswith("2023-11-14T22:13:20Z,0,0," + ",".join(allm[0].record) + "\n")
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
