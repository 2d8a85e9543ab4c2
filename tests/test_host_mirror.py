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
