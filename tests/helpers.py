"""Shared helpers for the parity tests: run the oracle and the CUDA path on the same bytes."""
from __future__ import annotations

import numpy as np

import oracle
from rtlamr_b200 import capi, synth


def whole_blocks(iq: np.ndarray, bs2: int) -> np.ndarray:
    return iq[: (iq.size // bs2) * bs2]


def oracle_run(msgtypes, cl, iq, search=None):
    if search is None:
        search = oracle.SEARCH_GO if (2 * cl) % 8 == 0 else oracle.SEARCH_EXACT
    o = oracle.Oracle(msgtypes, cl, search)
    iq = whole_blocks(iq, o.cfg.block_size2)
    cands, msgs = o.decode(iq, cand_cap=1 << 18, msg_cap=1 << 16)
    return o, cands, msgs


def cand_key_oracle(c, nbytes, pk_symbols):
    data = bytearray(c.data[:nbytes])
    if pk_symbols % 8:  # stale bits of the reused d.pkt in the last byte (decode.go:363-366): not comparable
        data[-1] &= (0xFF >> (8 - pk_symbols % 8))
    return (c.block, c.preamble_id, c.idx, bytes(data))


def cand_key_gpu(row, nbytes, pk_symbols):
    data = bytearray(row["bytes"][:nbytes].tobytes())
    if pk_symbols % 8:
        data[-1] &= (0xFF >> (8 - pk_symbols % 8))
    return (int(row["block"]), int(row["preamble_id"]), int(row["idx"]), bytes(data))


def synth_stream(msgtypes, cl, nsamples, seed=0x5EED0001, spacing=1 << 19, pkt_seed=7):
    pk, truth = synth.make_packets(msgtypes, cl, nsamples, seed=pkt_seed, spacing=spacing)
    return synth.host_fill(0, nsamples, seed, pk), pk, truth


# ---- the integrity screens, restated with the oracle's own crc/gf functions (crc/crc.go, r900/gf/gf.go) ----
def screen_expected(proto_name: str, data: bytes, digits=None) -> bool:
    """What bit i of check_mask must be for a candidate filed under parser `proto_name`:
    scm/scm.go:76, scmplus/scmplus.go:77, idm/idm.go:77-87, netidm/netidm.go:88-98, r900/r900.go:199-221."""
    if proto_name == "scm":
        return oracle.crc_checksum(0, data[2:12], 0x6F63) == 0
    if proto_name == "scm+":
        return oracle.crc_checksum(0xFFFF, data[2:16], 0x1021) == 0x1D0F
    if proto_name in ("idm", "netidm"):
        return (oracle.crc_checksum(0xFFFF, data[4:92], 0x1021) == 0x1D0F
                and oracle.crc_checksum(0xFFFF, data[9:13] + data[88:90], 0x1021) == 0x1D0F)
    if proto_name in ("r900", "r900bcd"):
        sym = [digits[2 * k] * 6 + digits[2 * k + 1] for k in range(21)]
        if max(sym) > 31:
            return False
        msg = bytes(sym[:16]) + bytes(10) + bytes(sym[16:])
        return oracle.gf32_syndrome(msg, 5, 29) == bytes(5)
    raise ValueError(proto_name)


PREAMBLES = {"scm": "111110010101001100000", "scm+": "0001011010100011", "idm": "01010101010101010001011010100011",
             "netidm": "01010101010101010001011010100011", "r900": "00000000000000001110010101100100",
             "r900bcd": "00000000000000001110010101100100"}


def assert_check_masks_exact(got, msgtypes):
    """BOTH directions: bit i of check_mask is set iff the candidate is filed under parser i's preamble and
    passes that parser's integrity check (recomputed here from the returned bytes / digits)."""
    if isinstance(msgtypes, str):
        msgtypes = [m.strip() for m in msgtypes.split(",") if m.strip()]
    pre_ids, order = {}, []
    for m in msgtypes:          # distinct preambles in registration order (decode.go:121-124)
        if PREAMBLES[m] not in order:
            order.append(PREAMBLES[m])
        pre_ids[m] = order.index(PREAMBLES[m])
    cache = {}
    for r in got:
        data, dig, pid = r["bytes"].tobytes(), r["r900_digits"].tobytes(), int(r["preamble_id"])
        want = 0
        for i, m in enumerate(msgtypes):
            if pre_ids[m] != pid:
                continue
            key = (m, data if m not in ("r900", "r900bcd") else dig)
            if key not in cache:
                cache[key] = screen_expected(m, data, list(dig))
            if cache[key]:
                want |= 1 << i
        assert int(r["check_mask"]) == want, (int(r["block"]), int(r["idx"]), pid, int(r["check_mask"]), want)
