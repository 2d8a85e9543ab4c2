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
