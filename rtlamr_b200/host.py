"""ctypes binding of libertgpu_host.so: the C++ mirror of rtlamr's protocol.Decoder + parsers
(rtlamr_b200/host/protocol.hpp) sitting above the C ABI.  Test/bench plumbing only."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libertgpu_host.so")


class _Msg(C.Structure):
    _fields_ = [("block", C.c_int64), ("idx", C.c_int32), ("meter_id", C.c_uint32), ("meter_type", C.c_uint32),
                ("nchecksum", C.c_int32), ("checksum", C.c_uint8 * 8), ("msgtype", C.c_char * 12),
                ("text", C.c_char * 1400), ("record", C.c_char * 1400)]


@dataclass(frozen=True)
class HostMessage:
    block: int
    idx: int
    msgtype: str
    meter_id: int
    meter_type: int
    checksum: bytes
    text: str
    record: tuple


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} missing: run __graft_entry__.build()")
        L = C.CDLL(LIB_PATH)
        L.erthost_new.restype = C.c_void_p
        L.erthost_new.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_char_p, C.c_int]
        L.erthost_free.argtypes = [C.c_void_p]
        L.erthost_error.restype = C.c_char_p
        L.erthost_error.argtypes = [C.c_void_p]
        L.erthost_config.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.erthost_reset.argtypes = [C.c_void_p]
        L.erthost_decode.restype = C.c_longlong
        L.erthost_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(_Msg), C.c_longlong]
        L.erthost_log.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.erthost_new_parse_only.restype = C.c_void_p
        L.erthost_new_parse_only.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        L.erthost_parse.restype = C.c_longlong
        L.erthost_parse.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.POINTER(_Msg), C.c_longlong]
        L.erthost_parse_dedup.restype = C.c_longlong
        L.erthost_parse_dedup.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.POINTER(_Msg), C.c_longlong,
                                          C.POINTER(C.c_longlong)]
        L.erthost_parse_filtered.restype = C.c_longlong
        L.erthost_parse_filtered.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_char_p, C.c_char_p, C.c_int, C.c_int,
                                             C.POINTER(_Msg), C.c_longlong, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        L.erthost_encode.restype = C.c_longlong
        L.erthost_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_longlong, C.c_int, C.c_longlong, C.c_int,
                                     C.c_int, C.c_char_p, C.c_longlong]
        _lib = L
    return _lib


def _messages(out, n):
    return [HostMessage(m.block, m.idx, m.msgtype.decode(), m.meter_id, m.meter_type, bytes(m.checksum[:m.nchecksum]),
                        m.text.decode(), tuple(m.record.decode().split(","))) for m in out[:n]]


class Parsers:
    """NewDecoder + RegisterProtocol(NewParser(..)) WITHOUT Allocate: no device is touched.  parse() is the second
    half of Decode (decode.go:177-187 and every parser's Parse) on candidate records from any source."""

    def __init__(self, msgtypes: str, chip_length: int = 72):
        self._L = lib()
        err = C.create_string_buffer(512)
        self._h = self._L.erthost_new_parse_only(msgtypes.encode(), chip_length, err, 512)
        if not self._h:
            raise RuntimeError(err.value.decode())

    def close(self):
        if self._h:
            self._L.erthost_free(self._h)
            self._h = None

    def parse(self, cands: np.ndarray, cap: int = 4096) -> list[HostMessage]:
        """cands: structured array with capi.CAND_DTYPE, sorted by (block, preamble_id, idx)."""
        cands = np.ascontiguousarray(cands)
        assert cands.dtype.itemsize == 160
        out = (_Msg * cap)()
        n = self._L.erthost_parse(self._h, cands.ctypes.data, len(cands), out, cap)
        if n < 0:
            raise RuntimeError(self._L.erthost_error(self._h).decode())
        if n > cap:
            raise OverflowError(f"{n} messages, cap {cap}")
        return _messages(out, n)


def _parse_dedup(self, cands: np.ndarray, block_dedup: bool = True, cap: int = 4096):
    """parse() followed by the receive loop's cross-block dedup (main.go:244-260,292): (messages, duplicates dropped)."""
    cands = np.ascontiguousarray(cands)
    out = (_Msg * cap)()
    dup = C.c_longlong(0)
    n = self._L.erthost_parse_dedup(self._h, cands.ctypes.data, len(cands), 1 if block_dedup else 0, out, cap, C.byref(dup))
    if n < 0:
        raise RuntimeError(self._L.erthost_error(self._h).decode())
    return _messages(out, min(n, cap)), int(dup.value)


Parsers.parse_dedup = _parse_dedup


def _parse_filtered(self, cands, filterid: str = "", filtertype: str = "", unique: bool = False, block_dedup: bool = True,
                    cap: int = 1 << 16):
    """erthost_parse_filtered: parsers, then the receive loop's filter chain (-filterid, -filtertype, -unique in the order
    flag.Visit adds them, main.go:97-113) and the cross-block dedup (main.go:236-260).  Returns (messages, dropped by the
    dedup, rejected by the filters)."""
    cands = np.ascontiguousarray(cands)
    out = (_Msg * cap)()
    dup, flt = C.c_longlong(0), C.c_longlong(0)
    n = self._L.erthost_parse_filtered(self._h, cands.ctypes.data, len(cands), filterid.encode(), filtertype.encode(),
                                       1 if unique else 0, 1 if block_dedup else 0, out, cap, C.byref(dup), C.byref(flt))
    if n < 0:
        raise RuntimeError(self._L.erthost_error(self._h).decode())
    return _messages(out, min(n, cap)), int(dup.value), int(flt.value)


def _encode(self, cands, fmt: str = "plain", unix_seconds: int = 0, nanos: int = 0, offset: int = 0, length: int = 0,
            sample_file_is_devnull: bool = True) -> str:
    """erthost_encode: every parsed message through the plain or csv encoder (flags.go:140-151) as one string of lines."""
    cands = np.ascontiguousarray(cands)
    cap = 1 << 22
    buf = C.create_string_buffer(cap)
    n = self._L.erthost_encode(self._h, cands.ctypes.data, len(cands), 1 if fmt == "csv" else 0, unix_seconds, nanos, offset, length,
                               1 if sample_file_is_devnull else 0, buf, cap)
    if n < 0 or n >= cap:
        raise RuntimeError("erthost_encode failed" if n < 0 else "encode buffer too small")
    return buf.value.decode()


Parsers.parse_filtered = _parse_filtered
Parsers.encode = _encode


class Receiver:
    """NewDecoder + RegisterProtocol(NewParser(..)) + Allocate, then Decode (main.go:59-86,235)."""

    CFG = ("DataRate", "BlockSize", "BlockSize2", "ChipLength", "SymbolLength", "SampleRate", "PreambleSymbols",
           "PacketSymbols", "PreambleLength", "PacketLength", "BufferLength", "CenterFreq")

    def __init__(self, msgtypes: str, chip_length: int = 72, device: int = 0, max_blocks: int = 0, max_cands: int = 0):
        self._L = lib()
        err = C.create_string_buffer(512)
        self._h = self._L.erthost_new(msgtypes.encode(), chip_length, device, max_blocks, max_cands, err, 512)
        if not self._h:
            raise RuntimeError(err.value.decode())
        v = (C.c_int32 * 12)()
        self._L.erthost_config(self._h, v)
        self.cfg = dict(zip(self.CFG, [int(x) for x in v]))
        self.cfg["CenterFreq"] &= 0xFFFFFFFF

    def close(self):
        if self._h:
            self._L.erthost_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        if self._L.erthost_reset(self._h) != 0:
            raise RuntimeError(self._L.erthost_error(self._h).decode())

    def log(self) -> str:
        b = C.create_string_buffer(2048)
        self._L.erthost_log(self._h, b, 2048)
        return b.value.decode()

    def decode(self, iq, cap: int = 4096) -> list[HostMessage]:
        iq = np.ascontiguousarray(iq, dtype=np.uint8)
        out = (_Msg * cap)()
        n = self._L.erthost_decode(self._h, iq.ctypes.data, iq.size, out, cap)
        if n == -2:
            raise ValueError(self._L.erthost_error(self._h).decode())
        if n < 0:
            raise RuntimeError(self._L.erthost_error(self._h).decode())
        if n > cap:
            raise OverflowError(f"{n} messages, cap {cap}")
        return _messages(out, n)
