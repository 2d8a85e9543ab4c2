"""ctypes binding of the libertgpu C ABI (include/ertgpu.h).

This is plumbing for tests/ and bench.py: the product is libertgpu.so itself, bound from Go by
the cgo shim in go/ (INTEGRATION.md).  There is no fallback of any kind: if the shared library
is missing or CUDA is unusable every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libertgpu.so")

OK, EINVAL, ECUDA, ENOMEM, ECAPACITY, ESIZE = 0, -1, -2, -3, -4, -5
MAX_PROTOCOLS, MAX_PREAMBLE, MAX_PACKET_BYTES, R900_DIGITS = 8, 32, 92, 42
CHECK_NONE, CHECK_CRC16, CHECK_IDM, CHECK_R900 = 0, 1, 2, 3
DECODE_ONLY_VALID = 1
CAND_HAS_R900 = 1
TAP_SIGNAL, TAP_CSUM, TAP_QUANTIZED, TAP_PACKED, TAP_R900_QUANTIZED = range(5)


class Protocol(C.Structure):
    _fields_ = [("name", C.c_char * 16), ("preamble", C.c_char * (MAX_PREAMBLE + 1)),
                ("data_rate", C.c_int32), ("chip_length", C.c_int32),
                ("preamble_symbols", C.c_int32), ("packet_symbols", C.c_int32),
                ("center_freq", C.c_uint32), ("check_kind", C.c_int32),
                ("crc_init", C.c_uint16), ("crc_poly", C.c_uint16), ("crc_residue", C.c_uint16),
                ("reserved", C.c_uint16), ("crc_from", C.c_int32), ("crc_to", C.c_int32)]


class DecoderConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "data_rate", "block_size", "block_size2", "chip_length", "symbol_length", "sample_rate",
        "preamble_symbols", "packet_symbols", "preamble_length", "packet_length",
        "buffer_length")] + [("center_freq", C.c_uint32), ("n_protocols", C.c_int32),
                             ("n_preambles", C.c_int32), ("packet_bytes", C.c_int32)]


class Candidate(C.Structure):
    _fields_ = [("block", C.c_int64), ("idx", C.c_int32), ("preamble_id", C.c_int32),
                ("check_mask", C.c_uint32), ("flags", C.c_uint32),
                ("bytes", C.c_uint8 * MAX_PACKET_BYTES), ("r900_digits", C.c_uint8 * R900_DIGITS),
                ("pad", C.c_uint8 * 2)]


class Shard(C.Structure):
    _fields_ = [("first_block", C.c_int64), ("last_block", C.c_int64), ("first_fed_block", C.c_int64)]


class SynthPacket(C.Structure):
    _fields_ = [("start_sample", C.c_int64), ("n_chips", C.c_int32), ("chip_length", C.c_int32),
                ("amp_i", C.c_int16), ("amp_q", C.c_int16), ("chips", C.c_uint8 * 192),
                ("pad", C.c_int32)]


CAND_DTYPE = np.dtype([("block", "<i8"), ("idx", "<i4"), ("preamble_id", "<i4"),
                       ("check_mask", "<u4"), ("flags", "<u4"),
                       ("bytes", "u1", (MAX_PACKET_BYTES,)), ("r900_digits", "u1", (R900_DIGITS,)),
                       ("pad", "u1", (2,))])
assert CAND_DTYPE.itemsize == C.sizeof(Candidate) == 160

# every symbol include/ertgpu.h declares
EXPORTS = [
    "ertgpu_abi_version", "ertgpu_last_error", "ertgpu_create", "ertgpu_destroy",
    "ertgpu_register_protocol", "ertgpu_stock_protocol", "ertgpu_allocate", "ertgpu_get_config",
    "ertgpu_reset", "ertgpu_decode", "ertgpu_decode_device_async", "ertgpu_fetch",
    "ertgpu_last_counts", "ertgpu_last_launches", "ertgpu_set_stage_timing",
    "ertgpu_last_stage_ms", "ertgpu_stage_ms_mean", "ertgpu_tap", "ertgpu_set_demod_variant",
    "ertgpu_host_alloc", "ertgpu_host_free", "ertgpu_synth_fill",
    "ertgpu_bind_host_thread", "ertgpu_last_kernels", "ertgpu_plan_shards", "ertgpu_decode_sharded",
]

_lib = None


class ErtGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libertgpu error {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    """Load libertgpu.so (built in-tree by __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_size_t
    L.ertgpu_abi_version.restype = C.c_int
    L.ertgpu_last_error.restype = C.c_char_p
    L.ertgpu_last_error.argtypes = [vp]
    L.ertgpu_create.argtypes = [C.POINTER(vp)]
    L.ertgpu_destroy.restype = None
    L.ertgpu_destroy.argtypes = [vp]
    L.ertgpu_register_protocol.argtypes = [vp, C.POINTER(Protocol)]
    L.ertgpu_stock_protocol.argtypes = [C.c_char_p, i32, C.POINTER(Protocol)]
    L.ertgpu_allocate.argtypes = [vp, i32, i64, i64]
    L.ertgpu_get_config.argtypes = [vp, C.POINTER(DecoderConfig)]
    L.ertgpu_reset.argtypes = [vp]
    L.ertgpu_decode.argtypes = [vp, vp, sz, u32, vp, sz, C.POINTER(sz)]
    L.ertgpu_decode_device_async.argtypes = [vp, vp, sz, u32, vp]
    L.ertgpu_fetch.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.ertgpu_last_counts.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.ertgpu_last_launches.restype = i64
    L.ertgpu_last_launches.argtypes = [vp]
    L.ertgpu_tap.argtypes = [vp, i32, i64, vp, sz, C.POINTER(sz)]
    L.ertgpu_set_stage_timing.argtypes = [vp, i32]
    L.ertgpu_last_stage_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.ertgpu_stage_ms_mean.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i64)]
    L.ertgpu_set_demod_variant.argtypes = [vp, i32]
    L.ertgpu_host_alloc.argtypes = [C.POINTER(vp), sz]
    L.ertgpu_host_free.argtypes = [vp]
    L.ertgpu_synth_fill.argtypes = [i32, vp, i64, i64, u64, vp, i64, vp]
    L.ertgpu_bind_host_thread.argtypes = [i32, C.POINTER(i32), C.POINTER(i32)]
    L.ertgpu_last_kernels.restype = C.c_char_p
    L.ertgpu_last_kernels.argtypes = [vp]
    L.ertgpu_plan_shards.argtypes = [i64, i32, i32, i32, C.POINTER(Shard)]
    L.ertgpu_decode_sharded.argtypes = [C.POINTER(vp), i32, vp, sz, u32, vp, sz, C.POINTER(sz)]
    for name in EXPORTS:
        getattr(L, name)  # AttributeError if the library does not export it
    _lib = L
    return L


def stock_protocol(msgtype: str, chip_length: int) -> Protocol:
    p = Protocol()
    rc = lib().ertgpu_stock_protocol(msgtype.encode(), chip_length, C.byref(p))
    if rc != OK:
        raise ErtGpuError(rc, f"invalid message type: {msgtype!r}")
    return p


class Handle:
    """One ertgpu_handle == one protocol.Decoder (reference protocol/decode.go:45-63)."""

    def __init__(self):
        self._L = lib()
        self._h = C.c_void_p()
        self._check(self._L.ertgpu_create(C.byref(self._h)))
        self.cfg = None

    def _check(self, rc: int):
        if rc != OK:
            raise ErtGpuError(rc, (self._L.ertgpu_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            self._L.ertgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def register(self, proto: Protocol):
        self._check(self._L.ertgpu_register_protocol(self._h, C.byref(proto)))

    def allocate(self, device: int = 0, max_blocks_per_call: int = 0, max_candidates: int = 0):
        self._check(self._L.ertgpu_allocate(self._h, device, max_blocks_per_call, max_candidates))
        cfg = DecoderConfig()
        self._check(self._L.ertgpu_get_config(self._h, C.byref(cfg)))
        self.cfg = cfg
        return cfg

    def reset(self):
        self._check(self._L.ertgpu_reset(self._h))

    def set_demod_variant(self, variant: int):
        self._check(self._L.ertgpu_set_demod_variant(self._h, variant))

    def _buffer(self, cap):
        # reused between calls: allocating and zeroing tens of MB per call would dominate small decodes
        buf = getattr(self, "_outbuf", None)
        if buf is None or len(buf) < cap:
            buf = np.empty(cap, dtype=CAND_DTYPE)
            self._outbuf = buf
        return buf[:cap]

    def _deliver(self, call, cap):
        while True:
            out = self._buffer(cap)
            n = C.c_size_t(0)
            rc = call(out.ctypes.data, cap, C.byref(n))
            if rc == ECAPACITY and n.value > cap and "internal" not in (
                    self._L.ertgpu_last_error(self._h) or b"").decode():
                cap = n.value
                call = lambda p, c, nn: self._L.ertgpu_fetch(self._h, p, c, nn)  # noqa: E731
                continue
            self._check(rc)
            return out[:n.value].copy()

    def decode(self, iq, flags: int = 0, cap: int = 4096) -> np.ndarray:
        """ertgpu_decode on a HOST buffer (numpy uint8 or an address/size pair)."""
        if isinstance(iq, tuple):
            addr, nbytes = iq
        else:
            iq = np.ascontiguousarray(iq, dtype=np.uint8)
            addr, nbytes = iq.ctypes.data, iq.size
        return self._deliver(
            lambda p, c, nn: self._L.ertgpu_decode(self._h, addr, nbytes, flags, p, c, nn), cap)

    def decode_device_async(self, d_ptr: int, nbytes: int, flags: int = 0, stream: int = 0):
        self._check(self._L.ertgpu_decode_device_async(self._h, d_ptr, nbytes, flags, stream))

    def fetch(self, cap: int = 4096) -> np.ndarray:
        return self._deliver(lambda p, c, nn: self._L.ertgpu_fetch(self._h, p, c, nn), cap)

    def last_counts(self):
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self._L.ertgpu_last_counts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_launches(self) -> int:
        return int(self._L.ertgpu_last_launches(self._h))

    def set_stage_timing(self, enable: bool):
        self._check(self._L.ertgpu_set_stage_timing(self._h, 1 if enable else 0))

    def last_stage_ms(self):
        ms = (C.c_float * 4)()
        self._check(self._L.ertgpu_last_stage_ms(self._h, ms))
        return dict(zip(("demod", "search", "extract", "carry"), [float(x) for x in ms]))

    def stage_ms_mean(self):
        ms, n = (C.c_float * 4)(), C.c_int64(0)
        self._check(self._L.ertgpu_stage_ms_mean(self._h, ms, C.byref(n)))
        return dict(zip(("demod", "search", "extract", "carry"), [float(x) for x in ms])), n.value

    def demod_kernel_name(self) -> str:
        """The demod instantiation the last decode launched (first entry of ertgpu_last_kernels)."""
        return (self._L.ertgpu_last_kernels(self._h) or b"").decode().split(";")[0].strip()

    def last_kernels(self) -> str:
        return (self._L.ertgpu_last_kernels(self._h) or b"").decode()

    def tap(self, which: int, block: int) -> np.ndarray:
        n = C.c_size_t(0)
        self._check(self._L.ertgpu_tap(self._h, which, block, None, 0, C.byref(n)))
        buf = np.zeros(n.value, dtype=np.uint8)
        self._check(self._L.ertgpu_tap(self._h, which, block, buf.ctypes.data, buf.size, C.byref(n)))
        if which in (TAP_SIGNAL, TAP_CSUM):
            return buf.view(np.float32)
        return buf


def new_decoder(msgtypes, chip_length: int = 72, device: int = 0, max_blocks_per_call: int = 0,
                max_candidates: int = 0) -> Handle:
    """NewDecoder + RegisterProtocol(NewParser(name, chip_length)) for each name + Allocate
    (reference main.go:64-86)."""
    if isinstance(msgtypes, str):
        msgtypes = [m.strip() for m in msgtypes.split(",") if m.strip()]
    h = Handle()
    for m in msgtypes:
        h.register(stock_protocol(m, chip_length))
    h.allocate(device, max_blocks_per_call, max_candidates)
    return h


def bind_host_thread(device: int):
    """ertgpu_bind_host_thread: pin the calling thread to the CPUs next to `device`.  Returns a
    small dict for the bench line (None fields when the topology is unknown: nothing changed)."""
    n, node = C.c_int32(0), C.c_int32(-1)
    rc = lib().ertgpu_bind_host_thread(device, C.byref(n), C.byref(node))
    return {"bound": rc == OK, "cpus": n.value if rc == OK else None, "numa_node": node.value if node.value >= 0 else None}


def plan_shards(total_blocks: int, nshards: int, block_size: int, packet_length: int):
    arr = (Shard * nshards)()
    rc = lib().ertgpu_plan_shards(total_blocks, nshards, block_size, packet_length, arr)
    if rc != OK:
        raise ErtGpuError(rc, "ertgpu_plan_shards: bad request")
    return [(s.first_block, s.last_block, s.first_fed_block) for s in arr]


def decode_sharded(handles, iq, flags: int = 0, cap: int = 1 << 16) -> np.ndarray:
    """ertgpu_decode_sharded: one host buffer over several handles (one host thread per handle)."""
    if isinstance(iq, tuple):
        addr, nbytes = iq
    else:
        iq = np.ascontiguousarray(iq, dtype=np.uint8)
        addr, nbytes = iq.ctypes.data, iq.size
    hs = (C.c_void_p * len(handles))(*[h._h for h in handles])
    while True:
        out = np.empty(cap, dtype=CAND_DTYPE)
        n = C.c_size_t(0)
        rc = lib().ertgpu_decode_sharded(hs, len(handles), addr, nbytes, flags, out.ctypes.data, cap, C.byref(n))
        if rc == ECAPACITY and n.value > cap:
            cap = n.value
            continue
        if rc != OK:
            msgs = [(lib().ertgpu_last_error(h._h) or b"").decode() for h in handles]
            raise ErtGpuError(rc, "; ".join(m for m in msgs if m) or "ertgpu_decode_sharded failed")
        return out[:n.value].copy()


def synth_fill(device: int, d_ptr: int, first_sample: int, nsamples: int, seed: int,
               packets: np.ndarray | None, stream: int = 0):
    """ertgpu_synth_fill: packets is a numpy array of SYNTH_DTYPE sorted by start_sample."""
    if packets is None or len(packets) == 0:
        ptr, n = None, 0
    else:
        packets = np.ascontiguousarray(packets)
        ptr, n = packets.ctypes.data, len(packets)
    rc = lib().ertgpu_synth_fill(device, d_ptr, first_sample, nsamples, seed, ptr, n, stream)
    if rc != OK:
        raise ErtGpuError(rc, "ertgpu_synth_fill failed")


SYNTH_DTYPE = np.dtype([("start_sample", "<i8"), ("n_chips", "<i4"), ("chip_length", "<i4"),
                        ("amp_i", "<i2"), ("amp_q", "<i2"), ("chips", "u1", (192,)),
                        ("pad", "<i4")])
assert SYNTH_DTYPE.itemsize == C.sizeof(SynthPacket) == 216
