"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/ertgpu.h declares; argument validation that needs no GPU behaves as documented."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from rtlamr_b200 import capi, synth


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "ertgpu.h")).read()
    declared = set(re.findall(r"\b(ertgpu_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ertgpu_handle"}
    L = capi.lib()
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in ertgpu.h but not exported"
    assert set(capi.EXPORTS) == declared
    assert L.ertgpu_abi_version() == 2


def test_struct_layouts_match_header(built):
    assert C.sizeof(capi.Candidate) == 160
    assert C.sizeof(capi.SynthPacket) == 216
    assert capi.Candidate.bytes.offset == 24 and capi.Candidate.r900_digits.offset == 116


def test_stock_protocols_match_reference_parsers(built):
    want = {  # scm/scm.go:42-50, scmplus/scmplus.go:49-57, idm/idm.go:48-56, netidm/netidm.go:60-68, r900/r900.go:57-65
        "scm": ("111110010101001100000", 21, 96, 912600155),
        "scm+": ("0001011010100011", 16, 128, 912600155),
        "idm": ("01010101010101010001011010100011", 32, 736, 912600155),
        "netidm": ("01010101010101010001011010100011", 32, 736, 912600155),
        "r900": ("00000000000000001110010101100100", 32, 116, 912380000),
        "r900bcd": ("00000000000000001110010101100100", 32, 116, 912380000),
    }
    for name, (pre, ps, pk, cf) in want.items():
        p = capi.stock_protocol(name, 72)
        assert p.preamble.decode() == pre and p.preamble_symbols == ps and p.packet_symbols == pk
        assert p.center_freq == cf and p.data_rate == 32768 and p.chip_length == 72
        assert len(pre) == ps
    with pytest.raises(capi.ErtGpuError):  # parse.go:49 "invalid message type"
        capi.stock_protocol("bogus", 72)


def test_call_order_and_argument_errors_without_gpu(built):
    h = capi.Handle()
    with pytest.raises(capi.ErtGpuError) as e:
        h.allocate(0)  # Allocate before any RegisterProtocol
    assert e.value.code == capi.EINVAL
    bad = capi.stock_protocol("scm", 72)
    bad.preamble = b"10x1"
    with pytest.raises(capi.ErtGpuError):
        h.register(bad)
    with pytest.raises(capi.ErtGpuError):  # decode before allocate
        h.decode(np.zeros(8192, dtype=np.uint8))
    h.close()


def test_no_cpu_fallback_without_device(built):
    """On a machine without a CUDA device allocate must fail loudly (ERTGPU_ECUDA), never fall back."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("CUDA device present")
    h = capi.Handle()
    h.register(capi.stock_protocol("scm", 72))
    with pytest.raises(capi.ErtGpuError) as e:
        h.allocate(0)
    assert e.value.code == capi.ECUDA
    h.close()


def test_host_generator_is_deterministic_and_shardable(built):
    pk, truth = synth.make_packets("scm,r900", 72, 1 << 20, seed=5, spacing=1 << 18)
    whole = synth.host_fill(0, 1 << 20, 0x5EED0002, pk)
    parts = [synth.host_fill(a, b - a, 0x5EED0002, pk) for a, b in ((0, 300001), (300001, 777777), (777777, 1 << 20))]
    assert np.array_equal(whole, np.concatenate(parts))
    assert 126.5 < whole.mean() < 128.5
    assert len(truth) == 4


def test_go_shim_binds_only_declared_symbols():
    """The cgo shim cannot be compiled here (no Go toolchain); at least every C.ertgpu_* / C.ERTGPU_* name it
    uses must exist in include/ertgpu.h, and it must keep decode.go's exported surface."""
    hdr = open(os.path.join(ROOT, "include", "ertgpu.h")).read()
    go = open(os.path.join(ROOT, "go", "protocol", "decode_cuda.go")).read()
    used = set(re.findall(r"\bC\.((?:ertgpu|ERTGPU)_[A-Za-z0-9_]+)", go))
    assert len(used) >= 12
    for name in sorted(used):
        assert re.search(r"\b" + re.escape(name) + r"\b", hdr), f"{name} used by the Go shim but not in ertgpu.h"
    for exported in ("type PacketConfig struct", "type Decoder struct", "func NewDecoder() Decoder",
                     "func (d *Decoder) RegisterProtocol(p Parser)", "func (d *Decoder) Allocate()",
                     "Decode(input []byte) chan Message", "func (d Decoder) Log()", "func NextPowerOf2(v int) int",
                     "type Demodulator interface", "func NewMagLUT()"):
        assert exported in go, exported
    assert go.startswith("//go:build cuda")
