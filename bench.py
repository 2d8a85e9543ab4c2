#!/usr/bin/env python
"""bench.py -- headline benchmark of the protocol.Decoder hot path on B200.

Metric (BASELINE.json): IQ Msamples/s decoded (1 sample = I byte + Q byte) on synthetic uint8 IQ
with injected ERT packets, per GPU workload = BASELINE configs[1]: scm, ChipLength 72
(`-symbollength=72`), 1 GiB of IQ (536 870 912 samples, 131 072 reference blocks).

  python bench.py --gpus N --steps K --warmup W            our arm (libertgpu.so, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  CPU arm: the reference's algorithm on
                                                           the box's host cores (oracle port: Go
                                                           is not in this image)

For N > 1 the driver launches one rank per GPU with torch.distributed.run; the stream is cut into
N contiguous block-aligned shards (weak scaling: 1 GiB per GPU) with a leading halo, and there is no
collective on the data path; torch.distributed is used for the barrier and
the max-over-ranks of the device time only.

One JSON line is printed by rank 0.  `value` = whole-job Msamples/s with the IQ resident in HBM
(CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks);
`e2e` = the same through the C-ABI call `ertgpu_decode` with pinned HOST buffers (H2D inside the
timed region); `roofline` = the demod kernel against the measured HBM copy bandwidth;
`cpu_baseline` = the CPU restatement timed on this box (bounded sample).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MSGTYPE, CHIP_LENGTH = "scm", 72
BYTES_PER_GPU = 1 << 30
SEED = 0x5EED0002
PACKET_SPACING = 1 << 20
METRIC, UNIT = "IQ Msamples/s decoded (scm, ChipLength 72, synthetic uint8 IQ with injected ERT packets)", "Msamples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bytes-per-gpu", type=int, default=BYTES_PER_GPU)
    ap.add_argument("--cpu-sample-mib", type=int, default=64)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled every
    ~2 ms from a thread (the timed region of a default run is only a few ms long), with the
    `nvidia-smi -lms` recipe as fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.nvml, self.stop_flag, self.thread = None, False, None

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.gpu])
            except Exception:
                return self.gpu
        return self.gpu

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self._visible_index()}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append((time.time(), sm, mx, rs))
            except Exception:
                pass
            time.sleep(0.0005)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            n = self.nvml
            rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
            sm = sorted(r[1] for r in rows)
            bits = 0
            for r in rows:
                bits |= r[3]
            names = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            reasons = sorted(k for k, v in names.items() if bits & v)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": rows[0][2] if rows else None,
                    "reasons": reasons, "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.05] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


# ----------------------------------------------------------------------------- CPU arm
def cpu_rate(nthreads: int, sample_mib: int, repeats: int = 1):
    """Msamples/s of the CPU restatement (oracle/ert_oracle.c, Go-faithful Search, parsers included):
    `nthreads` independent decoders, each fed the same bounded sample of the workload in
    BlockSize2-byte Decode calls (ctypes releases the GIL, so the threads run in parallel)."""
    import numpy as np

    import oracle
    from rtlamr_b200 import synth

    nsamples = sample_mib << 19
    pk, _ = synth.make_packets(MSGTYPE, CHIP_LENGTH, nsamples, seed=1, spacing=PACKET_SPACING)
    iq = synth.host_fill(0, nsamples, SEED, pk)
    decs = [oracle.Oracle(MSGTYPE, CHIP_LENGTH, oracle.SEARCH_GO) for _ in range(nthreads)]
    bs2 = decs[0].cfg.block_size2
    iq = iq[: iq.size // bs2 * bs2]
    counts = [0] * nthreads

    def work(i):
        for _ in range(repeats):
            c, m = decs[i].decode(iq, cand_cap=1 << 18, msg_cap=1 << 14)
            counts[i] += len(m)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(nthreads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    total = nthreads * repeats * (iq.size // 2)
    return total / dt / 1e6, dt, counts[0], (f"first {sample_mib} MiB of the synthetic stream, decoded {repeats}x by each of "
                                             f"{nthreads} independent decoders")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle  # noqa: F401  (the reference arm is the one place bench.py executes oracle/)
    ncores = os.cpu_count() or 1
    nthreads = max(1, ncores)   # every host thread the box has
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_rate(nthreads, 8)
    vals, t_tot = [], 0.0
    steps = max(1, args.steps)
    for _ in range(steps):
        v, dt, _, sample = cpu_rate(nthreads, args.cpu_sample_mib)
        vals.append(v)
        t_tot += dt
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t_tot / steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"scm ChipLength=72, synthetic uint8 IQ with injected preambles; each step = "
                               f"{args.cpu_sample_mib} MiB sample x {nthreads} host threads"},
        "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": nthreads, "kind": "port",
                         "sample": f"{args.cpu_sample_mib} MiB of the stream per thread per step; C restatement of the Go "
                                   f"Decoder (oracle/ert_oracle.c) because no Go toolchain is present"},
        "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from rtlamr_b200 import capi, shard, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # geometry and this rank's shard of the global stream (weak scaling: bytes_per_gpu each)
    probe = capi.new_decoder(MSGTYPE, CHIP_LENGTH, device=local, max_blocks_per_call=1)
    cfg = probe.cfg
    bs, bs2, pkl = cfg.block_size, cfg.block_size2, cfg.packet_length
    probe.close()
    blocks_per_gpu = args.bytes_per_gpu // bs2
    plan = shard.plan(total_blocks=blocks_per_gpu * world, nranks=world, block_size=bs, packet_length=pkl)[rank]
    nblocks = plan.last_block - plan.first_fed_block          # halo + owned blocks
    nbytes = nblocks * bs2
    nsamples = nblocks * bs
    first_sample = plan.first_fed_block * bs

    h = capi.new_decoder(MSGTYPE, CHIP_LENGTH, device=local, max_blocks_per_call=nblocks, max_candidates=1 << 20)
    total_samples = blocks_per_gpu * world * bs
    pk, truth = synth.make_packets(MSGTYPE, CHIP_LENGTH, total_samples, seed=1, spacing=PACKET_SPACING)
    d_iq = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    capi.synth_fill(local, d_iq.data_ptr(), first_sample, nsamples, SEED, pk)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream

    def step_device(fetch=False):
        """One pass of the hot path over the resident batch: fresh stream, every kernel of the
        pipeline, then the per-step result read-back (candidate / valid counters, D2H).  The
        candidate records themselves are fetched when asked (correctness gate, e2e arm)."""
        h.reset()
        h.decode_device_async(d_iq.data_ptr(), nbytes, capi.DECODE_ONLY_VALID, st)
        counts = h.last_counts()
        return h.fetch(1 << 17) if fetch else counts

    # ---- correctness gate (untimed): every injected packet that lies in this rank's range decodes
    got = step_device(fetch=True)
    got = got[got["block"] >= (plan.first_block - plan.first_fed_block)]
    ids = {bytes(r["bytes"][:12]) for r in got}
    mine = [t for t in truth if plan.owns_start(t.start_sample + cfg.symbol_length, bs, cfg.buffer_length)
            and t.start_sample + cfg.buffer_length < total_samples]
    missing = [t for t in mine if t.data not in ids]
    if missing:
        raise SystemExit(f"rank {rank}: {len(missing)} of {len(mine)} injected packets not recovered")
    n_pkts_local = len({bytes(r["bytes"][:12]) for r in got})

    # ---- device-resident timing
    for _ in range(args.warmup):
        step_device()
    h.set_stage_timing(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.05)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    t_wall0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
        launches += h.last_launches()
    e1.record(stream)
    barrier()
    t_wall1 = time.time()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    stage_mean, stage_n = h.stage_ms_mean()   # CUDA events recorded by the library around each stage, every timed step
    assert stage_n == args.steps
    h.set_stage_timing(False)

    # ---- end to end through the C-ABI call with pinned host memory (H2D inside the timed region)
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.copy_(d_iq)
    torch.cuda.synchronize()
    hptr = host.data_ptr()
    d2h_bytes = 0

    def step_host():
        h.reset()
        return h.decode((hptr, nbytes), capi.DECODE_ONLY_VALID, 1 << 17)

    for _ in range(min(args.warmup, 2)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = step_host()
        d2h_bytes = r.nbytes + 24
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([ms_total, e2e_s * 1e3, float(n_pkts_local)], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, e2e_ms_total, n_pkts = float(tmax[0]), float(tmax[1]), int(tsum[2])
    else:
        e2e_ms_total, n_pkts = e2e_s * 1e3, n_pkts_local

    owned_samples_all = blocks_per_gpu * world * bs
    ms_per_step = ms_total / args.steps
    value = owned_samples_all / (ms_per_step * 1e-3) / 1e6
    e2e_value = owned_samples_all / (e2e_ms_total / args.steps * 1e-3) / 1e6

    if rank == 0:
        peaks = {}
        for p in (os.path.join(ROOT, "MEASURED_PEAKS.json"),):
            if os.path.exists(p):
                peaks = json.load(open(p))
        peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        dm = stage_mean["demod"]
        stages = {k: round(v, 4) for k, v in stage_mean.items()}
        achieved = 2.0 * nsamples / (dm * 1e-3) / 1e9   # 2 algorithmic bytes per sample
        traffic = None
        tp = os.path.join(ROOT, "profiles", "demod_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get("dram_bytes_per_sample", 0) * nsamples or None
            except Exception:
                traffic = None
        cpu = None
        if not args.skip_cpu_baseline and world == 1:
            v, dt, nmsg, sample = cpu_rate(1, args.cpu_sample_mib * 4, repeats=12)   # ~10-20 s of one core
            cpu = {"value": round(v, 1), "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": sample + f" ({dt:.1f} s; C restatement of the Go Decoder incl. Search and parsers, Go toolchain absent)"}
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"scm ChipLength=72 (SymbolLength 144, BlockSize {bs}), {args.bytes_per_gpu >> 20} MiB synthetic "
                                   f"uint8 IQ with injected SCM packets per GPU, device-resident",
                       "blocks_per_gpu": blocks_per_gpu, "halo_blocks": plan.first_block - plan.first_fed_block,
                       "l2": "input per step (1 GiB) exceeds the 126 MB L2: no flush needed",
                       "sharding": "contiguous block-aligned shards + halo, no collective on the data path"},
            "pkts_per_s": round(n_pkts / (ms_per_step * 1e-3), 1), "pkts_per_step": n_pkts,
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": nbytes * world, "d2h_bytes_per_step": int(d2h_bytes) * world,
                    "api": "ertgpu_decode (C ABI), pinned host input, chunked H2D overlapped with kernels"},
            "gpu_launches": int(launches),
            "stage_ms": stages,
            "roofline": {"bound": "hbm", "kernel": "demod_fast_kernel<72,W> (W = 7 or 8 resident warps, chosen per call by round count)", "achieved": round(achieved, 1), "peak": peak,
                         "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes": "2 B per IQ sample x samples per launch"},
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
