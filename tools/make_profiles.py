"""Turn the ncu artefacts a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.

Inputs (produced on the GPU box):
  (TAG = r1, r2, ...: `python tools/make_profiles.py TAG`)
  gpurun_out/bench_TAG.json        python bench.py --steps 30 --warmup 3
  gpurun_out/bench_TAG_ref.json    python bench.py --impl reference --steps 3 --warmup 1
  gpurun_out/launches_TAG.csv      ncu --metrics gpu__time_duration.sum --clock-control none --csv ... bench.py --steps 2 --warmup 1
  gpurun_out/prof_TAG_final.ncu-rep  ncu --set full --clock-control none --import-source on -k regex:... bench.py --steps 2 --warmup 1
"""
import collections, csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

# ---- launch list
rows = [r for r in csv.reader(l for l in open(os.path.join(G, f"launches_{tag}.csv")) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(list)
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").strip()
    v = float(r[vi].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r[ui], 1)
    agg[name].append(v)
ours = {k: v for k, v in agg.items() if not any(x in k for x in ("synth", "elementwise", "at::"))}
tot = sum(sum(v) for v in ours.values())
big = {k: [x for x in v if x > 0.5 * max(v)] for k, v in ours.items()}   # the 1 GiB launches only
tot_big = sum(sum(v) / len(v) for v in big.values())
tbl = ["| kernel | launches (all sizes) | share of all launches | mean of the 1 GiB launches | share of a 1 GiB step |", "|---|---|---|---|---|"]
for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
    b = big[k]
    tbl.append(f"| `{k[:60]}` | {len(v)} | {100*sum(v)/tot:.1f} % | {sum(b)/len(b)/1000:.1f} us | {100*(sum(b)/len(b))/tot_big:.1f} % |")
tbl = "\n".join(tbl)

# ---- full metrics
raw = subprocess.run(["ncu", "-i", os.path.join(G, f"prof_{tag}_final.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"), ("smsp__inst_executed.sum", "warp_inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1_data_pipe_pct"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pct"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pct"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
        ("sm__cycles_elapsed.avg.per_second", "sm_ghz")]
kern = []
for r in rr[2:]:
    d = {}
    for name, key in want:
        if name in h:
            i = h.index(name)
            d[key] = r[i] + ((" " + units[i]) if units[i] and key in ("time", "dram_read", "dram_write") else "")
    kern.append(d)
json.dump(kern, open(os.path.join(P, f"{tag}_kernels.json"), "w"), indent=1)
det = subprocess.run(["ncu", "-i", os.path.join(G, f"prof_{tag}_final.ncu-rep"), "--page", "details"], capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_ncu_details_demod_fast.txt"), "w").write("\n".join(det.splitlines()[:400]) + "\n")

def to_bytes(s):
    v, u = s.split()
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

d = kern[0]
rd, wr, ns = to_bytes(d["dram_read"]), to_bytes(d["dram_write"]), 536870912
json.dump({"kernel": re.sub(r"\(.*", "", d["kernel"]).replace("void ert::", "").strip(), "source": f"ncu --set full, profiles/{tag}_kernels.json (1 GiB launch)",
           "dram_bytes_read": rd, "dram_bytes_write": wr, "samples_per_launch": ns, "dram_bytes_per_sample": (rd + wr) / ns},
          open(os.path.join(P, "demod_traffic.json"), "w"), indent=1)
for f in (f"bench_{tag}.json", f"bench_{tag}_ref.json"):
    src = os.path.join(G, f)
    if os.path.exists(src):
        line = [l for l in open(src) if l.startswith("{")][-1]
        open(os.path.join(P, f.replace(f"bench_{tag}", f"{tag}_bench_line")), "w").write(line)
open(os.path.join(P, f"{tag}_launches.csv"), "w").write(open(os.path.join(G, f"launches_{tag}.csv")).read())
open(os.path.join(P, f"{tag}_launch_table.md"), "w").write(tbl + "\n")
print(tbl)
for k in kern:
    print(k["kernel"][:45], k["time"], k["dram_read"], k["dram_write"], k["issue_active_pct"], k["l1_data_pipe_pct"], k["regs"])
