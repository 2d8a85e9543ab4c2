"""Warp-stall evidence for one kernel of an ncu report (needs `--set full --import-source on`, built with -lineinfo).

  python tools/ncu_stalls.py REPORT.ncu-rep KERNEL_REGEX [out.md]

Reads `ncu --page source --csv` of the LAST profiled launch that matches, and writes: the sampled stall reasons of the
whole kernel, of the steady-state loop (instructions executed >= 90 % of the hottest instruction) and of everything
else, the same per opcode class inside the loop, the address ranges that hold the samples, and the hottest 25 SASS lines."""
import collections, csv, subprocess, sys

rep, pat = sys.argv[1], sys.argv[2]
out = open(sys.argv[3], "w") if len(sys.argv) > 3 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[1]
launches, cur, names = [], None, []
for r in rows:
    if len(r) >= 1 and r[0] == "Kernel Name":
        cur = []
        launches.append(cur)
        names.append(r[1] if len(r) > 1 else "")
        continue
    if r == h or len(r) != len(h):
        continue
    cur.append(r)
rows, name = launches[-1], names[-1]
ix = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
S = lambda r: int(r[ix["# Samples"]])
E = lambda r: int(r[ix["Instructions Executed"]])
tot = sum(S(r) for r in rows)
mx = max(E(r) for r in rows)
hot = [r for r in rows if E(r) >= 0.9 * mx]


def table(rs):
    c = collections.Counter()
    for r in rs:
        for s in stalls:
            c[s.replace("stall_", "")] += int(r[ix[s]])
    n = sum(c.values()) or 1
    return ", ".join(f"{k} {100 * v / n:.1f} %" for k, v in c.most_common() if v >= 0.01 * n)


def opclass(r):
    sp = r[ix["Source"]].split()
    op = (sp[1] if sp[0].startswith("@") else sp[0]).rstrip(";")
    return ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STG", "SHF")) else op.split(".")[0]


w = out.write
w(f"# Warp-stall sampling: `{name[:110]}`\n\n")
w(f"Source: `{rep}` (last matching launch), {tot} samples over {len(rows)} SASS instructions; hottest instruction executed {mx} times per warp-set.\n\n")
w(f"* whole kernel: {table(rows)}\n")
hs = sum(S(r) for r in hot)
w(f"* steady-state loop ({len(hot)} instructions, {100 * hs / tot:.1f} % of the samples): {table(hot)}\n")
cold = [r for r in rows if E(r) < 0.9 * mx]
w(f"* everything else ({100 * (tot - hs) / tot:.1f} % of the samples: lead-in bodies, barrier waits, prologue, word flush): {table(cold)}\n\n")
w("## Inside the loop, per opcode class\n\n| opcode | instructions | samples | share of loop | top stall reasons |\n|---|---|---|---|---|\n")
agg, cnt = collections.Counter(), collections.Counter()
by = collections.defaultdict(list)
for r in hot:
    k = opclass(r)
    agg[k] += S(r)
    cnt[k] += 1
    by[k].append(r)
for k, v in agg.most_common(12):
    w(f"| {k} | {cnt[k]} | {v} | {100 * v / hs:.1f} % | {table(by[k])} |\n")
w("\n## Where the samples are (runs of instructions with the same execution count)\n\n| instructions | executed | samples | share | first instruction |\n|---|---|---|---|---|\n")
reg = []
for i, r in enumerate(rows):
    e, s = E(r), S(r)
    if reg and abs(reg[-1][2] - e) <= 0.02 * max(e, 1):
        reg[-1][1] = i
        reg[-1][3] += s
    else:
        reg.append([i, i, e, s])
for a, b, e, s in reg:
    if s >= 0.004 * tot:
        w(f"| {a}-{b} | {e} | {s} | {100 * s / tot:.1f} % | `{rows[a][ix['Source']].strip()[:60]}` |\n")
w("\n## Hottest SASS lines\n\n| samples | executed | instruction | stalls |\n|---|---|---|---|\n")
for r in sorted(rows, key=S, reverse=True)[:25]:
    st = {k.replace("stall_", ""): int(r[ix[k]]) for k in stalls if int(r[ix[k]])}
    st = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
    w(f"| {S(r)} | {E(r)} | `{r[ix['Source']].strip()[:70]}` | {st} |\n")
