"""Where a kernel's warp-samples and instructions go, from the SASS page of an `ncu --set full --import-source on` report:
instructions are clustered by their execution count (a straight-line region executes every instruction the same number of
times), each cluster is printed with its share of samples / instructions, mean active threads and dominant stall reasons.

    python tools/ncu_source_breakdown.py gpurun_out/r02_final_200k.ncu-rep regex:shadow_light [launch-skip]
"""
import collections
import csv
import io
import subprocess
import sys

rep, kernel = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1] if rows and len(rows[0]) > 1 else "?")
h = rows[1]
col = {name: i for i, name in enumerate(h)}
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = []
for r in rows[2:]:
    try:
        data.append((r[col["Source"]].strip(), int(r[col["Warp Stall Sampling (All Samples)"]]), int(r[col["Instructions Executed"]]),
                     float(r[col["Avg. Threads Executed"]]), [int(r[i]) for i in stall_cols]))
    except (ValueError, IndexError):
        pass
half = len(data) // 2
if half and data[0][0] == data[half][0]:  # the page lists the function twice
    data = data[:half]
tot_s, tot_i = sum(d[1] for d in data), sum(d[2] for d in data)
print(f"{len(data)} SASS instructions, {tot_s} warp samples, {tot_i} warp-instructions executed")
groups = collections.OrderedDict()
for n, (src, ss, ex, av, st) in enumerate(data):
    g = groups.setdefault(ex, dict(n=0, s=0, i=0, av=0.0, lo=n, hi=n, st=[0] * len(stall_cols), ops=collections.Counter()))
    g["n"] += 1; g["s"] += ss; g["i"] += ex; g["av"] += av; g["hi"] = n
    g["st"] = [a + b for a, b in zip(g["st"], st)]
    t = src.split()
    g["ops"][(t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]] += 1
print(f"{'executions':>12} {'instr':>6} {'samples':>8} {'instr share':>11} {'avg thr':>7}  rows        top opcodes / top stalls")
for ex, g in sorted(groups.items(), key=lambda kv: -kv[1]["s"])[:18]:
    stalls = sorted(zip(g["st"], [h[i] for i in stall_cols]), reverse=True)[:3]
    print(f"{ex:>12} {g['n']:>6} {100 * g['s'] / tot_s:>7.2f}% {100 * g['i'] / tot_i:>10.2f}% {g['av'] / g['n']:>7.1f}  {g['lo']:>5}-{g['hi']:<5} "
          f"{' '.join(f'{o}x{c}' for o, c in g['ops'].most_common(4))} | {' '.join(f'{n[6:]}={100 * v / max(1, g['s']):.0f}%' for v, n in stalls if v)}")
