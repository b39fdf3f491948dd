"""Summarise an .ncu-rep (read here, no GPU): key throughput metrics + stall reasons per kernel."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sm__cycles_active.avg.pct_of_peak_sustained_elapsed" if False else "sm__cycles_active.avg"]
stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
for r in rows[2:]:
    print("=" * 100)
    print(r[idx["Kernel Name"]])
    for w in want:
        if w in idx:
            print(f"  {w:78s} {r[idx[w]]:>18s} {units[idx[w]]}")
    vals = sorted([(float(r[idx[h]].replace(",", "")), h) for h in stall if r[idx[h]] not in ("", "n/a")], reverse=True)[:9]
    print("  stall reasons (warps per issue-active cycle):")
    for v, h in vals:
        print(f"    {v:7.3f} {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')}")
