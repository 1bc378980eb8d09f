"""Summarise an ncu report: headline raw metrics + per-window SASS instruction counts and stall
reasons.  Usage: python profiles/ncu_summarize.py <report.ncu-rep> [images]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
images = float(sys.argv[2]) if len(sys.argv) > 2 else 1e6
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic"]
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"  {w:75s} {vals[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
recs = []
tot_stall = collections.Counter()
for r in rows[2:]:
    if len(r) < len(h):
        continue
    ie = float(r[ix["Instructions Executed"]] or 0)
    recs.append((r[ix["Source"]].strip(), ie, float(r[ix["Avg. Threads Executed"]] or 0), int(r[ix["# Samples"]] or 0)))
    for s_ in stalls:
        try:
            tot_stall[s_] += float(r[ix[s_]] or 0)
        except ValueError:
            pass
tot = sum(r[1] for r in recs)
print(f"warp instructions executed: {tot:.4g}  = {tot / images:.1f} per image")
T = sum(tot_stall.values())
print("stall samples:", ", ".join(f"{k[6:]} {100 * v / T:.1f}%" for k, v in tot_stall.most_common(8)))
win = 40
print("SASS windows (inst per image, avg active threads, samples, dominant opcodes):")
for s0 in range(0, len(recs), win):
    ch = recs[s0:s0 + win]
    c = sum(r[1] for r in ch) / images
    if c < 2:
        continue
    ops = collections.Counter()
    for r in ch:
        tok = r[0].split()
        op = tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "?")
        ops[op.split(".")[0]] += r[1] / images
    thr = sum(r[1] * r[2] for r in ch) / max(sum(r[1] for r in ch), 1e-9)
    print(f"  {s0:5d}-{s0 + win:5d} {c:7.1f} thr={thr:5.1f} smp={sum(r[3] for r in ch):6d}  " +
          " ".join(f"{k}:{v:.0f}" for k, v in ops.most_common(6)))
