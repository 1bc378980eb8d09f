"""Per CUDA source line: stall samples and warp instructions of an ncu report (needs -lineinfo and
--import-source on).  Usage: python profiles/ncu_lines.py <report.ncu-rep> [units] [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file, hdr, recs = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 10 and r[0] == "Line No":
        hdr = {k: i for i, k in enumerate(r)}
        stall_cols = [(k, i) for i, k in enumerate(r) if k.startswith("stall_")]
    elif len(r) > 10 and hdr and r[2] == "-":          # a CUDA source line (aggregated over its SASS)
        try:
            smp = float(r[hdr["# Samples"]] or 0)
            ins = float(r[hdr["Instructions Executed"]] or 0)
        except ValueError:
            continue
        st = sorted(((float(r[i] or 0), k[6:]) for k, i in stall_cols if r[i] not in ("", "0")), reverse=True)[:3]
        recs.append((smp, ins, cur_file, r[0], r[1].strip()[:90], st))
tot_s = sum(r[0] for r in recs)
tot_i = sum(r[1] for r in recs)
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.4g} = {tot_i / units:.1f} per unit")
for smp, ins, f, ln, src, st in sorted(recs, reverse=True)[:top]:
    print(f"{100 * smp / tot_s:5.1f}% smp {ins / units:8.1f} inst  {f}:{ln:>4s}  {src}   [{', '.join(f'{k} {v:.0f}' for v, k in st)}]")

# optional phase buckets: YH_BUCKETS="name:lo-hi,name:lo-hi" over lines of the main .cu file
import os
spec = os.environ.get("YH_BUCKETS")
if spec:
    main = os.environ.get("YH_BUCKET_FILE", "yh_decode_nms.cu")
    print("phase buckets (share of samples, warp instructions per unit):")
    rest_s, rest_i = tot_s, tot_i
    for item in spec.split(","):
        name, rng = item.split(":")
        lo, hi = (int(x) for x in rng.split("-"))
        ss = sum(r[0] for r in recs if r[2] == main and lo <= int(r[3]) <= hi)
        ii = sum(r[1] for r in recs if r[2] == main and lo <= int(r[3]) <= hi)
        rest_s -= ss
        rest_i -= ii
        print(f"  {name:28s} {100 * ss / tot_s:5.1f}%  {ii / units:8.1f}")
    print(f"  {'(other files / lines)':28s} {100 * rest_s / tot_s:5.1f}%  {rest_i / units:8.1f}")
