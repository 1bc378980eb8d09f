#!/bin/bash
# Rebuilds the text summaries under profiles/ from the reports gpurun brought back in gpurun_out/ (no GPU needed).
set -e
cd "$(dirname "$0")/.."
H=keras-object-detection_b200/csrc/yh_decode_nms_impl.cuh
ln() { grep -n "$1" "$H" | head -1 | cut -d: -f1; }
cp gpurun_out/launches_r1_final.csv profiles/launches_r1_final.csv
cp gpurun_out/bench_r1_final.json profiles/bench_r1_final.json
cp gpurun_out/bench_r1_ref.json profiles/bench_r1_reference_arm.json
python profiles/launch_summary.py gpurun_out/launches_r1_final.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 1" > profiles/launches_r1_final_summary.txt
A=$(ln "A': compaction of the survivors"); B=$(ln "B: stable descending rank (utils.py:98): r ="); T=$(ln "a duplicate rank <=> equal confidences"); K=$(ln "class key: the class id itself")
C=$(ln "C: same-class masks.  Slot t"); D=$(ln "D: suppression bits against same-class predecessors"); E=$(ln "E: greedy keep flags, fixed point of keep\[q\] = !any(supp\[q\] & keep) ----"); F=$(ln "F: output slot of every rank position")
DC=$(ln "^// Phase A: decode one cell"); DK=$(ln "^// Direct kernel: one warp per image"); TK=$(ln "decode_nms_tma_kernel(const E"); TE=$(ln "^// Cooperative kernel for big images")
IO=$(ln "^__device__ __forceinline__ bool suppresses")
(echo "# decode_nms_tma_kernel<2,20,2,float>, 1M dense VOC images (ncu --set full --clock-control none --import-source on), final kernel of round 1"
 python profiles/ncu_summarize.py gpurun_out/prof_r1_dense_final.ncu-rep 1000000 | sed -n 1,26p; echo
 YH_BUCKET_FILE=yh_decode_nms_impl.cuh YH_BUCKETS="iou_test:$IO-$((IO+20)),gt_bits:45-56,A_compact:$A-$((B-1)),B_rank:$B-$((T-1)),B_ties:$T-$((K-1)),keys_scatter:$K-$((C-1)),C_masks:$C-$((D-1)),D_preds:$D-$((E-1)),E_resolve:$E-$((F-1)),F_output:$F-$((DC-30)),decode_cell:$DC-$((DK-1)),tma_kernel:$TK-$((TE-1))" python profiles/ncu_lines.py gpurun_out/prof_r1_dense_final.ncu-rep 1000000 25) > profiles/ncu_r1_decode_nms_dense_final_summary.txt
PR=$(ln "// ---- producer$" ); PR=$(grep -n "// ---- producer" "$H" | tail -1 | cut -d: -f1)
CA=$(ln "A: decode own cell from the ring"); CA2=$(ln "A': compaction (utils.py:95, strict >)$"); CB=$(ln "B: stable descending rank (utils.py:98)$"); CC=$(ln "C: scatter to rank order (utils.py:24-32,40); class masks")
CD=$(ln "D: suppression words against same-class predecessors"); CE=$(ln "E: greedy keep flags, fixed point of keep\[q\] = !any(supp\[q\] & keep)$"); CF=$(ln "F: output slots (kws holds the final keep words)"); CEND=$(ln "^// host side")
(echo "# decode_nms_coop_kernel<80,3,float>, cfg5 data, 32,768 images (ncu --set full --clock-control none --import-source on), final kernel of round 1"
 python profiles/ncu_summarize.py gpurun_out/prof_r1_stress5_coop.ncu-rep 32768 | sed -n 1,24p; echo
 YH_BUCKET_FILE=yh_decode_nms_impl.cuh YH_BUCKETS="iou_test:$IO-$((IO+20)),gt_bits:45-56,producer:$PR-$((CA-25)),A_decode_wait:$CA-$((CA2-1)),A_compact:$CA2-$((CB-1)),B_rank:$CB-$((CC-1)),C_scatter_masks:$CC-$((CD-1)),D_preds:$CD-$((CE-1)),E_resolve:$CE-$((CF-1)),F_output:$CF-$((CEND-1)),decode_cell:$DC-$((DK-1))" python profiles/ncu_lines.py gpurun_out/prof_r1_stress5_coop.ncu-rep 32768 28) > profiles/ncu_r1_decode_nms_coop_cfg5_summary.txt
(echo "# loss_kernel<true>, batch 4096 (ncu --set full --clock-control none --import-source on), final kernel of round 1"
 python profiles/ncu_summarize.py gpurun_out/prof_r1_loss2.ncu-rep 4096 | sed -n 1,24p; echo
 python profiles/ncu_lines.py gpurun_out/prof_r1_loss2.ncu-rep 4096 20) > profiles/ncu_r1_loss_summary.txt
python - <<'PY'
import json,subprocess,csv,io
raw=subprocess.run(["ncu","-i","gpurun_out/prof_r1_dense_final.ncu-rep","--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw))); h,u,v=rows[0],rows[1],rows[2]
def g(k):
    i=h.index(k); x=float(v[i].replace(",","")); return x*{"Gbyte":1e9,"Mbyte":1e6,"Kbyte":1e3,"byte":1}[u[i]]
rd,wr=g("dram__bytes_read.sum"),g("dram__bytes_write.sum")
json.dump({"decode_nms_tma_kernel_dram_bytes_per_launch":rd+wr,
 "source":f"ncu --set full --clock-control none, profiles/ncu_r1_decode_nms_dense_final_summary.txt (1M dense VOC images, one launch): dram__bytes_read.sum {rd/1e9:.6f} GB + dram__bytes_write.sum {wr/1e9:.6f} GB",
 "algorithmic_bytes_per_launch":6843757120},open("profiles/traffic.json","w"),indent=1)
print("traffic", rd+wr)
PY
