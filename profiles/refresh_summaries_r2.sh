#!/bin/bash
# Round 2: rebuilds the text summaries under profiles/ from the reports gpurun brought back in gpurun_out/ (no GPU needed).
set -e
cd "$(dirname "$0")/.."
H=keras-object-detection_b200/csrc/yh_decode_nms_impl.cuh
REV=$(git rev-parse --short HEAD)
ln() { grep -n "$1" "$H" | head -1 | cut -d: -f1; }
cp gpurun_out/launches_r2_final.csv profiles/launches_r2_final.csv
cp gpurun_out/launches_r2_map.csv profiles/launches_r2_map_cfg4.csv
cp gpurun_out/bench_r2_final.json profiles/bench_r2_final.json
cp gpurun_out/bench_r2_ref.json profiles/bench_r2_reference_arm.json
python profiles/launch_summary.py gpurun_out/launches_r2_final.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 2 --warmup 1 (YH_BENCH_SUSTAINED=20)" > profiles/launches_r2_final_summary.txt
python profiles/launch_summary.py gpurun_out/launches_r2_map.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 200 python profiles/prof_map.py cfg4 (13 evaluator passes of 5,000 images: 2 launches each + 1 torch fill; fused update_state, reduce on its counting path)" > profiles/launches_r2_map_cfg4_summary.txt
A=$(ln "A': compaction of the survivors"); B=$(ln "B: stable descending rank (utils.py:98): r ="); T=$(ln "a duplicate rank <=> equal confidences"); K=$(ln "class key: the class id itself")
C=$(ln "C: same-class masks.  Slot t"); D=$(ln "D: suppression bits against same-class predecessors"); E=$(ln "E: greedy keep flags, fixed point of keep\[q\] = !any(supp\[q\] & keep) ----"); F=$(ln "F: output slot of every rank position")
DC=$(ln "^// Phase A: decode one cell"); DK=$(ln "^// Direct kernel: one warp per image"); TK=$(ln "decode_nms_tma_kernel(const E"); TE=$(ln "^// Cooperative kernel for big images")
IO=$(ln "^static __device__ __noinline__ unsigned suppresses_exact")
(echo "# decode_nms_tma_kernel<2,20,2,float>, 1M dense VOC images (ncu --set full --clock-control none --import-source on), round 2, commit $REV"
 python profiles/ncu_summarize.py gpurun_out/prof_r2_dense.ncu-rep 1000000 | sed -n 1,26p; echo
 YH_BUCKET_FILE=yh_decode_nms_impl.cuh YH_BUCKETS="iou_test:$IO-$((IO+32)),gt_bits:45-60,A_compact:$A-$((B-1)),B_rank:$B-$((T-1)),B_ties:$T-$((K-1)),keys_scatter:$K-$((C-1)),C_masks:$C-$((D-1)),D_preds:$D-$((E-1)),E_resolve:$E-$((F-1)),F_output:$F-$((DC-30)),decode_cell:$DC-$((DK-1)),tma_kernel:$TK-$((TE-1))" python profiles/ncu_lines.py gpurun_out/prof_r2_dense.ncu-rep 1000000 25) > profiles/ncu_r2_decode_nms_dense_summary.txt
PR=$(grep -n "// ---- producer" "$H" | tail -1 | cut -d: -f1)
CA=$(ln "A: decode own cell from the ring"); CA2=$(ln "A': compaction (utils.py:95, strict >)$"); CB=$(ln "B: stable descending rank (utils.py:98): r = #{s_j > s_i} + #{j < i : s_j = s_i}.$"); CC=$(ln "C: scatter to rank order (utils.py:24-32,40); class masks")
CD=$(ln "D: suppression words against same-class predecessors"); CE=$(ln "E: greedy keep flags, fixed point of keep\[q\] = !any(supp\[q\] & keep)$"); CF=$(ln "F: output slots (kws holds the final keep words)"); CEND=$(ln "^// host side")
(echo "# decode_nms_coop_kernel<80,3,float>, cfg5 data, 32,768 images (ncu --set full --clock-control none --import-source on), round 2, commit $REV (bucketed rank, straight-line suppresses())"
 python profiles/ncu_summarize.py gpurun_out/prof_r2_stress5_coop.ncu-rep 32768 | sed -n 1,24p; echo
 YH_BUCKET_FILE=yh_decode_nms_impl.cuh YH_BUCKETS="iou_test:$IO-$((IO+32)),producer:$PR-$((CA-25)),A_decode_wait:$CA-$((CA2-1)),A_compact:$CA2-$((CB-1)),B_rank_buckets:$CB-$((CC-1)),C_scatter_masks:$CC-$((CD-1)),D_preds:$CD-$((CE-1)),E_resolve:$CE-$((CF-1)),F_output:$CF-$((CEND-1)),decode_cell:$DC-$((DK-1))" python profiles/ncu_lines.py gpurun_out/prof_r2_stress5_coop.ncu-rep 32768 28) > profiles/ncu_r2_decode_nms_coop_cfg5_summary.txt
(echo "# loss_gather_kernel<true>, batch 4096 (cfg3) (ncu --set full --clock-control none --import-source on), round 2, commit $REV"
 python profiles/ncu_summarize.py gpurun_out/prof_r2_loss.ncu-rep 4096 | sed -n 1,24p; echo
 python profiles/ncu_lines.py gpurun_out/prof_r2_loss.ncu-rep 4096 20) > profiles/ncu_r2_loss_summary.txt
(echo "# map_radix_kernel on its COUNTING path (default), cfg4: 73,595 records of 5,000 images (ncu --set full --clock-control none --import-source on), round 2, commit $REV; units = records"
 python profiles/ncu_summarize.py gpurun_out/prof_r2_map_count_cfg4.ncu-rep 73595 | sed -n 1,24p; echo
 python profiles/ncu_lines.py gpurun_out/prof_r2_map_count_cfg4.ncu-rep 73595 24) > profiles/ncu_r2_map_count_cfg4_summary.txt
(echo "# loss_stream_kernel<true>, batch 16,384 (ncu --set full --clock-control none --import-source on), round 2, commit $REV"
 python profiles/ncu_summarize.py gpurun_out/prof_r2_loss_stream_b16k.ncu-rep 16384 | sed -n 1,24p; echo
 python profiles/ncu_lines.py gpurun_out/prof_r2_loss_stream_b16k.ncu-rep 16384 20) > profiles/ncu_r2_loss_stream_b16k_summary.txt
(echo "# map_radix_kernel with YH_MAP_COUNT=0 (radix passes only), cfg4: 73,595 records of 5,000 images (ncu --set full --clock-control none --import-source on), round 2, commit $REV; units = records"
 python profiles/ncu_summarize.py gpurun_out/prof_r2_map_radix_cfg4.ncu-rep 73595 | sed -n 1,24p; echo
 python profiles/ncu_lines.py gpurun_out/prof_r2_map_radix_cfg4.ncu-rep 73595 24) > profiles/ncu_r2_map_radix_cfg4_summary.txt
(echo "# map_radix_kernel, 14.7 M records of 1 M images (ncu --set full --clock-control none --import-source on), round 2, commit $REV; units = records"
 python profiles/ncu_summarize.py gpurun_out/prof_r2_map_radix_big.ncu-rep 14668340 | sed -n 1,24p; echo
 python profiles/ncu_lines.py gpurun_out/prof_r2_map_radix_big.ncu-rep 14668340 24) > profiles/ncu_r2_map_radix_1M_images_summary.txt
(echo "# eval_state_kernel<2,20,2> (update_state in one launch), cfg4: one batch of 5,000 images (ncu --set full --clock-control none --import-source on), round 2, commit $REV; units = images"
 python profiles/ncu_summarize.py gpurun_out/prof_r2_eval_state_cfg4.ncu-rep 5000 | sed -n 1,24p; echo
 python profiles/ncu_lines.py gpurun_out/prof_r2_eval_state_cfg4.ncu-rep 5000 20) > profiles/ncu_r2_eval_state_cfg4_summary.txt
(echo "# eval_update_kernel<true> (three-launch path, YH_EVAL_FUSED=0), cfg4: one batch of 5,000 images (ncu --set full --clock-control none --import-source on), round 2, commit $REV; units = images"
 python profiles/ncu_summarize.py gpurun_out/prof_r2_eval_update_cfg4.ncu-rep 5000 | sed -n 1,24p; echo
 python profiles/ncu_lines.py gpurun_out/prof_r2_eval_update_cfg4.ncu-rep 5000 20) > profiles/ncu_r2_eval_update_cfg4_summary.txt
python - <<PY
import json,subprocess,csv,io
raw=subprocess.run(["ncu","-i","gpurun_out/prof_r2_dense.ncu-rep","--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw))); h,u,v=rows[0],rows[1],rows[2]
def g(k):
    i=h.index(k); x=float(v[i].replace(",","")); return x*{"Gbyte":1e9,"Mbyte":1e6,"Kbyte":1e3,"byte":1}[u[i]]
rd,wr=g("dram__bytes_read.sum"),g("dram__bytes_write.sum")
json.dump({"decode_nms_tma_kernel_dram_bytes_per_launch":rd+wr,
 "source":f"ncu --set full --clock-control none, profiles/ncu_r2_decode_nms_dense_summary.txt (1M dense VOC images, one launch): dram__bytes_read.sum {rd/1e9:.6f} GB + dram__bytes_write.sum {wr/1e9:.6f} GB",
 "commit":"$REV","captured_by":"profiles/final_capture_r2.sh",
 "algorithmic_bytes_per_launch":6843757120},open("profiles/traffic.json","w"),indent=1)
print("traffic", rd+wr)
PY
