#!/bin/bash
# Runs on the GPU box (gpurun): the plain bench lines first, then the ncu passes whose summaries profiles/refresh_summaries.sh rebuilds.
# Numbers printed under ncu are never bench values; every profiled command has exited 0 without ncu first.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --impl reference > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err || exit 1
python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err || exit 1
python bench.py --steps 2 --warmup 1 > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
python profiles/prof_decode_nms.py dense > gpurun_out/plain_dense_final.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:decode_nms_tma_kernel --launch-skip 2 --launch-count 1 \
    -o gpurun_out/prof_r1_dense_final -f python profiles/prof_decode_nms.py dense > gpurun_out/ncu_dense_final.log 2>&1
YH_PROF_IMAGES=32768 python profiles/prof_decode_nms.py stress5 > gpurun_out/plain_stress5_final.log 2>&1 || exit 1
YH_PROF_IMAGES=32768 ncu --set full --clock-control none --import-source on -k regex:decode_nms_coop_kernel --launch-skip 2 --launch-count 1 \
    -o gpurun_out/prof_r1_stress5_coop -f python profiles/prof_decode_nms.py stress5 > gpurun_out/ncu_stress5.log 2>&1
python profiles/prof_loss.py > gpurun_out/plain_loss_final.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:loss_kernel --launch-skip 30 --launch-count 1 \
    -o gpurun_out/prof_r1_loss2 -f python profiles/prof_loss.py > gpurun_out/ncu_loss.log 2>&1
tail -n 2 gpurun_out/plain_dense_final.log gpurun_out/plain_stress5_final.log gpurun_out/plain_loss_final.log
