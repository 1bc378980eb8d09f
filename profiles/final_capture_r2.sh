#!/bin/bash
# Round 2.  Runs on the GPU box (gpurun): the plain bench lines first, then the ncu passes whose summaries
# profiles/refresh_summaries_r2.sh rebuilds in the build container.  Numbers printed under ncu are never bench values;
# every profiled command has exited 0 without ncu first.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err || exit 1
python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err || exit 1
python bench.py --steps 2 --warmup 1 > /dev/null 2>&1 || exit 1
YH_BENCH_SUSTAINED=20 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2_final.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench_r2.log 2>&1
python profiles/prof_decode_nms.py dense > gpurun_out/plain_dense_r2.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:decode_nms_tma_kernel --launch-skip 2 --launch-count 1 \
    -o gpurun_out/prof_r2_dense -f python profiles/prof_decode_nms.py dense > gpurun_out/ncu_dense_r2.log 2>&1
YH_PROF_IMAGES=32768 python profiles/prof_decode_nms.py stress5 > gpurun_out/plain_stress5_r2.log 2>&1 || exit 1
YH_PROF_IMAGES=32768 ncu --set full --clock-control none --import-source on -k regex:decode_nms_coop_kernel --launch-skip 2 --launch-count 1 \
    -o gpurun_out/prof_r2_stress5_coop -f python profiles/prof_decode_nms.py stress5 > gpurun_out/ncu_stress5_r2.log 2>&1
python profiles/prof_loss.py > gpurun_out/plain_loss_r2.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:loss_gather_kernel --launch-skip 30 --launch-count 1 \
    -o gpurun_out/prof_r2_loss -f python profiles/prof_loss.py > gpurun_out/ncu_loss_r2.log 2>&1
YH_PROF_BATCH=16384 python profiles/prof_loss.py > gpurun_out/plain_loss_b16k_r2.log 2>&1 || exit 1
YH_PROF_BATCH=16384 ncu --set full --clock-control none --import-source on -k regex:loss_stream_kernel --launch-skip 30 --launch-count 1 \
    -o gpurun_out/prof_r2_loss_stream_b16k -f python profiles/prof_loss.py > gpurun_out/ncu_loss_stream_r2.log 2>&1
python profiles/prof_map_count.py 5000 20000 50000 100000 > gpurun_out/plain_map_count_r2.log 2>&1 || exit 1
python profiles/prof_map_host.py > gpurun_out/plain_map_host_r2.log 2>&1 || exit 1
python profiles/prof_map.py cfg4 > gpurun_out/plain_map_r2.log 2>&1 || exit 1
python profiles/prof_map.py big >> gpurun_out/plain_map_r2.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:map_radix_kernel --launch-skip 4 --launch-count 1 \
    -o gpurun_out/prof_r2_map_count_cfg4 -f python profiles/prof_map.py cfg4 > gpurun_out/ncu_map_r2.log 2>&1
YH_MAP_COUNT=0 ncu --set full --clock-control none --import-source on -k regex:map_radix_kernel --launch-skip 4 --launch-count 1 \
    -o gpurun_out/prof_r2_map_radix_cfg4 -f python profiles/prof_map.py cfg4 >> gpurun_out/ncu_map_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_state_kernel --launch-skip 4 --launch-count 1 \
    -o gpurun_out/prof_r2_eval_state_cfg4 -f python profiles/prof_map.py cfg4 >> gpurun_out/ncu_map_r2.log 2>&1
YH_EVAL_FUSED=0 ncu --set full --clock-control none --import-source on -k regex:eval_update_kernel --launch-skip 4 --launch-count 1 \
    -o gpurun_out/prof_r2_eval_update_cfg4 -f python profiles/prof_map.py cfg4 >> gpurun_out/ncu_map_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:map_radix_kernel --launch-skip 2 --launch-count 1 \
    -o gpurun_out/prof_r2_map_radix_big -f python profiles/prof_map.py big >> gpurun_out/ncu_map_r2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2_map.csv \
    python profiles/prof_map.py cfg4 >> gpurun_out/ncu_map_r2.log 2>&1
tail -n 3 gpurun_out/plain_dense_r2.log gpurun_out/plain_stress5_r2.log gpurun_out/plain_loss_r2.log gpurun_out/plain_loss_b16k_r2.log gpurun_out/plain_map_r2.log gpurun_out/plain_map_count_r2.log
