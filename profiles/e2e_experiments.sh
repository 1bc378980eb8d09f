# e2e experiments under torchrun (N ranks): bash profiles/e2e_experiments.sh N
N=${1:-8}
run() { env "$@" YH_BENCH_EXTRAS=0 YH_BENCH_SUSTAINED=0 YH_BENCH_E2E_HALF=0 YH_BENCH_E2E_IMAGES=500000 \
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 5 --warmup 3 2>/dev/null \
  | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); e=d['e2e']; print('  e2e %.2f M img/s  %.1f ms/step  concurrent bare H2D %.1f GB/s total  frac %.3f' % (e['value']/1e6, e['ms_per_step'], e['bare_pinned_h2d_all_ranks_at_once_GBps_total'], e['frac_of_concurrent_h2d_ceiling']))"; }
echo "baseline (regular pinned input)"; run A=1
echo "write-combined pinned input"; run YH_BENCH_E2E_WC=1
