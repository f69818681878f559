#!/bin/bash
# per-iteration time at the per-GPU size of an N-way strong-scaling split, on ONE GPU
for m in 4096 2896 2048 1448; do
  PSB_BENCH_M=$m PSB_CPU_ITERS=1 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('single m=$m n=%d  %.1f us/iter  spmv %.1f us  frac %.3f' % (d['config']['n'], 1e3*d['iter_roofline']['ms_per_iteration'], 1e3*d['roofline']['ms_per_launch'], d['iter_roofline']['frac']))"
  PSB_BENCH_M=$m PSB_BENCH_WORKLOAD=c3d PSB_BENCH_SKIP_E2E=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 1 --steps 5 --warmup 3 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dist   m=$m n=%d  %.1f us/iter  frac %.3f' % (d['config']['n'], 1e3*d['roofline']['ms_per_iteration'], d['roofline']['frac']))"
done
