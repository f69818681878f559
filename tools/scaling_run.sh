#!/bin/bash
# usage: tools/scaling_run.sh OUTFILE "N:workload:mode:skip_e2e" ...
# runs bench.py under torchrun for each spec and appends the JSON lines to OUTFILE
out=$1; shift
mkdir -p "$(dirname "$out")"
port=29600
for spec in "$@"; do
  IFS=: read -r n wl mode skip <<< "$spec"
  port=$((port+1))
  echo "== $spec" >&2
  PSB_BENCH_WORKLOAD=$wl PSB_DIST_MODE=$mode PSB_BENCH_SKIP_E2E=$skip timeout 600 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus "$n" --steps 3 --warmup 3 2> "$out.$n.$wl.$mode.err" | grep '^{' >> "$out"
  tail -2 "$out.$n.$wl.$mode.err" | grep -i -E "error|Traceback|fail" >&2
done
python - "$out" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    print('%d GPUs  %-10s %-12s value %9.1f it/s  ms/iter %.4f  frac %.3f  e2e %s' % (
        d['n_gpus'], d['config']['workload'][:10], d['config']['parallelism'].split(': ')[-1], d['value'],
        d['roofline']['ms_per_iteration'], d['roofline']['frac'], d['e2e']['value']))
PY
