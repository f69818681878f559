#!/bin/bash
# usage: tools/scaling_run.sh OUTFILE "N:mode:skip_e2e:skip_c4" ...
# runs bench.py (under torchrun for N > 1) for each spec and appends the JSON lines to OUTFILE
out=$1; shift
mkdir -p "$(dirname "$out")"
port=29600
for spec in "$@"; do
  IFS=: read -r n mode skip skipc4 <<< "$spec"
  port=$((port+1))
  echo "== $spec" >&2
  if [ "$n" = "1" ]; then
    PSB_BENCH_SKIP_E2E=${skip:-0} PSB_BENCH_SKIP_C4=${skipc4:-0} PSB_BENCH_SKIP_IC=1 timeout 900 \
      python bench.py --gpus 1 --steps ${STEPS:-5} --warmup 3 2> "$out.$n.$mode.err" | grep '^{' >> "$out"
  else
    PSB_DIST_MODE=$mode PSB_BENCH_SKIP_E2E=${skip:-0} PSB_BENCH_SKIP_C4=${skipc4:-0} timeout 900 \
      python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 \
      --master-port $port bench.py --gpus "$n" --steps ${STEPS:-5} --warmup 3 2> "$out.$n.$mode.err" | grep '^{' >> "$out"
  fi
  tail -2 "$out.$n.$mode.err" | grep -i -E "error|Traceback|fail" >&2
done
python - "$out" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    c4 = d.get('c4') or {}
    par = d.get('parity') or {}
    print('%d GPUs  %-12s value %9.1f it/s  us/iter %.2f  frac %.3f  e2e %s | c4 %s it/s | parity ok=%s hist %.1e' % (
        d['n_gpus'], d.get('collectives', '-'), d['value'], 1e3 * d['ms_per_step'] / d['config']['iters_per_step'],
        d['roofline']['frac'], d['e2e']['value'], c4.get('value'), par.get('ok'), par.get('hist_rel_err', float('nan'))))
PY
