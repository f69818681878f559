N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
f() { grep -E "^rank 0|iter_us|Error|error" | head -2; }
echo "== bench N=$N"; $TR --master-port 29802 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 2> gpurun_out/r2n_bench_n$N.err | grep '^{' > gpurun_out/r2n_bench_n$N.json; python - <<PY
import json
d=json.load(open('gpurun_out/r2n_bench_n$N.json'))
c4=d.get('c4') or {}
print('N=%d value %.1f it/s (%.2f us/iter) frac %.3f e2e %s | c4 %.1f it/s frac %.3f | parity ok %s' % (d['n_gpus'], d['value'], 1e3*d['ms_per_step']/200, d['roofline']['frac'], d['e2e']['value'], c4.get('value',0), (c4.get('roofline') or {}).get('frac',0), d['parity']['ok']))
PY
echo "== timeline N=$N m=4096"; $TR --master-port 29803 tools/mega_timeline.py --gridm 4096 --out gpurun_out/tl${N}f_m4096 2>&1 | f
echo "== bratu 2048 N=$N"; $TR --master-port 29804 tools/dist_bratu.py --gridm 2048 2> gpurun_out/r2n_bratu_n$N.err | grep -E "^\{" | tee gpurun_out/r2n_bratu_n$N.json | cut -c1-900
