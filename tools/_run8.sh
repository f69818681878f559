N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== dist_worker N=$N"; $TR --master-port 29801 tests/dist_worker.py > gpurun_out/r2i_worker_n$N.log 2>&1; echo rc=$?; grep -E "^lap|^dh|^bratu|^newton|FAIL" gpurun_out/r2i_worker_n$N.log | cut -c1-300
echo "== bench N=$N"; $TR --master-port 29802 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > gpurun_out/r2i_bench_n$N.json 2> gpurun_out/r2i_bench_n$N.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2i_bench_n$N.json'))
    c4=d.get('c4') or {}
    print('N=%d value %.1f it/s (%.2f us/iter) frac %.3f e2e %s | c4 %.1f it/s frac %.3f | parity %s' % (d['n_gpus'], d['value'], 1e3*d['ms_per_step']/200, d['roofline']['frac'], d['e2e']['value'], c4.get('value',0), (c4.get('roofline') or {}).get('frac',0), {k:d['parity'][k] for k in ('ok','hist_rel_err','spmv_bitexact','second_solve_bitexact','iters')}))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2i_bench_n$N.err').read()[-1500:])
PY
f() { grep -E "^rank 0|iter_us|Error|error" | head -3; }
echo "== timeline N=$N m=4096"; $TR --master-port 29803 tools/mega_timeline.py --gridm 4096 --out gpurun_out/tl${N}_m4096 2>&1 | f
echo "== bratu 2048 N=$N"; $TR --master-port 29804 tools/dist_bratu.py --gridm 2048 2> gpurun_out/r2i_bratu_n$N.err | grep -E "^\{" | tee gpurun_out/r2i_bratu_n$N.json | cut -c1-1200
tail -2 gpurun_out/r2i_bratu_n$N.err | cut -c1-300
