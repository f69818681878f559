python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2f_tests.log 2>&1; tail -12 gpurun_out/r2f_tests.log | cut -c1-600
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
f() { grep -E "^rank|iter_us|Error|error" | tail -4; }
echo "== N=2 m=2048"; $TR --master-port 29704 tools/mega_timeline.py --gridm 2048 --out gpurun_out/tl2c_m2048 2>&1 | f
echo "== bratu 512 N=1"; python tools/dist_bratu.py --gridm 512 2>&1 | tail -1 | cut -c1-900
echo "== bratu 512 N=2"; $TR --master-port 29705 tools/dist_bratu.py --gridm 512 2>&1 | grep -E "^\{|Error|error" | tail -2 | cut -c1-900
