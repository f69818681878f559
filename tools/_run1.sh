python -m pytest tests/test_gpu_trsv.py tests/test_gpu_amg.py tests/test_gpu_gmres.py -x -q > gpurun_out/r2q_tests.log 2>&1; tail -4 gpurun_out/r2q_tests.log | cut -c1-300
python tools/amg_profile.py 512 2> gpurun_out/r2q_amg512.err | tee gpurun_out/r2q_amg512.json
python tools/amg_profile.py 2048 2> gpurun_out/r2q_amg2048.err | tee gpurun_out/r2q_amg2048.json
PSB_TRSV_NO_SUBWARP=1 python tools/amg_profile.py 2048 2>/dev/null | cut -c1-300
python tools/config_bench.py c2 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)['c2']; print([(x['config'][-22:], x['iters'], x['ref_iters'], round(1e3*x['gpu_ilut_apply_s'],3), x['hist_max_rel_err']) for x in d])"
