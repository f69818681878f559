# IC at m=2048: the host SuperLU setup takes minutes -> in the background while the rest runs
(python tools/ic_large.py 2048 > gpurun_out/r2e_ic2048.json 2> gpurun_out/r2e_ic2048.err; echo ic2048 done) &
ICPID=$!
python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; tail -4 gpurun_out/r2e_tests.log
python tools/config_bench.py gmres4096 > gpurun_out/r2e_gmres.json 2> gpurun_out/r2e_gmres.err; python -c "
import json; d=json.load(open('gpurun_out/r2e_gmres.json'))['gmres4096']; print('fused', {k: (round(v['ms_per_iteration'],3), round(v['achieved_GBps'])) for k,v in d.items() if isinstance(v, dict)})"
PSB_GMRES_NOFUSE=1 python tools/config_bench.py gmres4096 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)['gmres4096']; print('unfused', {k: (round(v['ms_per_iteration'],3), round(v['achieved_GBps'])) for k,v in d.items() if isinstance(v, dict)})"
python tools/amg_profile.py 512 2> gpurun_out/r2e_amg512.err | tee gpurun_out/r2e_amg512.json
python tools/amg_profile.py 2048 2> gpurun_out/r2e_amg2048.err | tee gpurun_out/r2e_amg2048.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_amg2048_launches.csv python tools/amg_profile.py 2048 --profile > gpurun_out/r2e_ncu_amg.log 2>&1; tail -2 gpurun_out/r2e_ncu_amg.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench1.json 2> gpurun_out/r2e_bench1.err; tail -c 600 gpurun_out/r2e_bench1.json; tail -3 gpurun_out/r2e_bench1.err
wait $ICPID
cat gpurun_out/r2e_ic2048.json; tail -3 gpurun_out/r2e_ic2048.err
