python -m pytest tests/test_gpu_amg.py tests/test_gpu_trsv.py -x -q > gpurun_out/r2h_tests.log 2>&1; tail -3 gpurun_out/r2h_tests.log
PSB_SWEEP_LOOKAHEAD="2 4 8 16 32 48 96" python tools/amg_profile.py 2048 2> gpurun_out/r2h_amg2048.err | tee gpurun_out/r2h_amg2048.json
PSB_SWEEP_LOOKAHEAD="2 4 8 16 32 48 96" python tools/amg_profile.py 512 2> gpurun_out/r2h_amg512.err | tee gpurun_out/r2h_amg512.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_amg2048_launches.csv python tools/amg_profile.py 2048 --profile > gpurun_out/r2h_ncu_amg.log 2>&1; tail -2 gpurun_out/r2h_ncu_amg.log
