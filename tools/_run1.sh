python -m pytest tests/test_gpu_trsv.py tests/test_gpu_amg.py -x -q > gpurun_out/r2r_tests.log 2>&1; tail -3 gpurun_out/r2r_tests.log | cut -c1-300
python tools/amg_profile.py 2048 2> gpurun_out/r2r_amg2048.err | tee gpurun_out/r2r_amg2048.json
python tools/amg_profile.py 512 2>/dev/null
python tools/trsv_probe.py 2>/dev/null | tail -12
