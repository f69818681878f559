#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_amg.py tests/test_gpu_spmv.py tests/test_gpu_trsv.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python tools/amg_profile.py 2048 > gpurun_out/amg2048_c.json 2>gpurun_out/amg2048_c.err; cat gpurun_out/amg2048_c.json
timeout 600 python tools/amg_profile.py 512 2>/dev/null
