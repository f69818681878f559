#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_amg.py tests/test_gpu_spmv.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python tools/amg_profile.py 2048 --ops > gpurun_out/amg2048_d.json 2>gpurun_out/amg2048_d.err; cat gpurun_out/amg2048_d.json
