#!/bin/bash
PSB_TRSV_STAGE=1 timeout 900 python -m pytest tests/test_gpu_trsv.py tests/test_gpu_amg.py -x -q -m gpu 2>&1 | tail -2
PSB_TRSV_STAGE=1 PSB_TRSV_KERNEL=grid timeout 900 python -m pytest tests/test_gpu_trsv.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python tools/trsv_levels.py 2048 --stage-ab > gpurun_out/trsv_stage2_2048.txt 2>&1
grep '^{' gpurun_out/trsv_stage2_2048.txt | cut -c1-250
