#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_trsv.py tests/test_gpu_amg.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python tools/trsv_levels.py 2048 --levelset-ab > gpurun_out/trsv_levelset_2048.txt 2>&1
grep -E "^(L11|U11)|^\{" gpurun_out/trsv_levelset_2048.txt | cut -c1-250
timeout 600 python tools/amg_profile.py 2048 2>/dev/null
timeout 600 python tools/amg_profile.py 512 2>/dev/null
