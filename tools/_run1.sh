#!/bin/bash
PSB_TRSV_STAGE=1 timeout 600 python tools/trsv_levels.py 2048 > gpurun_out/trsv_levels_2048_stage3.txt 2>&1
grep -E "^(L11|U11)|^\{" gpurun_out/trsv_levels_2048_stage3.txt | cut -c1-200
