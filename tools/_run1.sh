#!/bin/bash
timeout 800 python tools/trsv_levels.py 2048 --smin 8,4,2 > gpurun_out/trsv_smin_2048.txt 2>&1
grep -E "^(L11|U11)|^\{" gpurun_out/trsv_smin_2048.txt | cut -c1-260
