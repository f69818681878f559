#!/bin/bash
python tools/amg_profile.py 2048 --profile > gpurun_out/plain_amg.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/round2b_amg2048_launches.csv python tools/amg_profile.py 2048 --profile > gpurun_out/ncu_amg.log 2>&1
tail -2 gpurun_out/ncu_amg.log; wc -l gpurun_out/round2b_amg2048_launches.csv
