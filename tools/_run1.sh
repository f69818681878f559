export PSB_BENCH_SKIP_IC=1 PSB_BENCH_SKIP_C4=1 PSB_BENCH_SKIP_PARITY=1 PSB_CPU_ITERS=4
python bench.py --steps 2 --warmup 3 > gpurun_out/r2p_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2p_ncu1.log 2>&1
python bench.py --steps 2 --warmup 3 > gpurun_out/r2p_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pcg_mega -s 1 -c 1 -o gpurun_out/r2p_prof python bench.py --steps 2 --warmup 3 > gpurun_out/r2p_ncu2.log 2>&1
tail -c 400 gpurun_out/r2p_plain.log; tail -3 gpurun_out/r2p_ncu2.log; ls -la gpurun_out/r2p_prof.ncu-rep
