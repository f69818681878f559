python -m pytest tests -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1; tail -3 gpurun_out/r2s_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; tail -c 300 gpurun_out/r2s_bench.json; echo; tail -2 gpurun_out/r2s_bench.err
python bench.py --impl reference > gpurun_out/r2s_bench_ref.json 2> gpurun_out/r2s_bench_ref.err; cut -c1-400 gpurun_out/r2s_bench_ref.json
