python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; tail -6 gpurun_out/r2j_tests.log | cut -c1-400
python tools/dist_bratu.py --gridm 2048 2> gpurun_out/r2j_bratu_n1.err | grep -E "^\{" | tee gpurun_out/r2j_bratu_n1.json | cut -c1-1200
python tools/mega_timeline.py --gridm 1448 --out gpurun_out/tl1c_m1448 2>&1 | grep -E "^rank|iter_us"
