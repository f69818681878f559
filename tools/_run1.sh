python -m pytest tests -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1; tail -4 gpurun_out/r2o_tests.log | cut -c1-300
f() { grep -E "^rank|iter_us|Error|error" | tail -2; }
for wts in "0" "" "1.05,1.03,0.98,0.94" "1.02,1.01,0.99,0.98"; do
  for m in 4096 1448; do
    echo "== weights='$wts' m=$m"
    if [ -z "$wts" ]; then python tools/mega_timeline.py --gridm $m --out gpurun_out/tlw_default_m$m 2>&1 | f
    else PSB_MEGA_WEIGHTS=$wts python tools/mega_timeline.py --gridm $m --out gpurun_out/tlw_${wts}_m$m 2>&1 | f; fi
  done
done
python tools/config_bench.py gmres4096 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)['gmres4096']; print('cgs2 fused-all', {k: (round(v['ms_per_iteration'],3), round(v['achieved_GBps'])) for k,v in d.items() if isinstance(v, dict)})"
