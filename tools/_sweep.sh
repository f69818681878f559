python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; tail -3 gpurun_out/r2b_tests.log
for cfg in "5 4" "5 5" "5 6" "5 8" "4 4" "4 5" "4 6" "4 8"; do
  set -- $cfg
  for m in 1448 4096; do
    echo "== MINB=$1 G=$2 m=$m"
    PSB_MEGA_MINB=$1 PSB_MEGA_G=$2 python tools/mega_timeline.py --m $m --out gpurun_out/tl_b$1_g$2_m$m 2>&1 | tail -2
  done
done
echo "== 3D m=256 (5,4) (4,8)"
PSB_MEGA_MINB=5 PSB_MEGA_G=4 python tools/mega_timeline.py --dim 3 --m 256 --out gpurun_out/tl3_b5g4 2>&1 | tail -2
PSB_MEGA_MINB=4 PSB_MEGA_G=8 python tools/mega_timeline.py --dim 3 --m 256 --out gpurun_out/tl3_b4g8 2>&1 | tail -2
