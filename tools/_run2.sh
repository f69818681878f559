TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
f() { grep -E "^rank|iter_us|Error|error" | tail -4; }
echo "== N=1 m=2896"; python tools/mega_timeline.py --gridm 2896 --out gpurun_out/tl1_m2896 2>&1 | f
echo "== N=2 m=4096"; $TR --master-port 29701 tools/mega_timeline.py --gridm 4096 --out gpurun_out/tl2_m4096 2>&1 | f
echo "== N=2 m=4096 flags=3"; PSB_MEGA_FLAGS=3 $TR --master-port 29702 tools/mega_timeline.py --gridm 4096 --out gpurun_out/tl2_m4096_f3 2>&1 | f
echo "== N=2 m=4096 MINB=5"; PSB_MEGA_MINB=5 $TR --master-port 29703 tools/mega_timeline.py --gridm 4096 --out gpurun_out/tl2_m4096_b5 2>&1 | f
echo "== N=2 m=2048"; $TR --master-port 29704 tools/mega_timeline.py --gridm 2048 --out gpurun_out/tl2_m2048 2>&1 | f
echo "== N=2 m=2048 nccl-mode off mega"; PSB_DIST_MEGA=0 $TR --master-port 29705 tools/mega_timeline.py --gridm 2048 --out gpurun_out/tl2_m2048_nomega 2>&1 | f
