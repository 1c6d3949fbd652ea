cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for m in 0 1; do
B2ME_TC_TMA=$m timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -x -q -m gpu > gpurun_out/pytest_12epi_$m.log 2>&1; echo "TMA=$m conv+model rc=$?"
tail -3 gpurun_out/pytest_12epi_$m.log
done
SH=27:384:384,1:416:384,8:384:384,1:256:1024,27:128:128,27:256:256,27:32:32
for v in new old new old newtma; do
  echo "== $v"
  unset B2ME_LIB_PATH; unset B2ME_TC_TMA
  if [ $v = old ]; then export B2ME_LIB_PATH=$GRAFT_REPO_ROOT/markerless-robot-camera-calibration_b200/lib_debug/libb2me.so; fi
  if [ $v = newtma ]; then export B2ME_TC_TMA=1; fi
  timeout 300 python tools/conv_probe.py --frames 8 --shapes $SH 2>&1 | tail -7
done
unset B2ME_TC_TMA
for v in new old; do
  echo "== bench $v"
  if [ $v = old ]; then export B2ME_LIB_PATH=$GRAFT_REPO_ROOT/markerless-robot-camera-calibration_b200/lib_debug/libb2me.so; else unset B2ME_LIB_PATH; fi
  timeout 300 python bench.py --frames 32 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
r = d['roofline']
print(d['value'], d['ms_per_step'], 'alg', r['achieved'], 'frac', r['frac'], 'exec', r['mma_executed'], 'share', r['share_of_step'])"
done
