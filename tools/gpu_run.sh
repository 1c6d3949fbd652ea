cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_conv.py -x -q -m gpu > gpurun_out/pytest_smn.log 2>&1; rc=$?; echo "conv rc=$rc"
tail -3 gpurun_out/pytest_smn.log
if [ $rc -ne 0 ]; then exit 0; fi
SH=27:384:384,1:416:384,8:384:384,1:256:1024,27:128:128,27:256:256,27:32:32
for v in new old new old; do
  echo "== $v"
  unset B2ME_LIB_PATH
  if [ $v = old ]; then export B2ME_LIB_PATH=$GRAFT_REPO_ROOT/markerless-robot-camera-calibration_b200/lib_debug/libb2me.so; fi
  timeout 90 python tools/conv_probe.py --frames 8 --shapes $SH 2>&1 | tail -7
done
