cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
DBG=$GRAFT_REPO_ROOT/markerless-robot-camera-calibration_b200/lib_debug/libb2me.so
B2ME_LIB_PATH=$DBG timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2; do
echo "== product"; timeout 300 python tools/conv_probe.py --frames 8 --reps 10 2>&1 | tail -5
echo "== variant"; B2ME_LIB_PATH=$DBG timeout 300 python tools/conv_probe.py --frames 8 --reps 10 2>&1 | tail -5
done
