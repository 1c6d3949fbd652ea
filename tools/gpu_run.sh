set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -m gpu > gpurun_out/pytest_conv.log 2>&1; echo "conv rc=$?"
tail -3 gpurun_out/pytest_conv.log
B2ME_LIB_PATH=$GRAFT_REPO_ROOT/markerless-robot-camera-calibration_b200/lib_debug/libb2me.so timeout 300 python tools/conv_probe.py --frames 8 --shapes 27:384:384,1:256:1024 > gpurun_out/probe_x.log 2>&1; grep -E "^---|rank0 mma" gpurun_out/probe_x.log
timeout 600 python tools/conv_probe.py --frames 8 --reps 10 --out gpurun_out/probe11.json > gpurun_out/probe11.log 2>&1; echo "rc=$?"
tail -8 gpurun_out/probe11.log
