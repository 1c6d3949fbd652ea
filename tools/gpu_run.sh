set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== product (try_wait)"
timeout 300 python tools/conv_probe.py --frames 8 --reps 10 --shapes 27:384:384,1:256:1024 2>&1 | tail -2
echo "== spin test_wait (profile build)"
export B2ME_LIB_PATH=$GRAFT_REPO_ROOT/markerless-robot-camera-calibration_b200/lib_debug/libb2me.so
for d in 0 7; do
B2ME_TC_DEBUG=$d timeout 300 python tools/conv_probe.py --frames 8 --reps 10 --shapes 27:384:384,1:256:1024 2>&1 | grep -E "^---|rank0 mma"
done
