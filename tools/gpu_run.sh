cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
DBG=$GRAFT_REPO_ROOT/markerless-robot-camera-calibration_b200/lib_debug/libb2me.so
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2; do
echo "== product (reflected keys)"; timeout 300 python tools/conv_probe.py --frames 8 --reps 10 --shapes 27:384:384,8:384:384 2>&1 | tail -2
echo "== plain keys"; B2ME_LIB_PATH=$DBG timeout 300 python tools/conv_probe.py --frames 8 --reps 10 --shapes 27:384:384,8:384:384 2>&1 | tail -2
done
echo "== bench product"; timeout 600 python bench.py --frames 32 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120
echo "== bench plain"; B2ME_LIB_PATH=$DBG timeout 600 python bench.py --frames 32 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-120
