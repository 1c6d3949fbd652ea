mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest5_full.log 2>&1
tail -5 gpurun_out/r2_pytest5_full.log > gpurun_out/r2_pytest5.log
timeout 900 python tools/conv_probe.py --frames 16 --reps 2 --rounds 5 --sweep --shapes 27:384:384,8:384:384 > gpurun_out/r2_probe5_l1.log 2>&1
timeout 600 python tools/conv_probe.py --frames 16 --reps 2 --rounds 5 --sweep --level 2 --shapes 27:384:384 > gpurun_out/r2_probe5_l2.log 2>&1
B2ME_LIB_PATH=$PWD/markerless-robot-camera-calibration_b200/lib_debug/libb2me.so timeout 600 python tools/conv_probe.py --frames 16 --reps 3 --shapes 27:384:384 > gpurun_out/r2_roles5_far.log 2>&1
B2ME_LIB_PATH=$PWD/markerless-robot-camera-calibration_b200/lib_debug/libb2me.so timeout 600 python tools/conv_probe.py --frames 16 --reps 3 --shapes 27:384:384 --prefetch near > gpurun_out/r2_roles5_near.log 2>&1
timeout 600 python bench.py --stages --conv-table gpurun_out/r2_conv_table5.json > gpurun_out/r2_bench5.log 2>&1
timeout 300 python bench.py --config icp1k > gpurun_out/r2_icp1k5.log 2>&1
timeout 400 python bench.py --dtype tf32 --no-cpu-baseline --steps 3 > gpurun_out/r2_bench5_tf32.log 2>&1
tail -3 gpurun_out/r2_pytest5.log; grep "median" gpurun_out/r2_probe5_l1.log gpurun_out/r2_probe5_l2.log; tail -c 300 gpurun_out/r2_bench5.log
