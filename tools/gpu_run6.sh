mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest6_full.log 2>&1
tail -5 gpurun_out/r2_pytest6_full.log > gpurun_out/r2_pytest6.log
timeout 900 python tools/conv_probe.py --frames 16 --reps 2 --rounds 5 --sweep --shapes 27:384:384,8:384:384,1:416:384 > gpurun_out/r2_probe6_l1.log 2>&1
timeout 600 python tools/conv_probe.py --frames 16 --reps 2 --rounds 5 --sweep --level 2 --shapes 27:384:384 > gpurun_out/r2_probe6_l2.log 2>&1
B2ME_LIB_PATH=$PWD/markerless-robot-camera-calibration_b200/lib_debug/libb2me.so timeout 600 python tools/conv_probe.py --frames 16 --reps 3 --shapes 27:384:384,1:256:1024 > gpurun_out/r2_roles6.log 2>&1
B2ME_LIB_PATH=$PWD/markerless-robot-camera-calibration_b200/lib_debug/libb2me.so B2ME_TC_DEBUG=1 timeout 600 python tools/conv_probe.py --frames 16 --reps 3 --shapes 27:384:384 > gpurun_out/r2_roles6_noA.log 2>&1
B2ME_LIB_PATH=$PWD/markerless-robot-camera-calibration_b200/lib_debug/libb2me.so B2ME_TC_DEBUG=4 timeout 600 python tools/conv_probe.py --frames 16 --reps 3 --shapes 27:384:384 > gpurun_out/r2_roles6_noMMA.log 2>&1
timeout 600 python bench.py --stages --conv-table gpurun_out/r2_conv_table6.json > gpurun_out/r2_bench6.log 2>&1
timeout 300 python bench.py --config icp1k > gpurun_out/r2_icp1k6.log 2>&1
tail -3 gpurun_out/r2_pytest6.log; grep "median" gpurun_out/r2_probe6_l1.log gpurun_out/r2_probe6_l2.log; grep -v Warn gpurun_out/r2_roles6.log | tail -14; tail -c 300 gpurun_out/r2_bench6.log
