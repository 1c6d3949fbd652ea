mkdir -p gpurun_out
P=$PWD/markerless-robot-camera-calibration_b200
V=$P/lib_variants
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest13_full.log 2>&1
tail -5 gpurun_out/r2_pytest13_full.log > gpurun_out/r2_pytest13.log
PR="timeout 900 python tools/conv_probe.py --frames 16 --reps 2 --rounds 6 --sweep"
$PR --lib-b $V/libb2me_ns192.so --shapes 27:384:384,8:384:384 > gpurun_out/r2_probe13_l1.log 2>&1
$PR --lib-b $V/libb2me_ns192.so --level 2 --shapes 27:384:384 > gpurun_out/r2_probe13_l2.log 2>&1
timeout 600 python bench.py --stages --conv-table gpurun_out/r2_conv_table13.json > gpurun_out/r2_bench13.log 2>&1
timeout 600 python bench.py --depth 2 --no-cpu-baseline > gpurun_out/r2_bench13_depth2.log 2>&1
B2ME_LIB_PATH=$V/libb2me_ns192.so timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench13_ns192.log 2>&1
timeout 600 python bench.py --tc-path cpasync --no-cpu-baseline > gpurun_out/r2_bench13_cpasync.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench13_again.log 2>&1
tail -3 gpurun_out/r2_pytest13.log; grep -h "median" gpurun_out/r2_probe13_*.log | cut -c1-100
for f in r2_bench13 r2_bench13_depth2 r2_bench13_ns192 r2_bench13_cpasync r2_bench13_again; do echo $f; grep "^{" gpurun_out/$f.log | cut -c1-200; done
