mkdir -p gpurun_out
P=$PWD/markerless-robot-camera-calibration_b200
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest8_full.log 2>&1
tail -5 gpurun_out/r2_pytest8_full.log > gpurun_out/r2_pytest8.log
timeout 900 python tools/conv_probe.py --frames 16 --reps 2 --rounds 6 --sweep --lib-b $P/lib_variants/libb2me_prev.so --shapes 27:384:384,8:384:384,1:416:384,1:384:256 > gpurun_out/r2_probe8_l1.log 2>&1
timeout 900 python tools/conv_probe.py --frames 16 --reps 2 --rounds 6 --sweep --lib-b $P/lib_variants/libb2me_prev.so --head 3 --shapes 1:256:1024 > gpurun_out/r2_probe8_head.log 2>&1
timeout 600 python tools/conv_probe.py --frames 16 --reps 2 --rounds 6 --sweep --lib-b $P/lib_variants/libb2me_prev.so --level 2 --shapes 27:384:384,27:128:128,27:32:32 > gpurun_out/r2_probe8_l2.log 2>&1
timeout 600 python bench.py --stages --conv-table gpurun_out/r2_conv_table8.json > gpurun_out/r2_bench8.log 2>&1
timeout 600 python bench.py --strong-frames 96 --steps 2 --warmup 1 > gpurun_out/r2_strong8.log 2>&1
timeout 600 python bench.py --vote --no-cpu-baseline --stages --steps 4 > gpurun_out/r2_bench8_vote.log 2>&1
tail -3 gpurun_out/r2_pytest8.log; grep "median" gpurun_out/r2_probe8_l1.log gpurun_out/r2_probe8_head.log gpurun_out/r2_probe8_l2.log | grep -E "default|sb" ; tail -c 300 gpurun_out/r2_bench8.log
