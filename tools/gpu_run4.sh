mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest4_full.log 2>&1
tail -5 gpurun_out/r2_pytest4_full.log > gpurun_out/r2_pytest4.log
timeout 900 python tools/conv_probe.py --frames 16 --reps 4 --sweep --shapes 27:384:384,1:256:1024,8:384:384 > gpurun_out/r2_probe4_l1.log 2>&1
timeout 600 python tools/conv_probe.py --frames 16 --reps 4 --sweep --level 2 --shapes 27:384:384 > gpurun_out/r2_probe4_l2.log 2>&1
timeout 600 python bench.py --stages --conv-table gpurun_out/r2_conv_table4.json > gpurun_out/r2_bench4.log 2>&1
timeout 300 python bench.py --config icp1k > gpurun_out/r2_icp1k4.log 2>&1
tail -3 gpurun_out/r2_pytest4.log; cat gpurun_out/r2_probe4_l1.log | tail -30; tail -c 300 gpurun_out/r2_bench4.log
