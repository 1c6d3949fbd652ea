mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest16_full.log 2>&1
tail -5 gpurun_out/r2_pytest16_full.log > gpurun_out/r2_pytest16.log
timeout 300 python bench.py --config icp1k > gpurun_out/r2_icp1k16.log 2>&1
timeout 600 python bench.py --stages --no-cpu-baseline > gpurun_out/r2_bench16.log 2>&1
tail -3 gpurun_out/r2_pytest16.log; grep "^{" gpurun_out/r2_icp1k16.log | cut -c1-400; tail -c 600 gpurun_out/r2_bench16.log
