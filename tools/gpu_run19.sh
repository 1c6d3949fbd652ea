mkdir -p gpurun_out
timeout 600 python bench.py --stages --no-cpu-baseline --conv-table gpurun_out/r2_conv_table19.json > gpurun_out/r2_bench19.log 2>&1
tail -c 300 gpurun_out/r2_bench19.log
