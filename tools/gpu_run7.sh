mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest7_full.log 2>&1
tail -5 gpurun_out/r2_pytest7_full.log > gpurun_out/r2_pytest7.log
timeout 600 python bench.py --stages --conv-table gpurun_out/r2_conv_table7.json --report gpurun_out/r2_eval_report7.json --crop both > gpurun_out/r2_bench7.log 2>&1
timeout 500 python bench.py --dtype tf32 --no-cpu-baseline --steps 3 > gpurun_out/r2_bench7_tf32.log 2>&1
timeout 900 python bench.py --config sweep --steps 3 --warmup 2 > gpurun_out/r2_sweep7.log 2>&1
timeout 600 python bench.py --strong-frames 96 --steps 2 --warmup 1 > gpurun_out/r2_strong7.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_ref7.log 2>&1
tail -3 gpurun_out/r2_pytest7.log; tail -c 300 gpurun_out/r2_bench7.log; tail -c 400 gpurun_out/r2_sweep7.log; tail -c 400 gpurun_out/r2_strong7.log
