mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi --query-gpu=index,name,clocks.sm,power.limit --format=csv > gpurun_out/r2_8gpu_smi.txt 2>&1
timeout 600 $T bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_8gpu.log 2>&1
timeout 600 $T bench.py --gpus 8 --strong-frames 256 --steps 3 --warmup 1 > gpurun_out/r2_strong256_8gpu.log 2>&1
timeout 600 $T bench.py --gpus 8 --strong-frames 1000 --steps 1 --warmup 1 > gpurun_out/r2_strong1000_8gpu.log 2>&1
timeout 600 $T bench.py --gpus 8 --config sweep --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_sweep_8gpu.log 2>&1
timeout 300 python bench.py --strong-frames 256 --steps 2 --warmup 1 > gpurun_out/r2_strong256_1gpu.log 2>&1
timeout 300 python bench.py --depth 2 --no-cpu-baseline --steps 10 > gpurun_out/r2_bench12_depth2.log 2>&1
for f in r2_bench_8gpu r2_strong256_8gpu r2_strong1000_8gpu r2_sweep_8gpu r2_strong256_1gpu r2_bench12_depth2; do echo $f; grep "^{" gpurun_out/$f.log | cut -c1-250; done
