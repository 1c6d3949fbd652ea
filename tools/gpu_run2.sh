set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --frames 16 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/bench_2gpu.log | cut -c1-400
