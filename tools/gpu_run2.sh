set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_4gpu.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/bench_4gpu.log | cut -c1-300
