set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -m gpu > gpurun_out/pytest_conv.log 2>&1; echo "conv rc=$?"
tail -5 gpurun_out/pytest_conv.log
timeout 600 python bench.py --frames 16 --steps 2 --warmup 2 --no-cpu-baseline --stages --conv-table gpurun_out/conv_v3.json > gpurun_out/bench_v3.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_v3.log | cut -c1-300
