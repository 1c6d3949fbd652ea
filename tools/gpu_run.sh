set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest10.log 2>&1; echo "all rc=$?"
tail -15 gpurun_out/pytest10.log
timeout 600 python bench.py --frames 16 --steps 2 --warmup 2 --no-cpu-baseline --stages --conv-table gpurun_out/conv_v4.json > gpurun_out/bench_v4.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_v4.log | cut -c1-300
