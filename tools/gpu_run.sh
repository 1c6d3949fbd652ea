set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest11.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/pytest11.log
timeout 600 python bench.py --frames 16 --steps 2 --warmup 2 --no-cpu-baseline --stages --conv-table gpurun_out/conv_v6.json > gpurun_out/bench_v6.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_v6.log | cut -c1-300
timeout 600 python bench.py --frames 32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v6_32.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_v6_32.log | cut -c1-300
