set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest15.log 2>&1; echo "all rc=$?"
tail -3 gpurun_out/pytest15.log
timeout 600 python bench.py --frames 32 --steps 3 --warmup 3 --no-cpu-baseline --stages > gpurun_out/bench_v9_32.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_v9_32.log | cut -c1-200
