set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest14.log 2>&1; echo "all rc=$?"
tail -3 gpurun_out/pytest14.log
timeout 600 python tools/conv_probe.py --frames 8 --reps 10 --out gpurun_out/probe8.json > gpurun_out/probe8.log 2>&1; echo "rc=$?"
tail -8 gpurun_out/probe8.log
timeout 600 python bench.py --frames 32 --steps 3 --warmup 3 --no-cpu-baseline --stages > gpurun_out/bench_v8_32.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_v8_32.log | cut -c1-200
