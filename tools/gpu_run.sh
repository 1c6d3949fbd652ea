cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest18.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/pytest18.log
timeout 600 python bench.py > gpurun_out/bench_default2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_default2.log | cut -c1-250
