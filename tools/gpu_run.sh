cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest19.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/pytest19.log
timeout 600 python bench.py > gpurun_out/bench_default3.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_default3.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref3.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref3.log | cut -c1-600
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
