# round-end validation on a GPU box: gpurun --timeout 1700 -- 'bash tools/gpu_run.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 500 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_default.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "reference arm rc=$?"; tail -c 400 gpurun_out/bench_ref.log
