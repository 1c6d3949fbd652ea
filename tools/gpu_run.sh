set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc; free -g | head -2
( time timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1 ) 2> gpurun_out/bench_default.time; echo "rc=$?"; cat gpurun_out/bench_default.time | tail -3
tail -1 gpurun_out/bench_default.log | cut -c1-200
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.log 2>&1 ) 2> gpurun_out/bench_ref.time; echo "rc=$?"; cat gpurun_out/bench_ref.time | tail -3
tail -1 gpurun_out/bench_ref.log | cut -c1-600
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke2.log
