cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --frames 32 --steps 3 --warmup 3 --no-cpu-baseline --stages --kp-backbone pointnet2 > gpurun_out/bench_pn2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_pn2.log | cut -c1-200
timeout 600 python bench.py --frames 32 --steps 3 --warmup 3 --no-cpu-baseline --stages > gpurun_out/bench_v10_32.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_v10_32.log | cut -c1-200
