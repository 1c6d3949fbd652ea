set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --frames 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_r01f.log 2>&1; echo "rc=$?"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --frames 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r01f.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spconv_tc --launch-skip 42 --launch-count 3 -o gpurun_out/prof_tc_r01f -f python bench.py --frames 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_r01f.log 2>&1; echo "ncu full rc=$?"
