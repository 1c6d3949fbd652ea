cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_icp_eval --launch-skip 4 --launch-count 1 -o gpurun_out/prof_icp_r01 -f python bench.py --frames 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_icp.log 2>&1; echo "rc=$?"
