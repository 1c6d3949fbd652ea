cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pointnet2.py -x -q -m gpu -s > gpurun_out/pytest_pn2.log 2>&1; echo "rc=$?"
tail -25 gpurun_out/pytest_pn2.log
