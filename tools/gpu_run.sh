cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest17.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/pytest17.log
