set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest13.log 2>&1; echo "all rc=$?"
grep -E "rel err|passed|failed|Error|error" gpurun_out/pytest13.log | tail -20
