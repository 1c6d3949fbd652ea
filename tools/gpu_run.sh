cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_configs.py tests/test_gpu_conv.py -x -q -m gpu > gpurun_out/pytest_morton.log 2>&1; echo "configs+conv rc=$?"
tail -4 gpurun_out/pytest_morton.log
for m in "" "--mask-no-morton" "" "--mask-no-morton"; do
  echo "== bench $m"
  timeout 200 python bench.py --frames 32 --steps 3 --warmup 2 --no-cpu-baseline $m 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
r = d['roofline']
print(d['value'], d['ms_per_step'], 'alg', r['achieved'], 'exec', r['mma_executed'], 'roweff', r['row_efficiency'], 'share', r['share_of_step'])"
done
