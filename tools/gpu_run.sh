cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for m in 0 1; do
B2ME_TC_TMA=$m timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -x -q -m gpu > gpurun_out/pytest_lean_$m.log 2>&1; echo "TMA=$m conv+model rc=$?"
tail -3 gpurun_out/pytest_lean_$m.log
done
for m in 0 1 0 1; do
  echo "== TMA=$m"
  B2ME_TC_TMA=$m timeout 300 python tools/conv_probe.py --frames 8 --shapes 27:384:384,1:416:384,27:32:32,8:384:384,1:256:1024,27:128:128,27:256:256 2>&1 | tail -7
done
for m in 0 1; do
  echo "== bench TMA=$m"
  B2ME_TC_TMA=$m timeout 300 python bench.py --frames 32 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['share_of_step'])"
done
