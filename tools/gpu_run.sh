cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for mb in 0 131072 65536 262144 0; do
  echo "== mask-block $mb"
  timeout 200 python bench.py --frames 32 --steps 3 --warmup 2 --no-cpu-baseline --mask-block $mb 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
r = d['roofline']
print(d['value'], d['ms_per_step'], 'alg', r['achieved'], 'exec', r['mma_executed'], 'roweff', r['row_efficiency'], 'share', r['share_of_step'])"
done
