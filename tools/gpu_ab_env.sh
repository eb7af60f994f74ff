# usage: bash tools/gpu_ab_env.sh VAR   -- full GPU tests, then the bench with VAR unset and VAR=1
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in 0 1; do echo "$1=$v"; env $1=$v timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>&1 | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"; done
