cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>&1 | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['step']['frac'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})"
