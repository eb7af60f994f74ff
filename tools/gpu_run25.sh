cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for mode in inregister cached; do echo "mode $mode"; THZ_KERNEL_MODE=$mode timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>&1 | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline'].get('step'), d.get('kernels_ms', d['roofline']))"; done
