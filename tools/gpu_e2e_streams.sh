cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
for n in 1 2 4; do echo "copy streams $n"; THZ_E2E_COPY_STREAMS=$n timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>&1 | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e'])"; done
