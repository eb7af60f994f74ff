# block widths of the two intermediates (log2 columns per block; T2 0 = row-major)
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
for cfg in "2 0" "2 3" "2 2"; do set -- $cfg; echo "T1_LOG2=$1 T2_LOG2=$2"; THZ_T1_LOG2=$1 THZ_T2_LOG2=$2 timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>&1 | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"; done
