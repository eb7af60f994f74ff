cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
SUMM='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), {k:round(v["ms_per_step"],2) for k,v in d["roofline"]["kernels"].items()})'
for pf in 0 1 2 4; do echo "== prefetch waves $pf"; THZ_PREFETCH=$pf timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$SUMM"; done
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
