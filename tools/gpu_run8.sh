cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
SUMM='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), {k:round(v["ms_per_step"],2) for k,v in d["roofline"]["kernels"].items()}, "e2e", round(d["e2e"]["ms_per_step"],2))'
echo "== pairs"; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$SUMM"
echo "== no pairs"; THZ_NO_PAIRS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$SUMM"
python tools/profile_step.py --c 2 --steps 1 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:thz_p2 -c 6 -o gpurun_out/prof_r01d python tools/profile_step.py --c 2 --steps 1 > gpurun_out/ncu_full.log 2>&1
