cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python tools/config_bench.py donn c2 2>&1 | grep "^{"
timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | grep "^{" | tee gpurun_out/bench_n1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['e2e'], d['gpu_launches'], d['cpu_baseline'])"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-250
