cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 1 2>&1 | tail -5 | cut -c1-400
