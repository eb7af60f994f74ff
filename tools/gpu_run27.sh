cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "czt or toeplitz" 2>&1 | tail -3
timeout 200 python tools/config_bench.py czt 2>&1 | grep "^{" | cut -c1-200
timeout 200 python tools/czt_accuracy.py 2>&1 | tail -3
timeout 200 python tools/czt_accuracy.py 2048 1024 2 2>&1 | tail -3
