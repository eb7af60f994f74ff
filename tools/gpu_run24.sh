cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "czt or toeplitz" 2>&1 | tail -8
for dbg in 0 1 2; do echo "debug $dbg"; THZ_CZT_DEBUG=$dbg timeout 200 python tools/config_bench.py czt 2>&1 | grep "^{" | cut -c1-200; done
timeout 200 python tools/czt_accuracy.py 2>&1 | tail -8
timeout 200 python tools/tc_timeline.py 2>&1 | tail -50
