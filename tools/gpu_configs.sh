# the other BASELINE configs (tools/config_bench.py): DONN, C2 step, CZT, depth sweep, whole iteration
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 600 python tools/config_bench.py donn c2 czt zsweep iteration 2>&1 | grep "^{" | cut -c1-420
