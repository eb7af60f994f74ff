cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 300 python tools/config_bench.py iteration 2>&1 | grep "^{"
