cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python tools/config_bench.py donn c2 2>&1 | grep "^{"
THZ_NO_P2=1 timeout 300 python tools/config_bench.py donn c2 2>&1 | grep "^{"
