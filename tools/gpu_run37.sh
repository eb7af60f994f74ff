cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
for c in 0 32 64 128 256; do echo "chunk $c"; THZ_BC_CHUNK=$c timeout 200 python tools/config_bench.py donn 2>&1 | grep "^{"; done
