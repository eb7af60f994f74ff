cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 200 python tools/tc_timeline.py 2>&1 | tail -60
