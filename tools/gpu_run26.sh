cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
echo GEMM1; timeout 200 python tools/tc_timeline.py 4 2>&1 | tail -44
echo GEMM2; timeout 200 python tools/tc_timeline.py 3 2>&1 | tail -44
