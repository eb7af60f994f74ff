cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 180 python -m pytest tests -m gpu -q -x -k "czt" 2>&1 | tail -2
for hm in "2048 1024 2" "2048 1024 16"; do timeout 200 python tools/czt_accuracy.py $hm 2>&1 | grep "^tc"; done
