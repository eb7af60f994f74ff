cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
for hm in "256 128" "1024 512" "2048 1024"; do timeout 120 python tools/czt_accuracy.py $hm 2 2>&1 | tail -2; done
