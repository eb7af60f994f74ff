cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
for l in ste gumbel psq full; do timeout 200 python examples/four_focal_spots.py --iters 600 --layer $l 2>&1 | tail -1; done
timeout 200 python examples/four_focal_spots.py --iters 2000 --layer ste --graph 2>&1 | tail -1
