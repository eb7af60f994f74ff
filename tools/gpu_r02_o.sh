cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for l in ste gumbel psq full; do timeout 200 python examples/four_focal_spots.py --iters 600 --layer $l 2>&1 | tail -1; done
timeout 200 python examples/four_focal_spots.py --iters 2000 --layer ste --graph 2>&1 | tail -1
python tools/config_bench.py iteration 2>&1 | tail -1
