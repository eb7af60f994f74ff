cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "level_tables" 2>&1 | grep -v "^$" | tail -30
python examples/four_focal_spots.py --graph --iters 200 2>&1 | tail -1
THZ_NO_DOE_LUT=1 python examples/four_focal_spots.py --graph --iters 200 2>&1 | tail -1
python examples/four_focal_spots.py --iters 200 2>&1 | tail -1
