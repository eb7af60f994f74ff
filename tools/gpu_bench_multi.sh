cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $NG --steps 10 --warmup 3 2>&1 | grep "^{" | cut -c1-420
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus $NG --steps 1 --warmup 1 2>&1 | grep "^{" | cut -c1-300
