cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
THZ_SLAB_N=8192 THZ_SLAB_C=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 tools/multi_gpu_check.py 2>&1 | grep -v -i "warn\|OMP\|\*\*\*" | tail -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $NG --steps 10 --warmup 3 2>&1 | grep "^{" | cut -c1-330
