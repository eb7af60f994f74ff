cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py 2>&1 | tail -12
