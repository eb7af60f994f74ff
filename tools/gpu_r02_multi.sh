# round 2: multi-GPU call (gpurun --gpus N): real multi-rank pytest, bench at N ranks with the secondary slab / DONN-DP lines
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -s > gpurun_out/r02/pytest_multi_n$N.log 2>&1
tail -12 gpurun_out/r02/pytest_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02/bench_n$N.json 2> gpurun_out/r02/bench_n$N.err
tail -c 3500 gpurun_out/r02/bench_n$N.json; tail -5 gpurun_out/r02/bench_n$N.err
