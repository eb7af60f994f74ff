cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 300 python tools/config_bench.py donn c2 czt 2>&1 | grep "^{"
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/bench_n2.json | cut -c1-400
fi
timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | grep "^{" | tee gpurun_out/bench_n1.json | cut -c1-300
