# slab-decomposed ASM over all GPUs of the box: peer-memory vs NCCL transport, vs one GPU (tools/multi_gpu_check.py)
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
timeout 600 python -m pytest tests -m gpu -x -q -k "slab" 2>&1 | tail -2
for n in 4096 8192; do for tr in nccl peer; do
THZ_SLAB_TRANSPORT=$tr THZ_SLAB_N=$n THZ_SLAB_C=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 tools/multi_gpu_check.py 2>&1 | grep -v -i "warn\|OMP\|\*\*\*" | grep -i "slab\|stages\|error\|Traceback" | tail -6
done; done
