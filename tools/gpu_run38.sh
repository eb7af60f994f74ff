cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
for n in 4096 8192; do
THZ_SLAB_N=$n THZ_SLAB_C=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 tools/multi_gpu_check.py 2>&1 | grep -v -i "warn\|OMP\|\*\*\*" | tail -4
done
python - <<'PY'
import torch
try:
    import torch.distributed._symmetric_memory as sm
    print("symmetric memory module:", [n for n in dir(sm) if not n.startswith('_')][:30])
except Exception as e:
    print("no symm mem", e)
print("p2p 0->1:", torch.cuda.can_device_access_peer(0,1))
PY
nvidia-smi topo -m 2>&1 | head -8
