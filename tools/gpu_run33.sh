cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 300 python tools/config_bench.py zsweep c2 donn 2>&1 | grep "^{"
python - <<'PY'
import cProfile, pstats, torch, sys
sys.path.insert(0,'.')
from quantizationawarethzdoe_b200 import ASM_prop, ElectricField
dev=torch.device('cuda:0')
asm=ASM_prop(z_distance=0.1, device=dev); asm.check_Zc=False
x=torch.randn(1,1,1000,1000,dtype=torch.complex64,device=dev)
f=ElectricField(x,wavelengths=torch.tensor([1e-3],device=dev),spacing=torch.tensor([1e-3,1e-3],device=dev),device=dev)
for i in range(3):
    asm.z=0.05+i*0.001; asm(f)
torch.cuda.synchronize()
pr=cProfile.Profile(); pr.enable()
for i in range(50):
    asm.z=0.06+i*0.001; asm(f)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
PY
