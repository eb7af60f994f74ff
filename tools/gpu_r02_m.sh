cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
rm -f gpurun_out/parity_errors.jsonl
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
cp gpurun_out/parity_errors.jsonl gpurun_out/r02/parity_errors_m.jsonl
python - <<PY
import time, torch, sys
sys.path.insert(0, ".")
from quantizationawarethzdoe_b200 import ASM_prop, ElectricField
dev=torch.device("cuda:0")
for n in (100,1000):
    for mode in ("auto","inregister","cached"):
        asm = ASM_prop(z_distance=0.1, device=dev, padding_scale=2 if n==100 else None, kernel_mode=mode); asm.check_Zc=False
        x=torch.randn(1,1,n,n,dtype=torch.complex64,device=dev)
        f=ElectricField(x, wavelengths=torch.tensor([1e-3],device=dev), spacing=torch.tensor([1e-3,1e-3],device=dev), device=dev)
        zs=[0.05+0.001*i for i in range(100)]
        for z in zs[:5]:
            asm.z=z; asm(f)
        torch.cuda.synchronize(); t0=time.perf_counter()
        for z in zs:
            asm.z=z; y=asm(f).data
        torch.cuda.synchronize()
        print(n, mode, "ms per z %.3f"%((time.perf_counter()-t0)/len(zs)*1e3), asm.resolved_kernel_mode)
t0=time.perf_counter()
asm = ASM_prop(z_distance=0.1, device=dev, kernel_mode="cached"); asm.check_Zc=False
x=torch.randn(1,16,2048,2048,dtype=torch.complex64,device=dev)
lam=[1e-3*(1+0.01*c) for c in range(16)]
y=asm(ElectricField(x, wavelengths=lam, spacing=0.5e-3, device=dev)).data
torch.cuda.synchronize()
print("bench-size cached plan build + first forward: %.2f s"%(time.perf_counter()-t0))
PY
