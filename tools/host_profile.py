"""cProfile of the eager optimisation iteration (host overhead per iteration at small grids)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizationawarethzdoe_b200 import ASM_prop, ElectricField, FusedAdam, STEQuantizedDOELayer, normalized_intensity_mse
mm = 1e-3
dev = torch.device("cuda:0")
n = 512
doe = STEQuantizedDOELayer(dict(doe_size=[n, n], doe_dxy=0.5 * mm, doe_level=4, height_constraint_max=1 * mm, tolerance=None, material=[2.66, 0.003]), {}, device=dev)
asm = ASM_prop(z_distance=0.1, device=dev)
asm.check_Zc = False
opt = FusedAdam(doe.parameters(), lr=0.02)
x = torch.randn(1, 1, n, n, dtype=torch.complex64, device=dev)
target = torch.rand(1, 1, n, n, device=dev)
lam_t, sp_t = torch.tensor([1 * mm], device=dev), torch.tensor([0.5 * mm, 0.5 * mm], device=dev)
field = ElectricField(x, wavelengths=lam_t, spacing=sp_t, device=dev)

def it():
    y = asm(doe(field)).data
    loss = normalized_intensity_mse(y, target)
    opt.zero_grad(set_to_none=False)
    loss.backward()
    opt.step()

for _ in range(20):
    it()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    it()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
