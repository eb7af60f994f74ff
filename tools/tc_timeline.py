"""Debug: per-stage clock64 timeline of one CTA of the tcgen05 Toeplitz GEMM (THZ_CZT_DEBUG=3)."""
import ctypes, os, sys
os.environ["THZ_CZT_DEBUG"] = sys.argv[1] if len(sys.argv) > 1 else "3"   # 3: last GEMM launched (GEMM 2), 4: GEMM 1
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quantizationawarethzdoe_b200 import _native as N
from quantizationawarethzdoe_b200.Props.CZT_Prop import CZT_prop
from quantizationawarethzdoe_b200.DataType.ElectricField import ElectricField

C, H, M = 2, 2048, 1024
wl = torch.linspace(0.4e-3, 0.8e-3, C)
x = torch.randn(1, C, H, H, dtype=torch.complex64, device="cuda")
f = ElectricField(data=x, wavelengths=wl, spacing=0.1e-3, device=torch.device("cuda"))
prop = CZT_prop(z_distance=0.5, device=torch.device("cuda"))
for _ in range(2):
    y = prop(f, M, M, 0.1e-3, 0.1e-3).data
torch.cuda.synchronize()
buf = np.zeros((3, 256, 4), dtype=np.int64)
rc = N.lib().thz_debug_tc_timeline(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
t0 = buf[0, 0, 0]
print("producer: stage  wait_data->  [empty wait]  [stores]  [fence+arrive]   stage period")
for kb in range(20, 36):
    p = buf[0, kb]
    print("  %3d  t=%7d  empty %5d  stores %5d  fence %5d  period %5d" % (kb, p[0] - t0, p[1] - p[0], p[2] - p[1], p[3] - p[2], buf[0, kb + 1, 0] - p[0]))
print("mma: stage  [full wait] [issue+commit]  period")
for kb in range(20, 36):
    m = buf[1, kb]
    print("  %3d  t=%7d  full %5d  issue %5d  period %5d" % (kb, m[0] - t0, m[1] - m[0], m[2] - m[1], buf[1, kb + 1, 0] - m[0]))
print("acc: chunk  [tfull wait] [drain]")
for c in range(8, 14):
    q = buf[2, c]
    print("  %3d  t=%7d  wait %5d  drain %5d" % (c, q[0] - t0, q[1] - q[0], q[2] - q[1]))
print("first stamps: prod %d mma %d ; last prod %d (total %d clk for 64 stages)" % (0, buf[1, 0, 0] - t0, buf[0, 63, 3] - t0, buf[0, 63, 3] - t0))
