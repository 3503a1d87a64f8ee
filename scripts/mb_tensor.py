"""ParTensor C3(ii) weight-streaming kernels alone (for ncu): 5 applies + 5 pullbacks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import parops
from parops import ComplexF32
g = torch.Generator(device="cuda").manual_seed(0)
ws = (20, 20, 16, 16, 16, 8)
Tn = parops.ParTensor(ComplexF32, tuple(range(1, 7)), ws, tuple(range(2, 7)), (20, 16, 16, 16, 8), (1,) + tuple(range(3, 7)), (20, 16, 16, 16, 8))
Tt = Tn(parops.gpu(parops.init(Tn, rng=2)))
x = torch.view_as_complex(torch.randn((1, Tn.Domain, 2), generator=g, device="cuda", dtype=torch.float32)).t()
yb = torch.view_as_complex(torch.randn((1, Tn.Range, 2), generator=g, device="cuda", dtype=torch.float32)).t()
for _ in range(5):
    y, back = parops.rrule(Tt, x)
    back(yb)
torch.cuda.synchronize()
print("ok")
