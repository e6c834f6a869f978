"""On-box timing of psvdvals (no Q, no singular vectors) against psvdfact on the C2 matrix.  Not a benchmark."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lowrankapprox.jl_b200"))
import numpy as np, torch, brapprox
from brapprox._binding import DeviceMatrix
from brapprox import _binding as B
ctx = brapprox.Context(0); dev = torch.device("cuda", 0)
n=8192; r=640
g = torch.Generator(device=dev); g.manual_seed(0)
U,_ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device=dev, generator=g))
V,_ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device=dev, generator=g))
s = 10.0 ** (-12.0 * torch.arange(r, dtype=torch.float64, device=dev) / 500)
At = ((U*s) @ V.T).T.contiguous()
A = DeviceMatrix(At.data_ptr(), n, n, n, keep=At)
torch.cuda.synchronize()
for fn in (brapprox.psvdvals, lambda A, **k: brapprox._frontend.psvdfact_device(A, **k)):
    for i in range(3): fn(A, rtol=1e-12, seed=i, ctx=ctx)
    torch.cuda.synchronize(); t0=time.time()
    for i in range(10): fn(A, rtol=1e-12, seed=i, ctx=ctx)
    torch.cuda.synchronize(); print((time.time()-t0)/10*1e3, "ms")
