"""On-box: QRCP time of the five C2 rounds (idfact at 8192^2) -- total ms from the stage profile."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import torch, brapprox
from brapprox._binding import DeviceMatrix
from brapprox._frontend import idfact_device
ctx = brapprox.Context(0); dev = torch.device("cuda", 0)
torch.manual_seed(0)
n, r = 8192, 640
U, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device=dev))
V, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device=dev))
s = 10.0 ** (-12.0 * torch.arange(r, dtype=torch.float64, device=dev) / 500.0)
At = ((V * s) @ U.T).contiguous(); A = DeviceMatrix(At.data_ptr(), n, n, n, keep=At)
for w in range(2): idfact_device(A, rtol=1e-12, seed=w, ctx=ctx)
ctx.profile_enable(True)
for i in range(5): inf = idfact_device(A, rtol=1e-12, seed=i, ctx=ctx)
prof = ctx.profile_read()
print(json.dumps({"lib": os.environ.get("BRA_LIB", "default"), "k": int(inf.k), "qrcp_ms": prof["qrcp"][0] / 5, "gemm_ms": prof["gemm"][0] / 5}))
