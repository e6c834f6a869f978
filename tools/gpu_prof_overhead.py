"""On-box: device time of psvdfact/idfact at C2 with the per-stage event profiling off vs on."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import torch, brapprox
from brapprox import _binding as B
from brapprox._binding import DeviceMatrix
from brapprox._frontend import psvdfact_device, idfact_device
ctx = brapprox.Context(0); dev = torch.device("cuda", 0)
torch.manual_seed(0)
n, r = 8192, 640
U, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device=dev))
V, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device=dev))
s = 10.0 ** (-12.0 * torch.arange(r, dtype=torch.float64, device=dev) / 500.0)
At = ((V * s) @ U.T).contiguous(); A = DeviceMatrix(At.data_ptr(), n, n, n, keep=At)
ext = torch.cuda.ExternalStream(int(B.lib.bra_stream(ctx.handle)), device=dev)
for name, fn in (("psvdfact", psvdfact_device), ("idfact", idfact_device)):
    for prof in (False, True, False):
        for w in range(2): fn(A, rtol=1e-12, seed=w, ctx=ctx)
        ctx.profile_enable(prof)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for i in range(10): fn(A, rtol=1e-12, seed=i, ctx=ctx)
        e1.record(ext); e1.synchronize()
        ctx.profile_read(); ctx.profile_enable(False)
        print(name, "profile", prof, "ms/step", e0.elapsed_time(e1) / 10, flush=True)
