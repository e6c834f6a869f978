"""On-box stress of the psvd tail across core sizes (team-size / block-count variants of the Jacobi kernel, blocked
Cholesky and triangular-inverse levels): reconstruction error and orthogonality, repeated to catch ordering races."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import numpy as np, torch, brapprox
from brapprox._binding import DeviceMatrix
ctx = brapprox.Context(0)
dev = torch.device("cuda", 0)
bad = 0
for (m, n, r, dec) in [(700, 600, 40, 9.0), (1500, 1300, 160, 10.0), (2500, 2200, 360, 11.0), (4096, 3500, 800, 11.0),
                       (5000, 4000, 1300, 11.0), (5000, 4500, 1800, 11.0)]:
    g = torch.Generator(device=dev); g.manual_seed(m)
    U, _ = torch.linalg.qr(torch.randn(m, r, dtype=torch.float64, device=dev, generator=g))
    V, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device=dev, generator=g))
    s = 10.0 ** (-dec * torch.arange(r, dtype=torch.float64, device=dev) / r)
    A = ((U * s) @ V.T)
    Ah = np.asfortranarray(A.cpu().numpy())
    nrm = float(s[0])
    for rep in range(3):
        F = brapprox.psvdfact(Ah, rtol=1e-10, seed=rep, ctx=ctx)
        kk = len(F.S)
        err = np.linalg.norm(Ah - F.matrix(), 2) / nrm
        ou = np.linalg.norm(F.U.T @ F.U - np.eye(kk))
        ov = np.linalg.norm(F.Vt @ F.Vt.T - np.eye(kk))
        ds = np.max(np.abs(F.S - s[:kk].cpu().numpy())) / nrm
        ok = err < 1e-8 and ou < 1e-10 and ov < 1e-8 and ds < 1e-8
        bad += (not ok)
        print(f"m={m} n={n} r={r} rep={rep}: k_id={F.k_id} kk={kk} err={err:.2e} orthU={ou:.1e} orthV={ov:.1e} dS={ds:.1e} {'ok' if ok else 'BAD'}",
              flush=True)
print("FAILURES", bad)
