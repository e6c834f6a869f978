"""On-box stress of the pheigfact and CUR tails at core sizes beyond the test suite's (many panels of the fused
Cholesky / triangular-inverse kernels, several tiles per CTA): reconstruction error, orthogonality, repeated seeds."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
import numpy as np, torch, brapprox
ctx = brapprox.Context(0)
dev = torch.device("cuda", 0)
bad = 0
for (n, r, dec) in [(900, 60, 9.0), (2000, 300, 10.0), (3000, 620, 10.0), (4000, 1000, 10.0), (4500, 1180, 10.0)]:
    g = torch.Generator(device=dev); g.manual_seed(n)
    V, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device=dev, generator=g))
    s = 10.0 ** (-dec * torch.arange(r, dtype=torch.float64, device=dev) / r)
    sg = torch.where(torch.arange(r, device=dev) % 3 == 1, -1.0, 1.0).to(torch.float64)
    lam = s * sg
    A = (V * lam) @ V.T
    A = 0.5 * (A + A.T)
    Ah = np.asfortranarray(A.cpu().numpy())
    for rep in range(2):
        F = brapprox.pheigfact(Ah, rtol=1e-9, seed=rep, ctx=ctx)
        kk = len(F.values)
        R = (F.vectors * F.values) @ F.vectors.T
        err = np.linalg.norm(Ah - R, 2)
        ov = np.linalg.norm(F.vectors.T @ F.vectors - np.eye(kk))
        ok = err < 1e-7 and ov < 1e-8
        bad += (not ok)
        print(f"pheig n={n} r={r} rep={rep}: kk={kk} err={err:.2e} orth={ov:.1e} {'ok' if ok else 'BAD'}", flush=True)
for (m, n, r, dec) in [(1200, 1000, 50, 8.0), (2500, 2000, 250, 9.0), (4000, 3000, 600, 9.0), (5000, 4500, 1450, 9.0)]:
    g = torch.Generator(device=dev); g.manual_seed(m + 1)
    U, _ = torch.linalg.qr(torch.randn(m, r, dtype=torch.float64, device=dev, generator=g))
    V, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device=dev, generator=g))
    s = 10.0 ** (-dec * torch.arange(r, dtype=torch.float64, device=dev) / r)
    Ah = np.asfortranarray(((U * s) @ V.T).cpu().numpy())
    for rep in range(2):
        Uc = brapprox.curfact(Ah, rtol=1e-7, seed=rep, ctx=ctx)
        F = brapprox.CUR(Ah, Uc, ctx=ctx)
        k = len(Uc.rows)
        err = np.linalg.norm(Ah - F.matrix(), 2)
        ok = err < 1e-3
        bad += (not ok)
        print(f"cur m={m} n={n} r={r} rep={rep}: k={k} err={err:.2e} {'ok' if ok else 'BAD'}", flush=True)
print("FAILURES", bad)
