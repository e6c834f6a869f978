import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, brapprox, lra_oracle as o
from brapprox import _binding as B
ctx = brapprox.Context(0)
A = o.decaying_matrix(300, 260, 30, 9.0, 30, seed=5)
rin = o.RandomInputs(2)
Fo = o.sketchfact(A, o.LRAOptions(rtol=1e-9, pqrfact_retval="t"), rin, "n")
V = brapprox.idfact(A, brapprox.LRAOptions(rtol=1e-9), rand=rin.drawn, ctx=ctx)
inf = ctx.info()
order, n = int(inf.orders[inf.rounds - 1]), int(inf.n)
steps = int(inf.steps[inf.rounds - 1])
Bg = ctx.fetch(B.F_BSKETCH, (order, n))
pg = ctx.fetch(B.F_P, (n,), dtype=np.int64)
Bo = o.apply_sketch("randn", A, order, rin.drawn[-1], "n")
B0 = Bo.copy(order="F")
p, tau_o, k_o = o.geqp3_adap(Bo, o.LRAOptions(rtol=1e-9))
k = k_o
print("V1" if os.environ.get("BRA_QRCP_V1") else "FAST", "steps", steps, "k", inf.k, k_o, "p[:k] equal", np.array_equal(pg[:k], p[:k]), "p equal", np.array_equal(pg, p))
inv_g, inv_o = np.argsort(pg), np.argsort(p)
D = np.abs(Bg[:k, inv_g] - Bo[:k, inv_o])
print("max diff per row:", " ".join(f"{x:.1e}" for x in D.max(axis=1)))
j = np.unravel_index(np.argmax(D), D.shape); print("argmax", j, "orig col", j[1], "pos_g", inv_g[j[1]], "pos_o", inv_o[j[1]], Bg[j[0], inv_g[j[1]]], Bo[j[0], inv_o[j[1]]])
# geqp3 directly on the same B through the stage-wise entry
pj, tau, Rg, kk, tr = brapprox.geqp3_adap(B0, rtol=1e-9, ctx=ctx)
print("stagewise k", kk, "steps", tr["steps"], "kb", tr.get("kb"))
