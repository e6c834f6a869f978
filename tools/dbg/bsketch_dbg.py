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
print("rounds", inf.rounds, "order", order, "n", n, "steps", steps, "k", inf.k, "oracle k", Fo.k, "n drawn", len(rin.drawn))
Bg = ctx.fetch(B.F_BSKETCH, (order, n))
Bo = o.apply_sketch("randn", A, order, rin.drawn[-1], "n")
p, tau_o, k_o = o.geqp3_adap(Bo, o.LRAOptions(rtol=1e-9))
print("p equal:", np.array_equal(p, V.p), "k_o", k_o)
d = np.abs(np.triu(Bg[:k_o, :]) - np.triu(Bo[:k_o, :])).max(axis=0)
bad = np.nonzero(d > 1e-10)[0]
print("bad columns", bad[:40], len(bad))
print("p[bad]", p[bad[:20]], "V.p[bad]", V.p[bad[:20]])
