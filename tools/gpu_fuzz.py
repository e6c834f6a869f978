"""Randomised cross-check of the fused factorizations against the oracle on identical random inputs (run on the GPU box):
shapes, ranks, tolerances, sketch kinds, transposes and rank caps drawn at random; reports every mismatch in k / p
(leading pivots) and every error ratio above 2x.  Usage: python tools/gpu_fuzz.py [cases] [seed] [maxdim] [maxrank]"""
import sys
import time

sys.path.insert(0, "oracle")
sys.path.insert(0, "lowrankapprox.jl_b200")
import numpy as np  # noqa: E402
import lra_oracle as o  # noqa: E402
import brapprox  # noqa: E402


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    maxdim = int(sys.argv[3]) if len(sys.argv) > 3 else 900
    maxrank = int(sys.argv[4]) if len(sys.argv) > 4 else 160
    rng = np.random.default_rng(seed)
    ctx = brapprox.Context(0)
    bad = 0
    t0 = time.time()
    for c in range(cases):
        m = int(rng.integers(1, maxdim))
        n = int(rng.integers(1, maxdim))
        r = int(rng.integers(1, min(m, n) + 1))
        r = min(r, maxrank)
        decades = float(rng.uniform(2.0, 14.0))
        rtol = float(10.0 ** rng.uniform(-12, -3))
        kind = str(rng.choice(["randn", "randn", "srft", "sprn", "sub", "none"]))
        trans = str(rng.choice(["n", "c"]))
        fn = str(rng.choice(["idfact", "pqrfact", "psvdfact"]))
        kw = dict(rtol=rtol, sketch=kind)
        if rng.random() < 0.25:
            kw["rank"] = int(rng.integers(1, r + 8))
            if rng.random() < 0.5:
                kw["sketchfact_adap"] = False
        if rng.random() < 0.2 and fn != "psvdfact":
            kw["maxdet_tol"] = 0.0
        if rng.random() < 0.2 and kind == "randn":
            kw["sketch_randn_niter"] = 1
        A = o.decaying_matrix(m, n, r, decades, r, seed=int(rng.integers(1 << 30)))
        tag = f"case {c}: {fn} {m}x{n} r={r} dec={decades:.1f} {kw} trans={trans}"
        try:
            rin = o.RandomInputs(c)
            if fn == "psvdfact":
                Fo = o.psvdfact(A, o.LRAOptions(**kw), rin)
                Fg = brapprox.psvdfact(A, brapprox.LRAOptions(**kw), rand=rin.drawn, ctx=ctx)
                ko, kg = len(Fo.S), len(Fg.S)
                nrm = Fo.S[0] if ko else 1.0
                eo = np.linalg.norm(A - (Fo.U * Fo.S) @ Fo.Vt, 2) / nrm if ko else 0.0
                eg = np.linalg.norm(A - (Fg.U * Fg.S) @ Fg.Vt, 2) / nrm if kg else 0.0
                ds = np.max(np.abs(Fg.S - Fo.S)) / nrm if ko == kg and ko else 0.0
                msg = f"k {kg}/{ko} err {eg:.2e}/{eo:.2e} dS {ds:.1e}"
                ok = ko == kg and eg <= 2 * eo + 1e-13 and ds <= 1e-10      # exact-rank inputs: rounding floor ~4e-14 vs LAPACK's 1e-14
            else:
                f_o = getattr(o, fn)
                f_g = getattr(brapprox, fn)
                Fo = f_o(A, o.LRAOptions(**kw), rin, trans)
                Fg = f_g(A, brapprox.LRAOptions(**kw), trans=trans, rand=rin.drawn, ctx=ctx)
                Aop = A if trans == "n" else A.T
                nrm = np.linalg.norm(Aop, 2)
                eo = np.linalg.norm(Aop - (Aop[:, Fo.sk - 1] @ Fo.matrix() if fn == "idfact" else Fo.matrix()), 2) / nrm
                eg = np.linalg.norm(Aop - (Aop[:, Fg.sk - 1] @ Fg.matrix() if fn == "idfact" else Fg.matrix()), 2) / nrm
                same_p = np.array_equal(Fg.p[:Fo.k], Fo.p[:Fo.k]) if Fg.k == Fo.k else False
                first = int(np.flatnonzero(Fg.p[:Fo.k] != Fo.p[:Fo.k])[0]) if (Fg.k == Fo.k and not same_p) else -1
                msg = f"k {Fg.k}/{Fo.k} p_equal {same_p} first_diff {first} err {eg:.2e}/{eo:.2e}"
                ok = Fg.k == Fo.k and eg <= 2 * eo + 1e-14 and (same_p or "maxdet_tol" in kw or first >= Fo.k - 4)
        except Exception as e:  # noqa: BLE001
            msg, ok = f"EXC {type(e).__name__}: {str(e)[:150]}", False
        if not ok:
            bad += 1
            print("MISMATCH", tag, "|", msg, flush=True)
    print(f"fuzz: {cases} cases, {bad} mismatches, {time.time() - t0:.1f}s")


if __name__ == "__main__":
    main()
