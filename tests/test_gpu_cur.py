"""GPU parity of the CUR factors (src/cur.jl:85-109: C, R and the k x k core's svd! / eigen!) and of pheigfact with
clustered and +-lambda spectra (pheigorth!, src/pheig.jl:342-364) against the oracle.

Criteria: C and R bit-exact (gathers); core singular values |ds| <= 1e-10 s_1 (the reference's gesdd is absolutely
accurate; 1 ./ s is compared through s); V and U' orthonormal to 1e-12 sqrt(k); C U2 R within 2x of the oracle's error;
the reference's own test inequality (test/cur.jl)."""
import numpy as np
import pytest

import lra_oracle as o

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n,r,rtol", [(600, 500, 50, 1e-8), (300, 420, 30, 1e-10), (900, 700, 120, 1e-9)])
def test_cur_factors_match_oracle(ctx, m, n, r, rtol):
    import brapprox
    A = o.decaying_matrix(m, n, r, 11.0, r, seed=m + 7 * n)
    r1, r2 = o.RandomInputs(1), o.RandomInputs(2)
    rows, cols = o.curfact(A, o.LRAOptions(rtol=rtol), r1, r2)
    Ug = brapprox.curfact(A, brapprox.LRAOptions(rtol=rtol), rand=(r1.drawn, r2.drawn), ctx=ctx)
    np.testing.assert_array_equal(Ug.rows, rows)
    np.testing.assert_array_equal(Ug.cols, cols)
    F = brapprox.CUR(A, Ug, ctx=ctx)
    Co, (Vo, sinv_o, Uto), Ro = o.cur_core(A, rows, cols)
    k = len(cols)
    np.testing.assert_array_equal(F.C, Co)
    np.testing.assert_array_equal(F.R, Ro)
    so, sg = 1.0 / sinv_o, 1.0 / F.U.S
    assert np.all(np.diff(sg) <= 0)
    assert np.max(np.abs(sg - so)) <= 1e-10 * so[0]
    assert np.linalg.norm(F.U.U.T @ F.U.U - np.eye(k)) <= 1e-12 * np.sqrt(k)
    assert np.linalg.norm(F.U.Vt @ F.U.Vt.T - np.eye(k)) <= 1e-12 * np.sqrt(k)
    # the factored pseudo-inverse reproduces the core: W (V diag(1/s) U') W = W
    W = A[np.ix_(rows - 1, cols - 1)]
    pin_o = np.linalg.norm(W @ ((Vo * sinv_o) @ Uto) @ W - W)
    assert np.linalg.norm(W @ F.U.matrix() @ W - W) <= 10 * pin_o + 1e-13 * np.linalg.norm(W)
    nrm = np.linalg.norm(A, 2)
    eo = np.linalg.norm(A - Co @ ((Vo * sinv_o) @ Uto) @ Ro, 2) / nrm
    eg = np.linalg.norm(A - F.matrix(), 2) / nrm
    assert eg <= 2 * eo + 1e-15
    assert F["k"] == k and F["C"] is F.C


def test_cur_reference_inequality(ctx):
    """test/cur.jl:27-29 on its own matrices (the 128 x 64 Fourier matrix, real part; the symmetrised 64 x 64 block) with
    its own options (maxdet_tol = 0, sketch_randn_niter = 1): norm(A - Matrix(CUR(A, curfact(A)))) < 1000 rtol norm(A).
    rtol = 1e-8 instead of the reference's 5 eps: the CUR error grows like sigma_{k+1} / sigma_min(core), and at 5 eps the
    k x k core has sigma_min ~ eps sigma_1, so 1 ./ s amplifies rounding by 1e13 for ANY svd (the oracle's own gesdd
    gives 1e-3 on this matrix) -- whether the reference's bound holds there depends on its random x, y."""
    import brapprox
    rng = np.random.default_rng(0)
    M = o.matrixlib_fourier(rng.random(128), rng.random(64))
    rtol = 1e-8
    for A in (np.asfortranarray(M.real), np.asfortranarray((M[:64, :64] + M[:64, :64].T).real)):
        U = brapprox.curfact(A, rtol=rtol, maxdet_tol=0.0, sketch_randn_niter=1, seed=2, ctx=ctx)
        F = brapprox.CUR(A, U, ctx=ctx)
        assert np.linalg.norm(A - F.matrix()) < 1000 * rtol * np.linalg.norm(A)


def test_hermitian_cur_matches_oracle(ctx):
    import brapprox
    n, r = 500, 40
    G = o.decaying_matrix(n, n, r, 9.0, r, seed=77)
    A = np.asfortranarray(G @ G.T * np.sign(np.linspace(-1, 1, n))[None, :])
    A = np.asfortranarray((A + A.T) / 2)                    # symmetric indefinite
    rin = o.RandomInputs(3)
    rows, cols = o.curfact(A, o.LRAOptions(rtol=1e-8), rin)
    Ug = brapprox.curfact(A, brapprox.LRAOptions(rtol=1e-8), rand=(rin.drawn, None), ctx=ctx)
    assert Ug.hermitian
    np.testing.assert_array_equal(Ug.cols, cols)
    F = brapprox.CUR(A, Ug, ctx=ctx)
    Co, (winv_o, Xo) = o.hermcur_core(A, cols)
    k = len(cols)
    np.testing.assert_array_equal(F.C, Co)
    wo, wg = 1.0 / winv_o, 1.0 / F.U.values
    assert np.all(np.diff(wg) >= 0)
    assert np.max(np.abs(wg - wo)) <= 1e-10 * np.max(np.abs(wo))
    assert np.linalg.norm(F.U.vectors.T @ F.U.vectors - np.eye(k)) <= 1e-12 * np.sqrt(k)
    nrm = np.linalg.norm(A, 2)
    eo = np.linalg.norm(A - Co @ ((Xo * winv_o) @ Xo.T) @ Co.T, 2) / nrm
    eg = np.linalg.norm(A - F.matrix(), 2) / nrm
    assert eg <= 2 * eo + 1e-15


def test_pheigfact_plus_minus_pairs_and_clusters(ctx):
    """Eigenvalues of equal magnitude and opposite sign (the singular subspace mixes the two eigenvectors: resolved by the
    Rayleigh-Ritz step on the group) and an exactly repeated eigenvalue (pheigorth!'s cluster)."""
    import brapprox
    n = 300
    rng = np.random.default_rng(5)
    Q, _ = np.linalg.qr(rng.standard_normal((n, 12)))
    lam = np.array([3.0, -3.0, 2.0, 2.0, 2.0, -1.5, 1.5, 0.7, -0.2, 0.2, 0.05, -0.01])
    A = (Q * lam) @ Q.T
    A = np.asfortranarray((A + A.T) / 2)
    rin = o.RandomInputs(9)
    wo, Xo, Vo = o.pheigfact(A, o.LRAOptions(rtol=1e-10), rin)
    F = brapprox.pheigfact(A, brapprox.LRAOptions(rtol=1e-10), rand=rin.drawn, ctx=ctx)
    assert len(F.values) == len(wo) == 12
    assert np.max(np.abs(F.values - wo)) <= 1e-10 * 3.0
    assert np.max(np.abs(np.sort(F.values) - np.sort(lam))) <= 1e-10 * 3.0
    k = len(wo)
    assert np.linalg.norm(F.vectors.T @ F.vectors - np.eye(k)) <= 1e-10
    assert np.linalg.norm(A - F.matrix(), 2) <= 1e-9 * 3.0
    # invariant subspaces agree group by group (individual vectors inside a repeated eigenvalue are arbitrary)
    for v in np.unique(np.round(wo, 8)):
        io = np.abs(wo - v) < 1e-6
        ig = np.abs(F.values - v) < 1e-6
        Po, Pg = Xo[:, io] @ Xo[:, io].T, F.vectors[:, ig] @ F.vectors[:, ig].T
        assert np.linalg.norm(Po - Pg) <= 1e-8


def test_pheigorth_kernel_matches_oracle(ctx):
    """pheigorth! itself: a deliberately non-orthogonal pair inside a cluster is orthogonalised exactly like the
    reference's sweep (same order of operations)."""
    import brapprox
    import ctypes as C
    from brapprox import _binding as B
    if not hasattr(B.lib, "bra_debug_pheigorth"):
        pytest.skip("debug entry not built")
    rng = np.random.default_rng(1)
    rows, kk = 200, 9
    V = np.asfortranarray(np.linalg.qr(rng.standard_normal((rows, kk)))[0])
    V[:, 3] += 1e-3 * V[:, 2]
    V[:, 4] += 2e-3 * V[:, 2] - 1e-3 * V[:, 3]
    vals = np.array([-2.0, -1.0, 0.5, 0.5 * (1 + 1e-10), 0.5 * (1 + 2e-10), 0.9, 1.0, 1.0, 4.0])
    Vo = V.copy(order="F")
    o.pheigorth(vals, Vo, o.LRAOptions())
    Vg = V.copy(order="F")
    B.lib.bra_debug_pheigorth.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_double]
    ctx.check(B.lib.bra_debug_pheigorth(ctx.handle, C.c_void_p(vals.ctypes.data), C.c_void_p(Vg.ctypes.data), rows, rows, kk,
                                        float(np.sqrt(o.EPS))))
    assert np.max(np.abs(Vg - Vo)) <= 1e-15
    assert abs(Vg[:, 2] @ Vg[:, 3]) <= 1e-15 and abs(Vg[:, 6] @ Vg[:, 7]) <= 1e-15
