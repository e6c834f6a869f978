"""GPU parity of the pqrfact / psvdfact tails against the oracle on identical Omega.

Criteria (hard part 5):
  pqrfact : k, p identical; ||Q'Q - I|| <= 1e-13; Q, R equal to the oracle's Householder factors AFTER normalising
            the per-column sign (Cholesky-QR gives diag(R) > 0, LAPACK's is -sign(alpha)), weighted like tau in the
            stage-wise tests (column i of Q is determined to eps*|R11|/|R_ii|); ||A - QRP'|| within 2x of the oracle's.
  psvdfact: rank after psvdrank identical; |dsigma| <= 1e-10*sigma_1; U diag(S)/sigma_1 and diag(S) Vt/sigma_1
            entrywise within 1e-10 after sign-fixing; reconstruction error within 2x of the oracle's.
"""
import numpy as np
import pytest

import lra_oracle as o

pytestmark = pytest.mark.gpu


def _pair(fn_o, fn_g, A, kw, seed=0, **extra):
    rin = o.RandomInputs(seed)
    Fo = fn_o(A, o.LRAOptions(**kw), rin, **extra)
    return Fo, rin


@pytest.mark.parametrize("m,n,r,rtol,trans", [(1024, 768, 100, 1e-11, "n"), (600, 900, 64, 1e-9, "n"),
                                               (512, 400, 80, 1e-10, "c"), (1600, 1300, 420, 1e-11, "n")])
def test_pqrfact_matches_oracle(ctx, m, n, r, rtol, trans):
    import brapprox
    A = o.decaying_matrix(m, n, r, 13.0, r, seed=m + n)
    rin = o.RandomInputs(1)
    Fo = o.pqrfact(A, o.LRAOptions(rtol=rtol), rin, trans)
    Fg = brapprox.pqrfact(A, brapprox.LRAOptions(rtol=rtol), trans=trans, rand=rin.drawn, ctx=ctx)
    assert Fg.k == Fo.k
    np.testing.assert_array_equal(Fg.p, Fo.p)
    k = Fo.k
    assert np.linalg.norm(Fg.Q.T @ Fg.Q - np.eye(k)) <= 1e-13 * np.sqrt(k)
    Aop = A if trans == "n" else A.T
    nrm = np.linalg.norm(Aop)
    eo = np.linalg.norm(Aop - Fo.matrix()) / nrm
    eg = np.linalg.norm(Aop - Fg.matrix()) / nrm
    assert eg <= 2 * eo + 1e-15
    # factor parity modulo the Householder sign convention
    d = np.sign(np.diag(Fo.R[:, :k]))
    Qo, Ro = Fo.Q * d, Fo.R * d[:, None]
    r11 = abs(Ro[0, 0])
    w = np.abs(np.diag(Ro)) / r11
    assert np.max(np.abs(Fg.R - Ro)) <= 1e-10 * r11
    assert np.max(np.abs(Fg.Q - Qo) * w[None, :]) <= 1e-10


def _signfix(U, Vt, Uref):
    s = np.sign(np.sum(U * Uref, axis=0))
    s[s == 0] = 1.0
    return U * s, Vt * s[:, None]


@pytest.mark.parametrize("m,n,r,rtol", [(1024, 768, 100, 1e-11), (700, 500, 64, 1e-8), (400, 640, 80, 1e-10),
                                        (1600, 1300, 420, 1e-11), (300, 260, 20, 1e-7)])
def test_psvdfact_matches_oracle(ctx, m, n, r, rtol):
    import brapprox
    A = o.decaying_matrix(m, n, r, 13.0, r, seed=3 * m + n)
    rin = o.RandomInputs(2)
    So = o.psvdfact(A, o.LRAOptions(rtol=rtol), rin)
    Sg = brapprox.psvdfact(A, brapprox.LRAOptions(rtol=rtol), rand=rin.drawn, ctx=ctx)
    assert Sg.k_id == So.k_id
    assert len(Sg.S) == len(So.S)
    s1 = So.S[0]
    assert np.max(np.abs(Sg.S - So.S)) <= 1e-10 * s1
    assert np.all(np.diff(Sg.S) <= 0)
    kk = len(So.S)
    assert np.linalg.norm(Sg.U.T @ Sg.U - np.eye(kk)) <= 1e-12 * np.sqrt(kk)
    assert np.linalg.norm(Sg.Vt @ Sg.Vt.T - np.eye(kk)) <= 1e-9      # rows for tiny sigma are noise-limited
    Ug, Vtg = _signfix(Sg.U, Sg.Vt, So.U)
    assert np.max(np.abs(Ug * Sg.S - So.U * So.S)) <= 1e-10 * s1
    assert np.max(np.abs(Vtg * Sg.S[:, None] - So.Vt * So.S[:, None])) <= 1e-10 * s1
    nrm = np.linalg.norm(A, 2)
    eo = np.linalg.norm(A - So.matrix(), 2) / nrm
    eg = np.linalg.norm(A - Sg.matrix(), 2) / nrm
    assert eg <= 2 * eo + 1e-15


def test_psvdfact_reference_inequality(ctx):
    """test/psvd.jl:28-32 on the 128 x 64 Fourier (real part) matrix, fast mode (device Omega)."""
    import brapprox
    rng = np.random.default_rng(0)
    A = np.asfortranarray(o.matrixlib_fourier(rng.random(128), rng.random(64)).real)
    rtol = 5 * o.EPS
    F = brapprox.psvdfact(A, rtol=rtol, seed=3, ctx=ctx)
    assert np.linalg.norm(A - F.matrix()) < 100 * rtol * np.linalg.norm(A)
    s = brapprox.psvdvals(A, rank=len(F.S), rtol=0.0, seed=4, ctx=ctx)
    assert np.linalg.norm(s[:len(F.S)] - F.S) < 100 * rtol * np.linalg.norm(F.S)
    G = brapprox.pqrfact(A, rtol=rtol, seed=5, ctx=ctx)
    assert np.linalg.norm(A - G.matrix()) < 100 * rtol * np.linalg.norm(A)


def test_tail_kernels_standalone(ctx):
    """The building blocks through their effect: Hilbert-256 pqrfact (kappa(C) ~ 1e12) keeps Q orthonormal."""
    import brapprox
    A = o.matrixlib_hilb(256)
    F = brapprox.pqrfact(A, rtol=1e-12, seed=1, ctx=ctx)
    k = F.k
    assert 15 <= k <= 25
    assert np.linalg.norm(F.Q.T @ F.Q - np.eye(k)) <= 1e-13 * np.sqrt(k)
    assert np.all(np.diag(F.R[:, :k]) > 0)
    assert np.linalg.norm(A - F.matrix(), 2) <= 1e-10 * np.linalg.norm(A, 2)


def test_psvdfact_with_reference_test_options(ctx):
    """psvdfact on a maxdet-refined, power-iterated ID (the options of the reference's own suite, test/psvd.jl:8:
    LRAOptions(maxdet_tol=0., sketch_randn_niter=1)).  After maxdet the skeleton QR is preconditioned by a fresh sketch of
    A[:, sk] factored WITHOUT pivoting (R1 stays triangular in the reference's column order).  Against the oracle on identical Omega
    (a case whose swap sequence is well separated): same k, same p, same rank, |dsigma| <= 1e-10 sigma_1, U S and S Vt
    entrywise; plus the reference inequality on the 128 x 64 Fourier matrix in fast mode."""
    import brapprox
    A = o.decaying_matrix(300, 200, 60, 10.0, 60, seed=3)
    rin = o.RandomInputs(3)
    kw = dict(rtol=1e-9, maxdet_tol=0.0, sketch_randn_niter=1)
    So = o.psvdfact(A, o.LRAOptions(**kw), rin)
    Sg = brapprox.psvdfact(A, brapprox.LRAOptions(**kw), rand=rin.drawn, ctx=ctx)
    assert ctx.maxdet_swaps() > 0
    assert Sg.k_id == So.k_id and len(Sg.S) == len(So.S)
    s1 = So.S[0]
    kk = len(So.S)
    assert np.linalg.norm(Sg.U.T @ Sg.U - np.eye(kk)) <= 1e-12 * np.sqrt(kk)
    eo = np.linalg.norm(A - So.matrix(), 2) / np.linalg.norm(A, 2)
    eg = np.linalg.norm(A - Sg.matrix(), 2) / np.linalg.norm(A, 2)
    assert eg <= 2 * eo + 1e-15
    # sigma of two IDs with (possibly) different skeletons agree to the ID's own accuracy; with equal skeletons to 1e-10
    Vo = o.idfact(A, o.LRAOptions(**kw), o.RandomInputs(3))
    Vg = brapprox.idfact(A, rand=rin.drawn, ctx=ctx, **kw)
    same = np.array_equal(Vg.p, Vo.p)
    assert np.max(np.abs(Sg.S - So.S)) <= (1e-10 if same else 100 * 1e-9) * s1
    if same:
        Ug, Vtg = _signfix(Sg.U, Sg.Vt, So.U)
        assert np.max(np.abs(Ug * Sg.S - So.U * So.S)) <= 1e-10 * s1
        assert np.max(np.abs(Vtg * Sg.S[:, None] - So.Vt * So.S[:, None])) <= 1e-10 * s1
    # pqrfact on the same refined ID: Q orthonormal, R1 upper triangular in the reference's column order
    Fo = o.pqrfact(A, o.LRAOptions(**kw), o.RandomInputs(3))
    Fg = brapprox.pqrfact(A, brapprox.LRAOptions(**kw), rand=rin.drawn, ctx=ctx)
    k = Fo.k
    assert Fg.k == k
    assert np.linalg.norm(Fg.Q.T @ Fg.Q - np.eye(k)) <= 1e-13 * np.sqrt(k)
    assert np.all(np.diag(Fg.R[:, :k]) > 0) and np.max(np.abs(np.tril(Fg.R[:, :k], -1))) == 0.0
    nrm = np.linalg.norm(A)
    assert np.linalg.norm(A - Fg.matrix()) / nrm <= 2 * np.linalg.norm(A - Fo.matrix()) / nrm + 1e-15
    if np.array_equal(Fg.p, Fo.p):
        d = np.sign(np.diag(Fo.R[:, :k]))
        Ro = Fo.R * d[:, None]
        assert np.max(np.abs(Fg.R - Ro)) <= 1e-10 * abs(Ro[0, 0])
    rng = np.random.default_rng(0)
    F64 = np.asfortranarray(o.matrixlib_fourier(rng.random(128), rng.random(64)).real)
    rtol = 5 * o.EPS
    F = brapprox.psvdfact(F64, rtol=rtol, maxdet_tol=0.0, sketch_randn_niter=1, seed=3, ctx=ctx)
    assert np.linalg.norm(F64 - F.matrix()) < 100 * rtol * np.linalg.norm(F64)


@pytest.mark.parametrize("kind", ["psd", "indefinite", "hilbert"])
def test_pheigfact_matches_oracle(ctx, kind):
    """pheigfact (src/pheig.jl:276-296) on identical Omega: same ID rank, same number of eigenvalues after pheigrank,
    values within 1e-10 |lambda|_max, vectors entrywise (after sign-fixing, scaled by lambda / |lambda|_max),
    orthonormal vectors, reconstruction error within 2x of the oracle's."""
    import brapprox
    rng = np.random.default_rng(4)
    if kind == "hilbert":
        A, rtol = o.matrixlib_hilb(200), 1e-10
    else:
        n, r = 240, 40
        Qm, _ = np.linalg.qr(rng.standard_normal((n, r)))
        lam = 10.0 ** (-8.0 * np.arange(r) / r)
        if kind == "indefinite":
            lam = lam * np.where(np.arange(r) % 3 == 0, -0.7, 1.0)
        A = (Qm * lam) @ Qm.T
        A, rtol = np.asfortranarray((A + A.T) / 2), 1e-9
    rin = o.RandomInputs(6)
    wo, Xo, Vo = o.pheigfact(A, o.LRAOptions(rtol=rtol), rin)
    F = brapprox.pheigfact(A, rtol=rtol, rand=rin.drawn, ctx=ctx)
    assert F.k_id == Vo.k
    assert len(F.values) == len(wo)
    wmax = np.abs(wo).max()
    assert np.max(np.abs(F.values - wo)) <= 1e-10 * wmax
    kk = len(wo)
    assert np.linalg.norm(F.vectors.T @ F.vectors - np.eye(kk)) <= 1e-11 * np.sqrt(kk)
    sgn = np.sign(np.sum(F.vectors * Xo, axis=0))
    sgn[sgn == 0] = 1.0
    assert np.max(np.abs(F.vectors * sgn * F.values - Xo * wo)) <= 1e-9 * wmax
    nrm = np.linalg.norm(A, 2)
    eo = np.linalg.norm(A - (Xo * wo) @ Xo.T, 2) / nrm
    eg = np.linalg.norm(A - F.matrix(), 2) / nrm
    assert eg <= 2 * eo + 1e-12          # exact-rank inputs: both errors are rounding level (Q_z is orthonormal to ~1e-13)
    np.testing.assert_allclose(brapprox.pheigvals(A, rtol=rtol, rand=rin.drawn, ctx=ctx), F.values, rtol=0, atol=1e-13 * wmax)
    with pytest.raises(ValueError):
        brapprox.pheigfact(np.asfortranarray(rng.standard_normal((8, 8))), ctx=ctx)      # "matrix must be Hermitian"


def test_psvdfact_into_caller_buffers(ctx):
    """`out=`: the factors are written into caller-owned column-major buffers (the ABI's ownership model) and the
    returned arrays are views of them; results equal the allocating call's."""
    import brapprox
    A = o.decaying_matrix(320, 260, 50, 9.0, 50, seed=8)
    F0 = brapprox.psvdfact(A, rtol=1e-9, seed=4, ctx=ctx)
    Ub = np.zeros((320, 64), order="F")
    Sb = np.zeros(64)
    Vb = np.zeros((64, 260), order="F")
    F1 = brapprox.psvdfact(A, rtol=1e-9, seed=4, ctx=ctx, out=(Ub, Sb, Vb))
    kk = len(F0.S)
    assert len(F1.S) == kk and np.shares_memory(F1.U, Ub) and np.shares_memory(F1.Vt, Vb)
    np.testing.assert_array_equal(F1.S, F0.S)
    np.testing.assert_array_equal(F1.U, F0.U)
    np.testing.assert_array_equal(F1.Vt, F0.Vt)
    # pageable buffers are filled by bra_fetch after the call; PINNED buffers are registered with the library
    # (bra_psvd_set_outputs) and filled inside the call, each factor as soon as it exists; the registration is one shot
    assert brapprox.lib.bra_psvd_outputs_done(ctx.handle) == 0
    import torch
    Up = torch.zeros((64, 320), dtype=torch.float64, pin_memory=True).numpy().T          # 320 x 64, column-major
    Sp = torch.zeros((64,), dtype=torch.float64, pin_memory=True).numpy()
    Vp = torch.zeros((260, 64), dtype=torch.float64, pin_memory=True).numpy().T          # 64 x 260, column-major
    F2 = brapprox.psvdfact(A, rtol=1e-9, seed=4, ctx=ctx, out=(Up, Sp, Vp))
    assert brapprox.lib.bra_psvd_outputs_done(ctx.handle) == 7
    assert np.shares_memory(F2.U, Up) and np.shares_memory(F2.Vt, Vp)
    np.testing.assert_array_equal(F2.S, F0.S)
    np.testing.assert_array_equal(F2.U, F0.U)
    np.testing.assert_array_equal(F2.Vt, F0.Vt)
    brapprox.psvdfact(A, rtol=1e-9, seed=4, ctx=ctx)
    assert brapprox.lib.bra_psvd_outputs_done(ctx.handle) == 0
    # wide matrix (factors A'): U and Vt swap lanes
    Aw = np.asfortranarray(A.T)
    G0 = brapprox.psvdfact(Aw, rtol=1e-9, seed=4, ctx=ctx)
    Up2 = torch.zeros((64, 260), dtype=torch.float64, pin_memory=True).numpy().T
    Vp2 = torch.zeros((320, 64), dtype=torch.float64, pin_memory=True).numpy().T
    G1 = brapprox.psvdfact(Aw, rtol=1e-9, seed=4, ctx=ctx, out=(Up2, Sp, Vp2))
    assert brapprox.lib.bra_psvd_outputs_done(ctx.handle) == 7
    np.testing.assert_array_equal(G1.U, G0.U)
    np.testing.assert_array_equal(G1.Vt, G0.Vt)
    with pytest.raises(ValueError):
        brapprox.psvdfact(A, rtol=1e-9, seed=4, ctx=ctx, out=(np.zeros((320, 4), order="F"), Sb, Vb))


def _leverage_matrix(m, n, r, seed):
    """Rows of wildly different weight: a row-subset sketch misses most of the range, a sparse sketch has no
    oversampling -- the cases where the sketch's R11 is a poor preconditioner for A[:, sk]."""
    A = o.decaying_matrix(m, n, r, 10.0, r, seed=seed)
    w = 10.0 ** (-6.0 * np.random.default_rng(seed + 1).random(m))
    return np.asfortranarray(A * w[:, None])


@pytest.mark.parametrize("kind", ["sub", "sprn", "srft", "randn"])
def test_pqrfact_nonuniform_leverage_all_sketches(ctx, kind):
    """The reference's Householder qr! of A[:, sk] cannot fail (src/pqr.jl:297-305); neither may the preconditioned
    CholeskyQR2: k, p identical to the oracle, Q orthonormal, error within 2x."""
    import brapprox
    A = _leverage_matrix(900, 600, 60, 5)
    rin = o.RandomInputs(4)
    Fo = o.pqrfact(A, o.LRAOptions(rtol=1e-9, sketch=kind), rin)
    Fg = brapprox.pqrfact(A, brapprox.LRAOptions(rtol=1e-9, sketch=kind), rand=rin.drawn, ctx=ctx)
    assert Fg.k == Fo.k
    k = Fo.k
    same = int(np.flatnonzero(np.append(Fg.p[:k] != Fo.p[:k], True))[0])
    assert same >= k - 8          # pivots at the noise floor may flip (SRFT butterflies vs FFTW)
    assert np.linalg.norm(Fg.Q.T @ Fg.Q - np.eye(k)) <= 1e-13 * np.sqrt(k)
    nrm = np.linalg.norm(A)
    eo = np.linalg.norm(A - Fo.matrix()) / nrm
    eg = np.linalg.norm(A - Fg.matrix()) / nrm
    assert eg <= 2 * eo + 1e-15


@pytest.mark.parametrize("kind", ["sub", "sprn", "srft"])
def test_psvdfact_nonuniform_leverage_all_sketches(ctx, kind):
    import brapprox
    A = _leverage_matrix(800, 500, 50, 9)
    rin = o.RandomInputs(6)
    So = o.psvdfact(A, o.LRAOptions(rtol=1e-9, sketch=kind), rin)
    Sg = brapprox.psvdfact(A, brapprox.LRAOptions(rtol=1e-9, sketch=kind), rand=rin.drawn, ctx=ctx)
    assert Sg.k_id == So.k_id
    assert len(Sg.S) == len(So.S)
    assert np.max(np.abs(Sg.S - So.S)) <= 1e-10 * So.S[0]
    kk = len(So.S)
    assert np.linalg.norm(Sg.U.T @ Sg.U - np.eye(kk)) <= 1e-12 * np.sqrt(kk)
    nrm = np.linalg.norm(A, 2)
    eo = np.linalg.norm(A - So.matrix(), 2) / nrm
    eg = np.linalg.norm(A - Sg.matrix(), 2) / nrm
    assert eg <= 2 * eo + 1e-15


def test_skeleton_qr_retries_after_breakdown(ctx, monkeypatch):
    """With the a-priori choice switched off (BRA_SKELETON_NOFRESH) the :sub sketch's R11 preconditions A[:, sk]: whether
    or not the Gram Cholesky breaks down, the call must succeed (a breakdown is retried with a fresh preconditioner)."""
    import brapprox
    from brapprox import _binding as B
    monkeypatch.setenv("BRA_SKELETON_NOFRESH", "1")
    A = _leverage_matrix(900, 600, 60, 5)
    rin = o.RandomInputs(4)
    Fo = o.pqrfact(A, o.LRAOptions(rtol=1e-9, sketch="sub"), rin)
    B.lib.bra_debug_skeleton_retries.restype = int
    r0 = B.lib.bra_debug_skeleton_retries(ctx.handle)
    Fg = brapprox.pqrfact(A, brapprox.LRAOptions(rtol=1e-9, sketch="sub"), rand=rin.drawn, ctx=ctx)
    print("skeleton retries:", B.lib.bra_debug_skeleton_retries(ctx.handle) - r0)
    assert Fg.k == Fo.k
    assert np.linalg.norm(Fg.Q.T @ Fg.Q - np.eye(Fo.k)) <= 1e-12 * np.sqrt(Fo.k)
    nrm = np.linalg.norm(A)
    assert np.linalg.norm(A - Fg.matrix()) / nrm <= 2 * np.linalg.norm(A - Fo.matrix()) / nrm + 1e-15


def _spectrum_matrix(m, n, sig, seed):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((m, len(sig))))
    V, _ = np.linalg.qr(rng.standard_normal((n, len(sig))))
    return np.asfortranarray((U * sig) @ V.T)


@pytest.mark.parametrize("name", ["clusters", "steps", "one_gap", "flat_then_cliff"])
@pytest.mark.parametrize("noprecond", [False, True])
def test_psvdfact_core_on_non_geometric_spectra(ctx, monkeypatch, name, noprecond):
    """The k x k core SVD (Cholesky-of-Gram preconditioned one-sided Jacobi) on spectra that are NOT a
    geometric decay: repeated singular values, plateaus separated by big steps, a single 1e10 gap, a flat spectrum ending
    in a cliff -- with the preconditioner and (BRA_JACOBI_NOPRECOND) on the plain path.  Same criteria as the geometric
    cases: rank identical, |dsigma| <= 1e-10 sigma_1, U orthonormal, error within 2x of the oracle's."""
    import brapprox
    r = 120
    if name == "clusters":
        sig = np.repeat(10.0 ** -np.arange(0, 12, 1.0), 10)
    elif name == "steps":
        sig = np.concatenate([np.full(40, 1.0), np.full(40, 1e-4) * np.linspace(1, 0.5, 40), np.full(40, 1e-9) * np.linspace(1, 0.9, 40)])
    elif name == "one_gap":
        sig = np.concatenate([np.linspace(1.0, 0.5, 60), 1e-10 * np.linspace(1.0, 0.5, 60)])
    else:
        sig = np.concatenate([np.full(100, 1.0) * np.linspace(1.0, 0.99, 100), 10.0 ** -np.linspace(3, 13, 20)])
    assert len(sig) == r
    A = _spectrum_matrix(900, 700, sig, 17)
    if noprecond:
        monkeypatch.setenv("BRA_JACOBI_NOPRECOND", "1")
    rin = o.RandomInputs(3)
    So = o.psvdfact(A, o.LRAOptions(rtol=1e-12), rin)
    Sg = brapprox.psvdfact(A, brapprox.LRAOptions(rtol=1e-12), rand=rin.drawn, ctx=ctx)
    assert Sg.k_id == So.k_id
    assert len(Sg.S) == len(So.S)
    s1 = So.S[0]
    assert np.max(np.abs(Sg.S - So.S)) <= 1e-10 * s1
    kk = len(So.S)
    assert np.linalg.norm(Sg.U.T @ Sg.U - np.eye(kk)) <= 1e-12 * np.sqrt(kk)
    nrm = np.linalg.norm(A, 2)
    eo = np.linalg.norm(A - So.matrix(), 2) / nrm
    eg = np.linalg.norm(A - Sg.matrix(), 2) / nrm
    assert eg <= 2 * eo + 1e-15
    # the leading singular subspaces agree wherever the spectrum separates them (gap > 1e-3 sigma_1)
    true_sig = np.sort(sig)[::-1]
    cut = [i for i in range(1, min(kk, r)) if true_sig[i - 1] - true_sig[i] > 1e-3 * true_sig[i - 1] and true_sig[i - 1] > 1e-6]
    for c in cut[:3]:
        Pg, Po = Sg.U[:, :c] @ Sg.U[:, :c].T, So.U[:, :c] @ So.U[:, :c].T
        assert np.linalg.norm(Pg - Po, 2) <= 1e-6


def _graded(m, n, r, dec, seed):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((m, r)))
    V, _ = np.linalg.qr(rng.standard_normal((n, r)))
    s = 10.0 ** (-dec * np.arange(r) / r)
    return np.asfortranarray((U * s) @ V.T), s


def test_psvdfact_core_beyond_1024(ctx):
    """Cores of more than 1024 columns take the 6-column-block configuration of the Jacobi kernel (k <= 12 #SMs); the
    reference has no size limit (LAPACK gesdd).  Property test: rank, singular values against the construction,
    orthogonality, reconstruction."""
    import brapprox
    A, s = _graded(2300, 2100, 1500, 11.0, 3)
    F = brapprox.psvdfact(A, rtol=1e-10, seed=1, ctx=ctx)
    kk = len(F.S)
    assert F.k_id > 1024 and kk > 1024
    assert np.max(np.abs(F.S - s[:kk])) <= 1e-9
    assert np.linalg.norm(F.U.T @ F.U - np.eye(kk)) <= 1e-11
    assert np.linalg.norm(F.Vt @ F.Vt.T - np.eye(kk)) <= 1e-10
    assert np.linalg.norm(A - F.matrix(), 2) <= 5e-9


def test_pheigfact_core_beyond_1024(ctx):
    """The same for a core whose rotations are accumulated (pheigfact; blocks of 4 columns, 5 warps per pair,
    k <= 8 #SMs = 1184)."""
    import brapprox
    n, r = 2400, 1170
    rng = np.random.default_rng(5)
    V, _ = np.linalg.qr(rng.standard_normal((n, r)))
    lam = 10.0 ** (-10.0 * np.arange(r) / r) * np.where(np.arange(r) % 3 == 1, -1.0, 1.0)
    A = (V * lam) @ V.T
    A = np.asfortranarray(0.5 * (A + A.T))
    F = brapprox.pheigfact(A, rtol=1e-9, seed=2, ctx=ctx)
    kk = len(F.values)
    assert F.k_id > 1024 and kk > 1024
    assert np.linalg.norm(F.vectors.T @ F.vectors - np.eye(kk)) <= 1e-9
    assert np.linalg.norm(A - (F.vectors * F.values) @ F.vectors.T, 2) <= 1e-7


def test_psvdvals_skips_the_vectors(ctx):
    """psvdvals (src/psvd.jl:274-290) = the singular values of psvdfact on the same random inputs; the vectors are not
    formed (fetching them answers BRA_ERR_NOTREADY)."""
    import brapprox
    from brapprox import _binding as Bd
    for (m, n) in ((900, 700), (500, 820)):
        A = o.decaying_matrix(m, n, 90, 12.0, 90, seed=m)
        rin = o.RandomInputs(3)
        Fo = o.psvdfact(A, o.LRAOptions(rtol=1e-10), rin)
        F = brapprox.psvdfact(A, brapprox.LRAOptions(rtol=1e-10), rand=rin.drawn, ctx=ctx)
        s = brapprox.psvdvals(A, brapprox.LRAOptions(rtol=1e-10), rand=rin.drawn, ctx=ctx)
        assert len(s) == len(Fo.S)
        np.testing.assert_array_equal(s, F.S)
        assert np.max(np.abs(s - Fo.S)) <= 1e-10 * Fo.S[0]
        with pytest.raises(Bd.BraError):
            ctx.fetch(Bd.F_U, (m, len(s)))
    assert brapprox.psvdvals(A.astype(np.float32), seed=1, ctx=ctx).dtype == np.float32
