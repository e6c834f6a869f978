"""End-to-end GPU parity of idfact against the oracle on identical random inputs (same Omega):
k equal, p equal, C*T within 1e-10*||A|| entrywise (the attainable end-to-end factor check,
SURVEY.md section 7 hard part 3), spectral error within 2x of the oracle's."""
import numpy as np
import pytest

import lra_oracle as o

pytestmark = pytest.mark.gpu


def _run_pair(ctx, A, kw, seed=0, trans="n"):
    import brapprox
    rin = o.RandomInputs(seed)
    Vo = o.idfact(A, o.LRAOptions(**kw), rin, trans)
    Vg = brapprox.idfact(A, brapprox.LRAOptions(**kw), trans=trans, rand=rin.drawn, ctx=ctx)
    return Vo, Vg


def _check_pair(A, Vo, Vg, trans="n"):
    Aop = A if trans == "n" else A.T
    assert Vg.rounds == Vo.rounds
    assert Vg.k == Vo.k
    np.testing.assert_array_equal(Vg.p, Vo.p)
    C = Aop[:, Vo.sk - 1]
    nrm = np.linalg.norm(Aop, 2)
    assert np.max(np.abs(C @ Vg.T - C @ Vo.T)) <= 1e-10 * nrm
    eo = o.id_error(A, Vo, trans)
    eg = o.id_error(A, Vg, trans)
    assert eg <= 2 * eo + 1e-15


def test_idfact_hilbert_1024(ctx):
    """BASELINE config 1 (README.md:107,122 known answers)."""
    A = o.matrixlib_hilb(1024)
    Vo, Vg = _run_pair(ctx, A, dict(rtol=1e-12))
    assert Vo.k == 22
    _check_pair(A, Vo, Vg)
    # rtol = 1e-15 terminates inside rounding noise: k within 1, pivots equal on the above-noise prefix
    Vo, Vg = _run_pair(ctx, A, dict(rtol=1e-15))
    assert abs(Vg.k - Vo.k) <= 1 and Vg.k in (26, 27, 28)
    assert o.id_error(A, Vg) <= 2 * o.id_error(A, Vo) + 1e-15


@pytest.mark.parametrize("m,n,r,rtol", [(1024, 1024, 120, 1e-12), (2048, 1536, 300, 1e-10), (700, 900, 64, 1e-8)])
def test_idfact_decaying(ctx, m, n, r, rtol):
    A = o.decaying_matrix(m, n, r, 14.0, r, seed=m + n)
    Vo, Vg = _run_pair(ctx, A, dict(rtol=rtol))
    _check_pair(A, Vo, Vg)


def test_idfact_trans_c(ctx):
    A = o.decaying_matrix(600, 400, 80, 13.0, 80, seed=4)
    Vo, Vg = _run_pair(ctx, A, dict(rtol=1e-11), trans="c")
    _check_pair(A, Vo, Vg, trans="c")


def test_idfact_rank_and_nonadaptive(ctx):
    A = o.decaying_matrix(800, 800, 200, 10.0, 200, seed=8)
    for kw in (dict(rank=40), dict(rank=40, sketchfact_adap=False), dict(rank=0), dict(nb=16, rtol=1e-9)):
        Vo, Vg = _run_pair(ctx, A, kw)
        assert Vg.rounds == Vo.rounds and Vg.k == Vo.k
        np.testing.assert_array_equal(Vg.p, Vo.p)


def test_idfact_fast_mode_device_rng(ctx):
    """Fast mode (device Philox Omega): no oracle twin, so check the reference's own inequality
    (test/id.jl:27-32 style) and determinism."""
    import brapprox
    A = o.decaying_matrix(1500, 1200, 150, 14.0, 150, seed=2)
    V1 = brapprox.idfact(A, rtol=1e-12, seed=5, ctx=ctx)
    V2 = brapprox.idfact(A, rtol=1e-12, seed=5, ctx=ctx)
    np.testing.assert_array_equal(V1.p, V2.p)
    np.testing.assert_array_equal(V1.T, V2.T)
    err = np.linalg.norm(A - A[:, V1.sk - 1] @ V1.matrix()) / np.linalg.norm(A)
    assert err < 100 * 1e-12


def _golden(pattern):
    import glob
    import os
    return sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", pattern)))


@pytest.mark.parametrize("path", _golden("id_*.npz"))
def test_idfact_golden(ctx, path):
    """The committed fixtures (oracle outputs from the real LAPACK) replayed through the CUDA path."""
    import brapprox
    z = np.load(path, allow_pickle=True)
    kw = dict(z["opts"].tolist())
    rand = [{k[len(f"rand{t}_"):]: z[k] for k in z.files if k.startswith(f"rand{t}_")} for t in range(int(z["n_rand"]))]
    trans = str(z["trans"])
    V = brapprox.idfact(z["A"], brapprox.LRAOptions(**kw), trans=trans, rand=rand, ctx=ctx)
    np.testing.assert_array_equal(V.sk, z["sk"])
    np.testing.assert_array_equal(V.rd, z["rd"])
    Aop = z["A"] if trans == "n" else z["A"].T
    C = Aop[:, V.sk - 1]
    assert np.max(np.abs(C @ V.T - C @ z["T"])) <= 1e-10 * np.linalg.norm(Aop, 2)


@pytest.mark.parametrize("path", _golden("qrcp_*.npz"))
def test_qrcp_golden(ctx, path):
    import brapprox
    z = np.load(path, allow_pickle=True)
    kw = dict(z["opts"].tolist())
    Bg, pg, taug, kg, trg = brapprox.geqp3_adap(z["B0"], brapprox.LRAOptions(**kw), ctx=ctx)
    assert kg == int(z["k"]) and trg["kb"] == z["kb"].tolist()
    np.testing.assert_array_equal(pg, z["p"])
    ns = int(z["steps"])
    assert np.max(np.abs(np.triu(Bg[:ns]) - z["R"])) <= 1e-13 * abs(z["R"][0, 0])


def test_errors_mirror_reference(ctx):
    import brapprox
    A = np.zeros((8, 8))
    with pytest.raises(ValueError):
        brapprox.idfact(A, rtol=-1.0, ctx=ctx)          # ArgumentError("rtol")
    with pytest.raises(ValueError):
        brapprox.idfact(A, sketch="bogus", ctx=ctx)     # ArgumentError("sketch")
    with pytest.raises(ValueError):
        brapprox.idfact(A, trans="x", ctx=ctx)          # ArgumentError("trans")


@pytest.mark.parametrize("m,n,r,rtol,tol", [(300, 200, 60, 1e-9, 0.0), (512, 640, 100, 1e-10, 0.0), (256, 384, 40, 1e-8, 0.25)])
def test_idfact_maxdet_matches_oracle(ctx, m, n, r, rtol, tol):
    """Strong-RRQR post-processing (maxdet_swapcols!, src/pqr.jl:444-501).
    Stage-wise (the parity statement): the SAME (p, T) in => the same swap sequence, p and T out as the oracle's
    restatement.  End to end T carries a forward sensitivity of eps/rtol (hard part 3), so the
    arg-max sequence of two correct implementations may part ways at near-ties of |T|; there the invariants are
    checked: k equal, max|T| <= 1 + tol, reconstruction error within 2x of the oracle's."""
    import brapprox
    A = o.decaying_matrix(m, n, r, 10.0, r, seed=m + 7 * n)
    rin = o.RandomInputs(3)
    Vo = o.idfact(A, o.LRAOptions(rtol=rtol, maxdet_tol=tol), rin)
    V0 = brapprox.idfact(A, rtol=rtol, rand=rin.drawn, ctx=ctx)                       # GPU, no maxdet
    Vg = brapprox.idfact(A, rtol=rtol, maxdet_tol=tol, rand=rin.drawn, ctx=ctx)       # GPU, maxdet
    assert Vg.k == Vo.k == V0.k
    assert np.abs(V0.T).max() > 1 + tol, "the case must actually trigger swaps"
    assert np.abs(Vg.T).max() <= 1 + tol + 1e-12
    # stage-wise: the oracle's swap loop on the GPU's own pre-maxdet (p, T)
    p1, T1 = V0.p.copy(), np.array(V0.T, order="F", copy=True)
    nsw = o.maxdet_swapcols(None, p1, T1, o.LRAOptions(maxdet_tol=tol), False)
    assert nsw == ctx.maxdet_swaps() and nsw > 0
    np.testing.assert_array_equal(Vg.p, p1)
    assert np.max(np.abs(Vg.T - T1)) <= 1e-12 * max(1.0, np.abs(T1).max())
    # end to end
    assert o.id_error(A, Vg) <= 2 * o.id_error(A, Vo) + 1e-15
    assert o.id_error(A, Vg) <= 2 * o.id_error(A, V0) + 1e-15


def test_idfact_maxdet_niter_limit(ctx):
    """maxdet_niter caps the number of swaps (src/pqr.jl:456-461)."""
    import brapprox
    A = o.decaying_matrix(300, 200, 60, 10.0, 60, seed=3)
    rin = o.RandomInputs(0)
    Vo = o.idfact(A, o.LRAOptions(rtol=1e-9, maxdet_tol=0.0, maxdet_niter=2), rin)
    Vg = brapprox.idfact(A, rtol=1e-9, maxdet_tol=0.0, maxdet_niter=2, rand=rin.drawn, ctx=ctx)
    np.testing.assert_array_equal(Vg.p, Vo.p)
    assert ctx.maxdet_swaps() == 2


@pytest.mark.parametrize("case", ["decay_n", "decay_c", "symmetric", "rank_deficient", "niter2"])
def test_idfact_power_iteration_matches_oracle(ctx, case):
    """sketch_randn_niter > 0 (src/sketch.jl:140-149, 163-172): Omega op(A), then rounds of row orthonormalisation and
    multiplication by op(A)' and op(A).  The reference orthonormalises with LAPACK's LQ; the device builds another
    orthonormal basis of the same row space, which leaves the pivots, R (up to row signs) and T of the following QRCP
    unchanged up to rounding.  On identical Omega: rounds, k and p equal the oracle's, C*T within 1e-10 ||A||, error
    within 2x; power iteration must not make the ID worse than niter = 0."""
    import brapprox
    niter, trans, rtol = 1, "n", 1e-9
    if case == "decay_n":
        A = o.decaying_matrix(300, 200, 60, 10.0, 60, seed=3)
    elif case == "decay_c":
        A, trans = o.decaying_matrix(220, 320, 70, 10.0, 70, seed=5), "c"
    elif case == "symmetric":
        A, rtol = o.matrixlib_hilb(256), 1e-10
    elif case == "rank_deficient":
        rng = np.random.default_rng(7)
        A = np.asfortranarray(rng.standard_normal((260, 20)) @ rng.standard_normal((20, 180)))   # rank 20 < order 40
    else:
        A, niter = o.decaying_matrix(400, 300, 90, 11.0, 90, seed=9), 2
    rin = o.RandomInputs(4)
    Vo = o.idfact(A, o.LRAOptions(rtol=rtol, sketch_randn_niter=niter), rin, trans)
    Vg = brapprox.idfact(A, rtol=rtol, sketch_randn_niter=niter, trans=trans, rand=rin.drawn, ctx=ctx)
    V0 = brapprox.idfact(A, rtol=rtol, trans=trans, rand=rin.drawn, ctx=ctx)
    Aop = A if trans == "n" else np.asfortranarray(A.T)
    assert Vg.rounds == Vo.rounds
    assert Vg.k == Vo.k
    np.testing.assert_array_equal(Vg.p[:Vg.k], Vo.p[:Vo.k])            # the skeleton, in pivot order
    C = Aop[:, Vo.sk - 1]
    nrm = np.linalg.norm(Aop, 2)
    if case != "rank_deficient":        # beyond the numerical rank the order of p is rounding noise
        np.testing.assert_array_equal(Vg.p, Vo.p)
        assert np.max(np.abs(C @ Vg.T - C @ Vo.T)) <= 1e-10 * nrm
    eg, eo, e0 = o.id_error(Aop, Vg), o.id_error(Aop, Vo), o.id_error(Aop, V0)
    assert eg <= 2 * eo + 1e-14
    assert eg <= 2 * e0 + 1e-14


@pytest.mark.parametrize("sketch", ["randn", "sub", "srft", "sprn"])
def test_reference_id_test_with_its_own_options(ctx, sketch):
    """test/id.jl:4-32 with the options the reference's suite actually uses -- LRAOptions(maxdet_tol=0.,
    sketch_randn_niter=1), rtol = 5 eps -- on the 128 x 64 Fourier (real part) matrix, both transposes, Float64:
    ||A - A[:,sk] [I T] P'|| < 100 rtol ||A||.  (sketch = :none and the other element types are not built.)"""
    import brapprox
    rng = np.random.default_rng(11)
    A = np.asfortranarray(o.matrixlib_fourier(rng.random(128), rng.random(64)).real)
    rtol = 5 * o.EPS
    nrm = np.linalg.norm(A)
    for trans in ("n", "c"):
        V = brapprox.idfact(A, rtol=rtol, sketch=sketch, maxdet_tol=0.0, sketch_randn_niter=1, trans=trans, seed=5, ctx=ctx)
        Aop = A if trans == "n" else A.T
        assert np.linalg.norm(Aop - Aop[:, V.sk - 1] @ V.matrix()) < 100 * rtol * nrm      # ID(:n, A, V)
        if V.k < Aop.shape[1]:
            assert np.abs(V.T).max() <= 1.0 + 1e-12


@pytest.mark.parametrize("m,n,r,rtol,trans", [(300, 200, 60, 1e-9, "n"), (700, 150, 50, 1e-10, "n"), (180, 400, 40, 1e-8, "c"),
                                               (128, 64, 0, 5 * 2.220446049250313e-16, "n")])
def test_sketch_none_matches_oracle(ctx, m, n, r, rtol, trans):
    """sketch = :none (pqrfact_none, src/pqr.jl:323-327): the early-terminating QRCP on op(A) itself -- the real dlaqps
    through the oracle is the comparison.  k and p identical, C*T within 1e-10 ||A||, error within 2x; r = 0 is the
    reference's 128 x 64 Fourier test matrix with its inequality (test/id.jl:27-32)."""
    import brapprox
    if r == 0:
        rng = np.random.default_rng(5)
        A = np.asfortranarray(o.matrixlib_fourier(rng.random(m), rng.random(n)).real)
    else:
        A = o.decaying_matrix(m, n, r, 10.0, r, seed=m + n)
    Vo = o.idfact(A, o.LRAOptions(rtol=rtol, sketch="none"), None, trans)
    Vg = brapprox.idfact(A, rtol=rtol, sketch="none", trans=trans, ctx=ctx)
    Aop = A if trans == "n" else np.asfortranarray(A.T)
    assert Vg.k == Vo.k
    if r == 0:
        assert np.linalg.norm(Aop - Aop[:, Vg.sk - 1] @ Vg.matrix()) < 100 * rtol * np.linalg.norm(Aop)
        return
    # the generators have exact rank r: pivots taken after step r (the block runs on to its end) are rounding noise,
    # so the skeleton is compared in order and the rest as a set, T through the permutation
    np.testing.assert_array_equal(Vg.p[:Vg.k], Vo.p[:Vo.k])
    assert sorted(Vg.p) == sorted(Vo.p)
    C = Aop[:, Vo.sk - 1]
    Tg = np.zeros((Vg.k, Aop.shape[1]))
    Tg[:, Vg.rd - 1] = Vg.T
    To = np.zeros((Vo.k, Aop.shape[1]))
    To[:, Vo.rd - 1] = Vo.T
    assert np.max(np.abs(C @ Tg - C @ To)) <= 1e-10 * np.linalg.norm(Aop, 2)
    assert o.id_error(Aop, Vg) <= 2 * o.id_error(Aop, Vo) + 1e-15
    if trans == "n":
        F = brapprox.psvdfact(A, rtol=rtol, sketch="none", ctx=ctx)
        So = o.psvdfact(A, o.LRAOptions(rtol=rtol, sketch="none"), None)
        assert len(F.S) == len(So.S)
        assert np.max(np.abs(F.S - So.S)) <= 1e-10 * So.S[0]


@pytest.mark.parametrize("m,n", [(300, 200), (180, 260), (200, 200)])
def test_curfact_matches_oracle(ctx, m, n):
    """curfact (src/cur.jl:532-566): the reference's two-pass row/column selection with both passes on the device;
    identical random inputs => identical index sets; the CUR reconstruction C pinv(C) A pinv(R) R is as accurate as the
    oracle's.  The square case is symmetric and takes the Hermitian branch."""
    import brapprox
    A = o.decaying_matrix(m, n, 50, 9.0, 50, seed=2 * m + n)
    if m == n:
        A = np.asfortranarray(A @ A.T)
    r1, r2 = o.RandomInputs(1), o.RandomInputs(2)
    ro, co = o.curfact(A, o.LRAOptions(rtol=1e-8), r1, r2)
    U = brapprox.curfact(A, rtol=1e-8, rand=(r1.drawn, r2.drawn), ctx=ctx)
    np.testing.assert_array_equal(U.rows, ro)
    np.testing.assert_array_equal(U.cols, co)
    C, R = A[:, U.cols - 1], A[U.rows - 1, :]
    err = np.linalg.norm(A - C @ (np.linalg.pinv(C) @ A @ np.linalg.pinv(R)) @ R, 2) / np.linalg.norm(A, 2)
    assert err < 1e-6
    rows, cols = brapprox.cur(A, rtol=1e-8, seed=3, ctx=ctx)
    assert len(rows) == len(cols) > 0


@pytest.mark.parametrize("m,n", [(4096, 2304), (4097, 2100)])
def test_pipelined_upload_equals_plain_upload(ctx, m, n):
    """A large host-resident A is uploaded in column panels while the sketches of the first adaptive rounds are formed on
    the panels already on the device (bra_stage_A).  Same Philox Omega, same rounds, same k and p as the plain upload;
    T to rounding (the stacked product splits the contraction differently)."""
    import os
    import brapprox
    A = o.decaying_matrix(m, n, 120, 11.0, 120, seed=12)     # > 64 MB: pipelined; odd m: padded device ld, 2-D panel copies
    os.environ["BRA_NO_SPEC_UPLOAD"] = "1"
    try:
        V0 = brapprox.idfact(A, rtol=1e-10, seed=9, ctx=ctx)
    finally:
        del os.environ["BRA_NO_SPEC_UPLOAD"]
    V1 = brapprox.idfact(A, rtol=1e-10, seed=9, ctx=ctx)
    assert V1.rounds == V0.rounds and V1.k == V0.k
    np.testing.assert_array_equal(V1.p[:V1.k], V0.p[:V0.k])
    C = A[:, V0.sk - 1]
    T0 = np.zeros((V0.k, A.shape[1])); T0[:, V0.rd - 1] = V0.T
    T1 = np.zeros((V1.k, A.shape[1])); T1[:, V1.rd - 1] = V1.T
    assert np.max(np.abs(C @ T1 - C @ T0)) <= 1e-10 * np.linalg.norm(A, 2)
    F = brapprox.psvdfact(A, rtol=1e-10, seed=9, ctx=ctx)
    assert np.linalg.norm(A - F.matrix(), 2) <= 1e-8 * np.linalg.norm(A, 2)


def test_nested_sketch_rounds(ctx, monkeypatch):
    """Fast mode: the rounds of the adaptive loop share their Omega rows (round t = round t-1's rows + fresh ones), so the
    library multiplies max(order) rows with A where the reference schedule multiplies sum(order); BRA_SKETCH_FRESH=1
    restores independent draws per round.  Both satisfy the reference's error inequality with the same round structure;
    caller-supplied Omegas (parity mode) are multiplied round by round as given."""
    import brapprox
    A = o.decaying_matrix(1400, 1100, 200, 12.0, 200, seed=21)
    nrm = np.linalg.norm(A, 2)
    V = brapprox.idfact(A, rtol=1e-10, seed=3, ctx=ctx)
    orders = [r[0] for r in V.rounds]
    assert len(orders) >= 3
    assert brapprox.lib.bra_debug_sketch_rows(ctx.handle) == max(orders)
    assert np.linalg.norm(A - A[:, V.sk - 1] @ V.matrix(), 2) <= 1e3 * 1e-10 * nrm
    W = brapprox.idfact(A, rtol=1e-10, seed=3, sketch_fresh=True, ctx=ctx)            # BRA_OPT_FRESH_SKETCH
    monkeypatch.setenv("BRA_SKETCH_FRESH", "1")
    W2 = brapprox.idfact(A, rtol=1e-10, seed=3, ctx=ctx)                              # the same through the environment
    monkeypatch.delenv("BRA_SKETCH_FRESH")
    np.testing.assert_array_equal(W2.sk, W.sk)
    assert [r[0] for r in W.rounds] == orders
    assert brapprox.lib.bra_debug_sketch_rows(ctx.handle) == sum(orders)
    assert np.linalg.norm(A - A[:, W.sk - 1] @ W.matrix(), 2) <= 1e3 * 1e-10 * nrm
    assert abs(len(W.sk) - len(V.sk)) <= 2
    # the first round is the same in both (same stream, same rows): same k in round 1
    assert V.rounds[0] == W.rounds[0]
    # parity mode: every round's Omega is the caller's
    rin = o.RandomInputs(6)
    Vo = o.idfact(A, o.LRAOptions(rtol=1e-10), rin)
    Vg = brapprox.idfact(A, brapprox.LRAOptions(rtol=1e-10), rand=rin.drawn, ctx=ctx)
    np.testing.assert_array_equal(Vg.sk, Vo.sk)
    assert brapprox.lib.bra_debug_sketch_rows(ctx.handle) == sum(r[0] for r in Vg.rounds)
