"""Stage-wise GPU parity ("teacher forcing"): every CUDA kernel is fed the ORACLE's input for its
stage and compared with the oracle's output for that stage (parity plan).

Tolerances (FP64):
  sketch   B = Omega*A          : 1e-13 normwise (different summation order than OpenBLAS dgemm)
  QRCP     (k, p, kb) exact     ; R, tau to 1e-13 * |R11|
  T-solve  given identical R    : 1e-10 entrywise relative to max|T|
"""
import numpy as np
import pytest

import lra_oracle as o

pytestmark = pytest.mark.gpu


def _decay(m, n, r, decades, seed):
    return o.decaying_matrix(m, n, r, decades, r, seed)


@pytest.mark.parametrize("m,n,order", [(1024, 1024, 40), (1000, 777, 72), (512, 2048, 136), (2048, 512, 264),
                                       (333, 129, 40), (64, 40, 24)])
def test_sketch_randn_matches_dgemm(ctx, m, n, order):
    import brapprox
    rng = np.random.default_rng(m + n + order)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    Om = np.asfortranarray(rng.standard_normal((order, m)))
    Bg = brapprox.sketch(A, order, rand={"Omega": Om}, ctx=ctx)
    Bo = o.sketch_randn(A, Om)
    assert Bg.shape == (order, n)
    assert np.linalg.norm(Bg - Bo) <= 1e-13 * np.linalg.norm(Bo)


@pytest.mark.parametrize("m,n,order", [(300, 500, 40), (129, 64, 72)])
def test_sketch_randn_trans_c(ctx, m, n, order):
    import brapprox
    rng = np.random.default_rng(7)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    Om = np.asfortranarray(rng.standard_normal((order, n)))
    Bg = brapprox.sketch(A, order, trans="c", rand={"Omega": Om}, ctx=ctx)
    Bo = o.sketch_randn(A, Om, "c")
    assert Bg.shape == (order, m)
    assert np.linalg.norm(Bg - Bo) <= 1e-13 * np.linalg.norm(Bo)


def _qrcp_case(ctx, B0, opts_kw, check_tail=True):
    import brapprox
    oo = o.LRAOptions(**opts_kw)
    Bo = B0.copy(order="F")
    tr = o.QRCPTrace()
    po, tauo, ko = o.geqp3_adap(Bo, oo, o.dlaqps_real, tr)
    Bg, pg, taug, kg, trg = brapprox.geqp3_adap(B0, brapprox.LRAOptions(**opts_kw), ctx=ctx)
    assert kg == ko, (kg, ko, tr.kb, trg["kb"])
    assert trg["steps"] == tr.steps
    assert trg["kb"] == tr.kb
    np.testing.assert_array_equal(pg, po)
    ns = tr.steps
    if ns == 0:
        return
    r11 = max(abs(Bo[0, 0]), 1e-300)
    Rg, Ro = np.triu(Bg[:ns, :]), np.triu(Bo[:ns, :])
    assert np.max(np.abs(Rg - Ro)) <= 1e-13 * r11
    # tau_i and the reflector v_i are functions of the DIRECTION of the i-th residual column, which is
    # only determined to ~eps*|R11|/|R_ii|: weight the comparison by |R_ii|/|R11| (exact for i = 1).
    w = np.abs(np.diag(Ro)[:ns]) / r11
    assert np.max(np.abs(taug[:ns] - tauo[:ns]) * w) <= 1e-12
    if check_tail:
        # reflectors below the diagonal (weighted as above) and the trailing matrix, LAPACK layout
        D = np.abs(Bg - Bo)
        for i in range(ns):
            D[i + 1:, i] *= w[i]
        assert np.max(D) <= 1e-12 * max(r11, 1.0)


@pytest.mark.parametrize("l,n,decades,rtol", [(40, 1024, 14, 1e-12), (72, 777, 6, 1e-12), (136, 2048, 20, 1e-10),
                                              (264, 1500, 30, 1e-12), (40, 33, 3, 1e-8), (24, 64, 2, 1e-3)])
def test_qrcp_matches_dlaqps(ctx, l, n, decades, rtol):
    rng = np.random.default_rng(l * 7 + n)
    A = _decay(max(n, 300), n, min(n, 200), decades, 3)
    B0 = o.sketch_randn(A, np.asfortranarray(rng.standard_normal((l, A.shape[0]))))
    _qrcp_case(ctx, B0, dict(rtol=rtol))


def test_qrcp_rank_cap_and_nb(ctx):
    rng = np.random.default_rng(5)
    B0 = np.asfortranarray(rng.standard_normal((72, 400)))
    for kw in (dict(rank=10), dict(rank=0), dict(rank=1), dict(nb=8, rank=50), dict(nb=1, rank=20),
               dict(nb=64), dict(rtol=0.5)):
        _qrcp_case(ctx, B0, kw)


def test_qrcp_zero_and_duplicate_columns(ctx):
    rng = np.random.default_rng(9)
    B0 = np.asfortranarray(rng.standard_normal((40, 200)))
    B0[:, 17] = 0.0
    B0[:, 50] = B0[:, 3]           # exact tie: idamax must take the first
    B0[:, 120] = 0.0
    _qrcp_case(ctx, B0, dict(rtol=1e-12), check_tail=False)
    Bz = np.zeros((24, 50), order="F")
    _qrcp_case(ctx, Bz, dict(rtol=1e-12), check_tail=False)


def test_qrcp_hilbert_sketch(ctx):
    A = o.matrixlib_hilb(1024)
    rng = np.random.default_rng(0)
    B0 = o.sketch_randn(A, np.asfortranarray(rng.standard_normal((40, 1024))))
    # rtol = 1e-12 terminates above the rounding-noise floor: full parity is meaningful
    _qrcp_case(ctx, B0, dict(rtol=1e-12), check_tail=False)


def test_qrcp_tall_sketch(ctx):
    # l > n (sub-sampling style sketch): min(l, n) = n pivots possible
    rng = np.random.default_rng(11)
    B0 = np.asfortranarray(rng.standard_normal((136, 48)) @ np.diag(10.0 ** -np.linspace(0, 6, 48)))
    _qrcp_case(ctx, B0, dict(rtol=1e-13))


@pytest.mark.parametrize("k,n", [(27, 1024), (64, 300), (100, 100), (1, 10), (250, 2000)])
def test_trsolve_matches_dtrsm(ctx, k, n):
    import brapprox
    # a genuine pivoted-QR R (graded, |R_ij| <= |R_ii|): unpivoted random triangles blow up exponentially
    rng = np.random.default_rng(k + n)
    A = _decay(max(n, 300), n, min(n, max(k, 8)), 8, 1)
    B0 = o.sketch_randn(A, np.asfortranarray(rng.standard_normal((k + 8, A.shape[0]))))
    o.geqp3_adap(B0, o.LRAOptions(rank=k, rtol=0.0))
    R = np.asfortranarray(np.triu(B0[:k, :]))
    Tg = brapprox.trsolve_T(R, ctx=ctx)
    To = o.dtrsm_upper(R[:, :k], R[:, k:])
    assert Tg.shape == (k, n - k)
    if n > k:
        assert np.max(np.abs(Tg - To)) <= 1e-10 * max(np.max(np.abs(To)), 1e-300)
