"""Float32 matrices through the front-ends (the reference is generic in T: LRAOptions(T), src/LowRankApprox.jl:96-118,
f(A::AbstractMatOrLinOp{T}, opts = LRAOptions(T); ...)).  The library widens a Float32 A on the device
(bra_widen_f32), factors it with the FP64 kernels and the mirror rounds the factors.

Criteria: the factors come back as Float32; rank and pivots are those of the oracle run on the widened matrix with the
T = Float32 default options (rtol = 5 eps(Float32)) and the same random inputs; T / U S V' entries within 1e-6 (Float32
rounding of FP64 results that agree to 1e-10); reconstruction error of the order of rtol."""
import numpy as np
import pytest

import lra_oracle as o

pytestmark = pytest.mark.gpu

EPS32 = float(np.finfo(np.float32).eps)


def _a32(m, n, r, seed):
    return np.asfortranarray(o.decaying_matrix(m, n, r, 9.0, r, seed=seed).astype(np.float32))


def test_idfact_float32_matches_oracle_on_widened_matrix(ctx):
    import brapprox
    A32 = _a32(700, 560, 90, 3)
    A64 = np.asfortranarray(A32.astype(np.float64))
    rin = o.RandomInputs(4)
    Vo = o.idfact(A64, o.LRAOptions(rtol=5 * EPS32), rin)
    Vg = brapprox.idfact(A32, rand=rin.drawn, ctx=ctx)                   # no opts: LRAOptions(Float32) defaults
    assert Vg.T.dtype == np.float32
    np.testing.assert_array_equal(Vg.sk, Vo.sk)
    np.testing.assert_array_equal(Vg.rd, Vo.rd)
    assert np.max(np.abs(Vg.T - Vo.T)) <= 1e-6 * max(1.0, np.max(np.abs(Vo.T)))
    # identical to the FP64 call on the widened matrix, rounded
    V64 = brapprox.idfact(A64, brapprox.LRAOptions(rtol=5 * EPS32), rand=rin.drawn, ctx=ctx)
    np.testing.assert_array_equal(Vg.T, V64.T.astype(np.float32))
    k = len(Vg.sk)
    assert 0 < k < 90
    rec = np.zeros_like(A64)
    rec[:, Vg.sk - 1] = A64[:, Vg.sk - 1]
    rec[:, Vg.rd - 1] = A64[:, Vg.sk - 1] @ Vg.T.astype(np.float64)
    assert np.linalg.norm(A64 - rec, 2) <= 1e3 * 5 * EPS32 * np.linalg.norm(A64, 2)


def test_psvdfact_pqrfact_float32(ctx):
    import brapprox
    A32 = _a32(640, 720, 80, 5)
    A64 = A32.astype(np.float64)
    nrm = np.linalg.norm(A64, 2)
    F = brapprox.psvdfact(A32, seed=1, ctx=ctx)
    assert F.U.dtype == F.S.dtype == F.Vt.dtype == np.float32
    kk = len(F.S)
    s = np.linalg.svd(A64, compute_uv=False)
    assert 0 < kk < 80
    assert np.max(np.abs(F.S - s[:kk])) <= 1e-5 * s[0]
    rec = (F.U.astype(np.float64) * F.S.astype(np.float64)) @ F.Vt.astype(np.float64)
    assert np.linalg.norm(A64 - rec, 2) <= 1e3 * 5 * EPS32 * nrm
    # explicit options are honoured as given (FP64 tolerance on Float32 data -> full numerical rank of the data)
    F2 = brapprox.psvdfact(A32, rtol=1e-4, seed=1, ctx=ctx)
    assert len(F2.S) < kk
    Q = brapprox.pqrfact(A32, seed=2, ctx=ctx)
    assert Q.Q.dtype == np.float32 and Q.R.dtype == np.float32
    rec = np.zeros_like(A64)
    rec[:, Q.p - 1] = Q.Q.astype(np.float64) @ Q.R.astype(np.float64)
    assert np.linalg.norm(A64 - rec, 2) <= 1e3 * 5 * EPS32 * nrm
    with pytest.raises(TypeError):
        brapprox.psvdfact(A32, out=(None, None, None), ctx=ctx)
    with pytest.raises(TypeError):
        brapprox.idfact(A32.astype(np.complex64), ctx=ctx)


def test_pheigfact_cur_float32(ctx):
    import brapprox
    n, r = 500, 60
    rng = np.random.default_rng(8)
    V, _ = np.linalg.qr(rng.standard_normal((n, r)))
    lam = 10.0 ** (-8.0 * np.arange(r) / r) * np.where(np.arange(r) % 2 == 1, -1.0, 1.0)
    S32 = ((V * lam) @ V.T).astype(np.float32)
    S32 = np.asfortranarray(0.5 * (S32 + S32.T))
    S64 = S32.astype(np.float64)
    F = brapprox.pheigfact(S32, seed=3, ctx=ctx)
    assert F.values.dtype == np.float32 and F.vectors.dtype == np.float32
    Vv = F.vectors.astype(np.float64)
    assert np.linalg.norm(S64 - (Vv * F.values.astype(np.float64)) @ Vv.T, 2) <= 1e3 * 5 * EPS32
    A32 = _a32(520, 430, 40, 9)
    A64 = A32.astype(np.float64)
    U = brapprox.curfact(A32, seed=4, ctx=ctx)
    U64 = brapprox.curfact(A64, brapprox.LRAOptions.for_eltype(np.float32), seed=4, ctx=ctx)
    np.testing.assert_array_equal(U.rows, U64.rows)
    np.testing.assert_array_equal(U.cols, U64.cols)
    Fc = brapprox.CUR(A32, U, ctx=ctx)
    assert Fc.C.dtype == np.float32 and Fc.R.dtype == np.float32 and Fc.U.S.dtype == np.float32
    np.testing.assert_array_equal(Fc.C, A32[:, U.cols - 1])             # gathers are exact through widen + round
    np.testing.assert_array_equal(Fc.R, A32[U.rows - 1, :])
