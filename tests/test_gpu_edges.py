"""Edge cases through the C ABI: zero matrices (rank 0), single rows/columns, matrices smaller than the first sketch
order, padded leading dimensions, pure absolute tolerance.  The reference reaches these through the same code as
every other input (src/sketch.jl:213-236 falls back to the full QRCP when order >= n; src/pqr.jl:396-404 returns k = 0
when the first pivot norm is 0)."""
import numpy as np
import pytest

import lra_oracle as o

pytestmark = pytest.mark.gpu

SHAPES = [(5, 7), (1, 1), (1, 9), (9, 1), (50, 3), (3, 50), (41, 41)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("zero", [True, False])
def test_small_and_zero_inputs(ctx, shape, zero):
    import brapprox
    A = np.zeros(shape) if zero else np.random.default_rng(shape[0] * 100 + shape[1]).standard_normal(shape)
    A = np.asfortranarray(A)
    rin = o.RandomInputs(0)
    Vo = o.idfact(A, o.LRAOptions(rtol=1e-10), rin)
    Vg = brapprox.idfact(A, brapprox.LRAOptions(rtol=1e-10), rand=rin.drawn, ctx=ctx)
    assert Vg.k == Vo.k
    np.testing.assert_array_equal(Vg.sk, Vo.sk)
    assert sorted(Vg.p) == list(range(1, shape[1] + 1))
    if Vo.k and shape[1] > Vo.k:
        np.testing.assert_allclose(A[:, Vg.sk - 1] @ Vg.matrix(), A, atol=1e-12 * max(1.0, np.abs(A).max()) * 50)
    rin = o.RandomInputs(0)
    Qo = o.pqrfact(A, o.LRAOptions(rtol=1e-10), rin)
    Qg = brapprox.pqrfact(A, brapprox.LRAOptions(rtol=1e-10), rand=rin.drawn, ctx=ctx)
    assert Qg.k == Qo.k
    assert Qg.Q.shape == (shape[0], Qo.k) and Qg.R.shape == (Qo.k, shape[1])
    if Qo.k:
        np.testing.assert_allclose(Qg.matrix(), A, atol=1e-12 * 50)
    rin = o.RandomInputs(0)
    So = o.psvdfact(A, o.LRAOptions(rtol=1e-10), rin)
    Sg = brapprox.psvdfact(A, brapprox.LRAOptions(rtol=1e-10), rand=rin.drawn, ctx=ctx)
    assert len(Sg.S) == len(So.S)
    if len(So.S):
        np.testing.assert_allclose(Sg.S, So.S, atol=1e-11 * So.S[0])
        np.testing.assert_allclose((Sg.U * Sg.S) @ Sg.Vt, A, atol=1e-11 * So.S[0])


def test_padded_leading_dimension(ctx):
    """A device-resident operand with lda > m gives the same factorization as its packed copy."""
    import torch
    import brapprox
    m, n, pad = 700, 640, 37
    A = o.decaying_matrix(m, n, 90, 11.0, 90, seed=3)
    rin = o.RandomInputs(2)
    Vo = o.idfact(A, o.LRAOptions(rtol=1e-9), rin)
    t = torch.zeros((n, m + pad), dtype=torch.float64, device="cuda")
    t[:, :m] = torch.from_numpy(np.ascontiguousarray(A.T)).cuda()
    t[:, m:] = float("nan")                                            # the padding must never be read
    dA = brapprox.DeviceMatrix(t.data_ptr(), m, n, m + pad, keep=t)
    Vg = brapprox.idfact(dA, brapprox.LRAOptions(rtol=1e-9), rand=rin.drawn, ctx=ctx)
    assert Vg.k == Vo.k
    np.testing.assert_array_equal(Vg.p, Vo.p)
    Sg = brapprox.psvdfact(dA, brapprox.LRAOptions(rtol=1e-9), rand=rin.drawn, ctx=ctx)
    assert np.isfinite(Sg.S).all()
    np.testing.assert_allclose((Sg.U * Sg.S) @ Sg.Vt, A, atol=1e-7 * Sg.S[0])


def test_absolute_tolerance_only(ctx):
    import brapprox
    A = o.decaying_matrix(600, 500, 200, 12.0, 200, seed=8)
    for atol in (1e-3, 1e-7):
        rin = o.RandomInputs(4)
        Vo = o.idfact(A, o.LRAOptions(atol=atol, rtol=0.0), rin)
        Vg = brapprox.idfact(A, brapprox.LRAOptions(atol=atol, rtol=0.0), rand=rin.drawn, ctx=ctx)
        assert Vg.k == Vo.k
        np.testing.assert_array_equal(Vg.p, Vo.p)
