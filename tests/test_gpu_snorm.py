"""GPU parity of snorm / snormdiff (src/snorm.jl:14-53) against the oracle on the same start vector.

The iteration is a power method on A'A (or A for Hermitian A): every product is a sum whose order differs between the
device's split reductions and BLAS, so iterates agree to a few ulps times the iteration count.  Criteria: the returned
estimate within 1e-12 relative; the iteration count (visible through the estimate at an iteration limit) identical;
against the true norm the estimate is a lower bound within the reference's own test tolerance.
"""
import numpy as np
import pytest

import lra_oracle as o

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n", [(700, 500), (300, 900), (1, 40), (64, 1)])
def test_snorm_dense_matches_oracle(ctx, m, n):
    import brapprox
    rng = np.random.default_rng(m * 7 + n)
    A = o.decaying_matrix(m, n, min(m, n), 6.0, min(m, n), seed=m + n) if min(m, n) > 1 else rng.standard_normal((m, n))
    x0 = rng.standard_normal(n)
    so = o.snorm_dense(A, o.LRAOptions(), x0=x0)
    sg = brapprox.snorm(A, brapprox.LRAOptions(), x0=x0, ctx=ctx)
    assert abs(sg - so) <= 1e-12 * so
    true = np.linalg.norm(A, 2)
    assert sg <= true * (1 + 1e-12) and sg >= 0.9 * true


def test_snorm_hermitian_branch(ctx):
    import brapprox
    rng = np.random.default_rng(5)
    B = rng.standard_normal((400, 60))
    A = B @ B.T - 0.5 * np.eye(400)                                 # symmetric indefinite
    A = (A + A.T) / 2
    x0 = rng.standard_normal(400)
    so = o.snorm_dense(A, o.LRAOptions(), x0=x0)
    sg = brapprox.snorm(A, brapprox.LRAOptions(), x0=x0, ctx=ctx)
    assert abs(sg - so) <= 1e-12 * so
    # one entry off symmetry -> the general branch (sqrt of the A'A estimate)
    A2 = A.copy()
    A2[3, 7] += 1e-3
    so2 = o.snorm_dense(A2, o.LRAOptions(), x0=x0)
    sg2 = brapprox.snorm(A2, brapprox.LRAOptions(), x0=x0, ctx=ctx)
    assert abs(sg2 - so2) <= 1e-12 * so2


def test_snorm_iteration_limit(ctx):
    import brapprox
    rng = np.random.default_rng(9)
    A = rng.standard_normal((500, 500))                              # flat spectrum: slow convergence
    x0 = rng.standard_normal(500)
    for nit in (1, 3, 8):
        so = o.snorm_dense(A, o.LRAOptions(snorm_niter=nit), x0=x0)
        sg = brapprox.snorm(A, brapprox.LRAOptions(snorm_niter=nit), x0=x0, ctx=ctx)
        assert abs(sg - so) <= 1e-12 * so


def test_snorm_device_start_vector(ctx):
    """Without x0 the start vector is the device's Philox stream: deterministic per seed, a valid lower bound."""
    import brapprox
    A = o.decaying_matrix(600, 450, 100, 8.0, 100, seed=2)
    s1 = brapprox.snorm(A, brapprox.LRAOptions(seed=4), ctx=ctx)
    s2 = brapprox.snorm(A, brapprox.LRAOptions(seed=4), ctx=ctx)
    assert s1 == s2
    true = np.linalg.norm(A, 2)
    assert 0.95 * true <= s1 <= true * (1 + 1e-12)


@pytest.mark.parametrize("m,n,r,rtol", [(900, 700, 120, 1e-9), (500, 800, 60, 1e-6)])
def test_snormdiff_factorizations(ctx, m, n, r, rtol):
    """The reference's own accuracy check: snormdiff(A, F) <= approx_rtol * snorm(A) (test/{id,pqr,psvd}.jl)."""
    import brapprox
    A = o.decaying_matrix(m, n, r, 13.0, r, seed=m)
    rng = np.random.default_rng(1)
    x0 = rng.standard_normal(n)
    nrm = brapprox.snorm(A, x0=x0, ctx=ctx)
    opts = brapprox.LRAOptions(rtol=rtol)
    F = brapprox.psvdfact(A, opts, ctx=ctx)
    eo = o.snormdiff_lowrank(A, F.U * F.S, F.Vt, x0=x0)
    eg = brapprox.snormdiff(A, F, x0=x0, ctx=ctx)
    assert abs(eg - eo) <= 1e-9 * eo + 1e-15 * nrm
    assert eg <= 100 * rtol * nrm
    Q = brapprox.pqrfact(A, opts, ctx=ctx)
    eq = brapprox.snormdiff(A, Q, x0=x0, ctx=ctx)
    R = np.zeros_like(Q.R)
    R[:, Q.p - 1] = Q.R
    eo = o.snormdiff_lowrank(A, Q.Q, R, x0=x0)
    assert abs(eq - eo) <= 1e-9 * eo + 1e-15 * nrm
    assert eq <= 100 * rtol * nrm
    V = brapprox.idfact(A, opts, ctx=ctx)
    C = A[:, V.sk - 1]
    ei = brapprox.snormdiff(A, C, V.matrix(), x0=x0, ctx=ctx)
    eo = o.snormdiff_lowrank(A, C, V.matrix(), x0=x0)
    assert abs(ei - eo) <= 1e-9 * eo + 1e-15 * nrm
    assert ei <= 100 * rtol * nrm


def test_snormdiff_dimension_mismatch(ctx):
    import brapprox
    A = np.zeros((10, 8))
    with pytest.raises(ValueError):
        brapprox.snormdiff(A, np.zeros((9, 2)), np.zeros((2, 8)), ctx=ctx)
