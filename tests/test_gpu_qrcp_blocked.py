"""GPU parity of the BLOCKED pivoted QR for tall matrices (csrc/qrcp_blocked.cu: dlaqps with the block's trailing update
on the FP64 tensor cores) against the real LAPACK dlaqps driven exactly as the reference drives it
(src/pqr.jl:361-418).  Criteria as for the short-sketch kernel: k, p, the block-length trace and the number of pivot
steps exact; R within 1e-13 |R_11|; tau and the reflectors weighted by |R_ii| / |R_11|; the trailing matrix in LAPACK
layout."""
import numpy as np
import pytest

import lra_oracle as o
from test_gpu_stagewise import _decay, _qrcp_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n,decades,rtol", [(700, 300, 13, 1e-9), (3000, 400, 14, 1e-11), (1500, 900, 13, 1e-8),
                                               (2000, 64, 3, 1e-13), (1025, 1300, 12, 1e-10)])
def test_blocked_qrcp_matches_dlaqps(ctx, m, n, decades, rtol):
    # full rank with a decaying spectrum: the factorization ends by the rank test at a block end, well above the rounding
    # floor (pivots chosen in the noise of an exactly rank-deficient matrix are not comparable)
    r = min(m, n)
    A = _decay(m, n, r, decades, 3)
    _qrcp_case(ctx, np.asfortranarray(A), dict(rtol=rtol))


def test_blocked_qrcp_rank_cap_block_sizes_and_flags(ctx):
    rng = np.random.default_rng(5)
    B0 = np.asfortranarray(rng.standard_normal((900, 260)))
    for kw in (dict(rank=10), dict(rank=1), dict(nb=8, rank=50), dict(nb=1, rank=20), dict(nb=32, rank=100), dict(rtol=0.5),
               dict(nb=5, rank=33)):
        _qrcp_case(ctx, B0, kw)
    # strongly graded columns: the LAWN-176 test flags columns and ends blocks early (kb trace must match)
    B1 = np.asfortranarray(B0 * (10.0 ** -np.linspace(0, 14, 260))[None, :])
    _qrcp_case(ctx, B1, dict(rtol=1e-13), check_tail=False)


def test_blocked_qrcp_full_rank_and_ties(ctx):
    rng = np.random.default_rng(2)
    B0 = np.asfortranarray(rng.standard_normal((640, 96)))
    B0[:, 50] = B0[:, 3]                     # exact tie: idamax takes the first
    B0[:, 17] = 0.0
    _qrcp_case(ctx, B0, dict(rtol=1e-13), check_tail=False)
    _qrcp_case(ctx, np.asfortranarray(rng.standard_normal((1000, 40))), dict(rtol=0.0))      # all 40 steps, n < nb blocks


def test_blocked_qrcp_beyond_the_old_row_limit(ctx):
    """More than ~20000 rows: the shape the persistent on-chip kernels refuse (sketch = :none / prange on tall A)."""
    import brapprox
    A = np.asfortranarray(_decay(30000, 96, 96, 13, 4))
    _qrcp_case(ctx, A, dict(rtol=1e-9))
    rin = o.RandomInputs(1)
    Fo = o.pqrfact(A, o.LRAOptions(rtol=1e-9, sketch="none"), rin)
    Fg = brapprox.pqrfact(A, brapprox.LRAOptions(rtol=1e-9, sketch="none"), ctx=ctx)
    assert Fg.k == Fo.k
    np.testing.assert_array_equal(Fg.p[:Fo.k], Fo.p[:Fo.k])
    nrm = np.linalg.norm(A)
    assert np.linalg.norm(A - Fg.matrix()) / nrm <= 2 * np.linalg.norm(A - Fo.matrix()) / nrm + 1e-15
