"""GPU parity at the FULL sizes of BASELINE.json's single-GPU configs, against the oracle on identical random inputs
(caller-supplied Omega_t / (d, idx)), through the C ABI:
  C2  idfact + psvdfact, 8192 x 8192, sigma_j = 10^(-12 j / 500), rtol = 1e-12, sketch = :randn
  C3  idfact, 16384 x 16384, same spectrum, sketch = :srft
Criteria (BASELINE.json north_star): rounds, k and p exact; C*T within 1e-10 ||A|| entrywise (the attainable
end-to-end factor check, SURVEY.md section 7 hard part 3); spectral error snormdiff(A, F)/snorm(A) within 2x of the
oracle's; psvd: rank exact, |dsigma| <= 1e-10 sigma_1, U*S and S*Vt entrywise 1e-10 sigma_1 after sign fixing.
||A||_2 = sigma_1 = 1 by construction."""
import numpy as np
import pytest

import lra_oracle as o

pytestmark = pytest.mark.gpu

RANK_GEN, DECADES, JDIV, RTOL = 640, 12.0, 500.0, 1e-12


def _check_id(A, Vo, Vg):
    assert Vg.rounds == Vo.rounds
    assert Vg.k == Vo.k
    np.testing.assert_array_equal(Vg.p, Vo.p)
    C = np.asfortranarray(A[:, Vo.sk - 1])
    assert np.max(np.abs(C @ (Vg.T - Vo.T))) <= 1e-10        # ||A||_2 = 1
    eo, eg = o.id_error(A, Vo), o.id_error(A, Vg)
    assert eg <= 2 * eo + 1e-15, (eg, eo)
    return eo, eg


def test_c2_idfact_full_size(ctx):
    import brapprox
    A = o.decaying_matrix(8192, 8192, RANK_GEN, DECADES, JDIV, seed=1)
    rin = o.RandomInputs(0)
    Vo = o.idfact(A, o.LRAOptions(rtol=RTOL), rin)
    Vg = brapprox.idfact(A, brapprox.LRAOptions(rtol=RTOL), rand=rin.drawn, ctx=ctx)
    assert [l for l, _ in Vo.rounds] == [40, 72, 136, 264, 520]
    eo, eg = _check_id(A, Vo, Vg)
    assert eg < 100 * RTOL                                    # the reference's own test bound (test/id.jl:27-32)


def test_c2_psvdfact_full_size(ctx):
    import brapprox
    A = o.decaying_matrix(8192, 8192, RANK_GEN, DECADES, JDIV, seed=1)
    rin = o.RandomInputs(3)
    Fo = o.psvdfact(A, o.LRAOptions(rtol=RTOL), rin)
    Fg = brapprox.psvdfact(A, brapprox.LRAOptions(rtol=RTOL), rand=rin.drawn, ctx=ctx)
    assert len(Fg.S) == len(Fo.S)
    s1 = Fo.S[0]
    assert np.max(np.abs(Fg.S - Fo.S)) <= 1e-10 * s1
    USg, USo = Fg.U * Fg.S, Fo.U * Fo.S
    sgn = np.sign(np.sum(USg * USo, axis=0))
    assert np.max(np.abs(USg * sgn - USo)) <= 1e-10 * s1
    assert np.max(np.abs(Fg.Vt * (Fg.S * sgn)[:, None] - Fo.Vt * Fo.S[:, None])) <= 1e-10 * s1
    k = len(Fg.S)
    assert np.linalg.norm(Fg.U.T @ Fg.U - np.eye(k), 2) <= 1e-12 * np.sqrt(k)
    eo = o.snormdiff_lowrank(A, Fo.U * Fo.S, Fo.Vt)
    eg = o.snormdiff_lowrank(A, Fg.U * Fg.S, Fg.Vt)
    assert eg <= 2 * eo + 1e-15, (eg, eo)


def test_c3_idfact_srft_full_size(ctx):
    import brapprox
    A = o.decaying_matrix(16384, 16384, RANK_GEN, DECADES, JDIV, seed=3)
    rin = o.RandomInputs(1)
    opts = dict(rtol=RTOL, sketch="srft")
    Vo = o.idfact(A, o.LRAOptions(**opts), rin)
    Vg = brapprox.idfact(A, brapprox.LRAOptions(**opts), rand=rin.drawn, ctx=ctx)
    _check_id(A, Vo, Vg)
