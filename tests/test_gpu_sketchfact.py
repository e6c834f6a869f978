"""GPU parity of the public sketchfact(side, trans, A, opts) (src/sketch.jl:52-66) against the oracle on identical
random inputs: the factors of the SKETCH itself (not of A).

Left side: k, p identical; R = triu(B[1:k, :]) within 1e-12 |R_11|; Q (orgqr of the sketch's reflectors) within 1e-10
weighted by |R_jj| / |R_11|; T within the ID criterion (C T agrees to 1e-10 ||B||).  Right side: k, leading pivots, Q up to
the sign of each column, R with the same sign convention."""
import numpy as np
import pytest

import lra_oracle as o

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["randn", "srft", "sprn", "sub"])
@pytest.mark.parametrize("trans", ["n", "c"])
def test_sketchfact_left_matches_oracle(ctx, kind, trans):
    import brapprox
    A = o.decaying_matrix(420, 360, 40, 9.0, 40, seed=31)
    rin = o.RandomInputs(5)
    kw = dict(rtol=1e-8, sketch=kind, pqrfact_retval="qrt")
    Fo = o.sketchfact(A, o.LRAOptions(**kw), rin, trans)
    Fg = brapprox.sketchfact(A, brapprox.LRAOptions(**kw), side="left", trans=trans, rand=rin.drawn, ctx=ctx)
    assert isinstance(Fg, brapprox.PQRFactors)
    assert Fg.k == Fo.k and Fg.rounds == Fo.rounds
    k = Fo.k
    np.testing.assert_array_equal(Fg.p[:k], Fo.p[:k])
    r11 = abs(Fo.R[0, 0])
    assert np.max(np.abs(Fg.R[:, :k] - Fo.R[:, :k])) <= 1e-12 * r11
    # R12 column by column through each side's own permutation (noise pivots beyond k may differ)
    inv_g, inv_o = np.argsort(Fg.p), np.argsort(Fo.p)
    rest = np.setdiff1d(np.arange(len(Fo.p)), Fo.p[:k] - 1)
    assert np.max(np.abs(Fg.R[:, inv_g[rest]] - Fo.R[:, inv_o[rest]])) <= 1e-12 * r11
    w = np.abs(np.diag(Fo.R[:, :k])) / r11
    assert Fg.Q.shape == Fo.Q.shape
    assert np.max(np.abs(Fg.Q - Fo.Q) * w[None, :]) <= 1e-10
    assert np.linalg.norm(Fg.Q.T @ Fg.Q - np.eye(k)) <= 1e-13 * np.sqrt(k)
    if np.array_equal(Fg.p, Fo.p):
        C = Fo.R[:, :k]
        assert np.max(np.abs(C @ Fg.T - C @ Fo.T)) <= 1e-10 * r11


def test_sketchfact_default_retval_is_partialqr(ctx):
    import brapprox
    A = o.decaying_matrix(300, 200, 20, 8.0, 20, seed=3)
    rin = o.RandomInputs(1)
    Fo = o.sketchfact(A, o.LRAOptions(rtol=1e-7), rin, "n")
    Fg = brapprox.sketchfact(A, brapprox.LRAOptions(rtol=1e-7), rand=rin.drawn, ctx=ctx)
    assert isinstance(Fg, brapprox.PartialQR)
    assert Fg.k == Fo.k
    np.testing.assert_array_equal(Fg.p[:Fo.k], Fo.p[:Fo.k])
    # Q R P' reproduces the sketch B = Omega A
    Bsk = o.apply_sketch("randn", A, Fo.rounds[-1][0], rin.drawn[-1], "n")
    assert np.linalg.norm(Bsk - Fg.matrix()) <= 1e-7 * np.linalg.norm(Bsk) * 10


@pytest.mark.parametrize("kind", ["randn", "srft", "sprn"])
@pytest.mark.parametrize("trans", ["n", "c"])
def test_sketchfact_right_matches_oracle(ctx, kind, trans):
    import brapprox
    A = o.decaying_matrix(420, 360, 40, 9.0, 40, seed=33)
    rin = o.RandomInputs(7)
    kw = dict(rtol=1e-8, sketch=kind, pqrfact_retval="qr")
    Fo = o.sketchfact(A, o.LRAOptions(**kw), rin, trans, side="right")
    Fg = brapprox.sketchfact(A, brapprox.LRAOptions(**kw), side="right", trans=trans, rand=rin.drawn, ctx=ctx)
    assert Fg.k == Fo.k and Fg.rounds == Fo.rounds
    k = Fo.k
    d = np.sign(np.diag(Fo.R[:, :k]))
    Qo, Ro = Fo.Q * d, Fo.R * d[:, None]
    r11 = abs(Ro[0, 0])
    w = np.abs(np.diag(Ro)) / r11
    same = int(np.flatnonzero(np.append(Fg.p[:k] != Fo.p[:k], True))[0])
    if same < k - 8:
        # an early difference must be a TIE (north star: pivots exact except at column-norm ties within 1e-13): the
        # right-hand SRFT sketch has columns (Re, Im) of frequencies f and m - f, whose norms are equal in exact
        # arithmetic.  The residual norm of the device's choice, read off the oracle's R, equals the oracle's pivot norm.
        assert kind == "srft"
        inv_o = np.argsort(Fo.p)
        jg = Fg.p[same] - 1
        res_g = np.linalg.norm(Fo.R[same:, inv_o[jg]])
        assert abs(res_g - abs(Fo.R[same, same])) <= 1e-12 * r11
        # the two factorizations then span the same range to the accuracy of the approximation
        Bsk = Fo.Q @ Fo.R
        eo = np.linalg.norm(Bsk - Fo.Q @ (Fo.Q.T @ Bsk))
        eg = np.linalg.norm(Bsk - Fg.Q @ (Fg.Q.T @ Bsk))
        assert eg <= 10 * max(eo, 1e-8 * np.linalg.norm(Bsk))
    assert Fg.Q.shape == Qo.shape
    assert np.linalg.norm(Fg.Q.T @ Fg.Q - np.eye(k)) <= 1e-13 * np.sqrt(k)
    assert np.max(np.abs(Fg.Q[:, :same] - Qo[:, :same]) * w[None, :same]) <= 1e-10
    assert np.max(np.abs(Fg.R[:same, :same] - Ro[:same, :same])) <= 1e-11 * r11


def test_sketchfact_argument_errors(ctx):
    import brapprox
    A = np.zeros((8, 8))
    with pytest.raises(ValueError):
        brapprox.sketchfact(A, side="middle", ctx=ctx)
    with pytest.raises(ValueError):
        brapprox.sketchfact(A, trans="t", ctx=ctx)
    with pytest.raises(ValueError):
        brapprox.sketchfact(A, sketch="none", ctx=ctx)
