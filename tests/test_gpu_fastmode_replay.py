"""Fast mode against the oracle.  In fast mode the library draws its own random numbers on the device (nested Gaussian
rounds, distinct SRFT frequencies / subset rows, a host Fisher-Yates permutation for :sprn), so there is no "same Omega"
to hand to the oracle -- unless the very numbers are read back: bra_debug_randn / bra_debug_meta return the streams the
factorization used, this file rebuilds every round's random inputs from them exactly as csrc/api.cu defines them and
replays the factorization through the oracle.  Criteria as in parity mode: rounds, k, p identical; C T within
1e-10 ||A||; error within 2x."""
import ctypes as C

import numpy as np
import pytest

import lra_oracle as o

pytestmark = pytest.mark.gpu


class DeviceDraws(o.RandomInputs):
    """RandomInputs whose draws are the library's own fast-mode streams for (seed, round)."""

    def __init__(self, ctx, seed):
        super().__init__(seed)
        self.ctx = ctx
        self.rows = None           # nested Gaussian rows so far

    def _randn(self, count, stream):
        import brapprox
        out = np.empty(count)
        self.ctx.check(brapprox.lib.bra_debug_randn(self.ctx.handle, C.c_void_p(out.ctypes.data), count, self.seed, stream))
        return out

    def _meta(self, kind, count, rng_, stream, dtype):
        import brapprox
        out = np.empty(count, dtype=dtype)
        self.ctx.check(brapprox.lib.bra_debug_meta(self.ctx.handle, kind, C.c_void_p(out.ctypes.data), count, rng_, self.seed,
                                                   stream))
        return out

    def draw(self, kind, rnd, order, m):
        if kind == "randn":
            # nested: the rows of round t are the rows of round t-1 followed by (order - have) rows of stream t
            have = 0 if self.rows is None else self.rows.shape[0]
            ldt = (m + 1) & ~1
            fresh = self._randn((order - have) * ldt, rnd).reshape(order - have, ldt)[:, :m]
            self.rows = fresh if self.rows is None else np.vstack([self.rows, fresh])
            out = {"Omega": np.asfortranarray(self.rows.copy())}
        elif kind == "sub":
            out = {"r": self._meta(1, order, m, rnd, np.int64)}
        elif kind == "srft":
            out = {"d": self._meta(0, m, 0, rnd, np.float64), "idx": self._meta(3, order, m, rnd, np.int64)}
        elif kind == "sprn":
            out = {"perm": self._meta(2, m, m, rnd, np.int64), "s": self._randn(m, rnd)}
        else:
            raise ValueError(kind)
        self.drawn.append({"kind": kind, "round": rnd, "order": order, **out})
        return out


@pytest.mark.parametrize("kind", ["randn", "srft", "sub", "sprn"])
@pytest.mark.parametrize("trans", ["n", "c"])
def test_fast_mode_replayed_through_the_oracle(ctx, kind, trans):
    import brapprox
    m, n, r = (900, 760, 150) if trans == "n" else (700, 980, 150)
    A = o.decaying_matrix(m, n, r, 12.0, r, seed=31 + m)
    rtol, seed = 1e-10, 12345
    Vg = brapprox.idfact(A, rtol=rtol, sketch=kind, seed=seed, trans=trans, ctx=ctx)
    rin = DeviceDraws(ctx, seed)
    Vo = o.idfact(A, o.LRAOptions(rtol=rtol, sketch=kind), rin, trans)
    assert Vg.rounds == Vo.rounds
    assert len(Vg.sk) == len(Vo.sk)
    np.testing.assert_array_equal(Vg.sk, Vo.sk)
    np.testing.assert_array_equal(Vg.rd, Vo.rd)
    Aop = A if trans == "n" else A.T
    Cs = Aop[:, Vo.sk - 1]
    assert np.max(np.abs(Cs @ Vg.T - Cs @ Vo.T)) <= 1e-10 * np.linalg.norm(Aop, 2)       # the parity-mode criterion
    assert o.id_error(A, Vg, trans) <= 2 * o.id_error(A, Vo, trans) + 1e-15
    if kind == "randn":
        assert len(Vg.rounds) >= 3                                  # the nested path was exercised
        assert brapprox.lib.bra_debug_sketch_rows(ctx.handle) == max(rr[0] for rr in Vg.rounds)
    if kind == "srft":
        idx = rin.drawn[-1]["idx"][0::2] - 1                        # the entries srft_apply! reads: distinct classes
        mm = rin.drawn[-1]["d"].shape[0]
        cls = np.minimum(idx, mm - idx)
        assert len(set(cls.tolist())) == len(cls) and cls.min() >= 1
    if kind == "sub":
        for dr in rin.drawn:                                       # distinct whenever the order fits the range
            rr, mm = dr["r"], (A.shape[0] if trans == "n" else A.shape[1])
            if len(rr) <= mm:
                assert len(set(rr.tolist())) == len(rr)


def test_fast_mode_psvdfact_replayed_through_the_oracle(ctx):
    import brapprox
    A = o.decaying_matrix(1100, 840, 180, 12.0, 180, seed=5)
    rtol, seed = 1e-10, 777
    Fg = brapprox.psvdfact(A, rtol=rtol, seed=seed, ctx=ctx)
    Fo = o.psvdfact(A, o.LRAOptions(rtol=rtol), DeviceDraws(ctx, seed))
    assert len(Fg.S) == len(Fo.S)
    assert np.max(np.abs(Fg.S - Fo.S)) <= 1e-10 * Fo.S[0]
    nrm = Fo.S[0]
    eo = np.linalg.norm(A - (Fo.U * Fo.S) @ Fo.Vt, 2) / nrm
    eg = np.linalg.norm(A - Fg.matrix(), 2) / nrm
    assert eg <= 2 * eo + 1e-14
