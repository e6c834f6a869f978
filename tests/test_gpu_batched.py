"""GPU parity of the fused batched idfact (BASELINE config 5: independent Cauchy blocks, sketch = :sprn) against
the oracle run block by block on identical (perm, s): k and p exact, C*T to 1e-10*||A||, error within 2x."""
import numpy as np
import pytest

import lra_oracle as o

pytestmark = pytest.mark.gpu


def _cauchy_block(m, n, seed, gap=0.02):
    rng = np.random.default_rng(seed)
    x = np.sort(rng.random(m))
    y = np.sort(rng.random(n)) + 1.0 + gap
    return o.matrixlib_cauchy(x, y)


def _check_block(A, Vo, Vg):
    assert Vg.k == Vo.k
    np.testing.assert_array_equal(Vg.p, Vo.p)
    C = A[:, Vo.sk - 1]
    assert np.max(np.abs(C @ Vg.T - C @ Vo.T)) <= 1e-10 * np.linalg.norm(A, 2)
    assert o.id_error(A, Vg) <= 2 * o.id_error(A, Vo) + 1e-15


@pytest.mark.parametrize("m,n,nblocks,rtol", [(512, 512, 24, 1e-12), (256, 384, 9, 1e-10), (512, 200, 5, 1e-8)])
def test_batched_matches_blockwise_oracle(ctx, m, n, nblocks, rtol):
    import brapprox
    blocks = np.stack([_cauchy_block(m, n, 100 + b) for b in range(nblocks)])
    kw = dict(rtol=rtol, sketch="sprn")
    rands, Vos = [], []
    for b in range(nblocks):
        rin = o.RandomInputs(b)
        Vos.append(o.idfact(blocks[b], o.LRAOptions(**kw), rin))
        assert len(rin.drawn) == 1, "these blocks must finish in the first (fused) round"
        rands.append(rin.drawn[0])
    Vgs = brapprox.idfact_batched(blocks, brapprox.LRAOptions(**kw), rand=rands, ctx=ctx)
    for b in range(nblocks):
        _check_block(blocks[b], Vos[b], Vgs[b])


def test_batched_equals_single_matrix_path(ctx):
    """The fused kernel and the general single-matrix path are two implementations of the same reference loop."""
    import brapprox
    m = n = 512
    blocks = np.stack([_cauchy_block(m, n, 7 + b, gap=0.05) for b in range(6)])
    kw = dict(rtol=1e-11, sketch="sprn")
    rands = [o.RandomInputs(50 + b).draw("sprn", 0, 32, m) for b in range(6)]
    Vb = brapprox.idfact_batched(blocks, brapprox.LRAOptions(**kw), rand=rands, ctx=ctx)
    for b in range(6):
        Vs = brapprox.idfact(blocks[b], brapprox.LRAOptions(**kw), rand=[rands[b]], ctx=ctx)
        assert Vs.k == Vb[b].k
        np.testing.assert_array_equal(Vs.p, Vb[b].p)
        C = blocks[b][:, Vs.sk - 1]
        assert np.max(np.abs(C @ Vs.T - C @ Vb[b].T)) <= 1e-10 * np.linalg.norm(blocks[b], 2)


def test_batched_rank_cap_and_flat_spectrum(ctx):
    """rank cap below the numerical rank (kcap binds) and a block that does NOT finish in the fused round
    (k >= nb: continues through the general path, fast-mode inputs for the later rounds)."""
    import brapprox
    m = n = 512
    blocks = np.stack([_cauchy_block(m, n, 3), o.decaying_matrix(m, n, 120, 10.0, 120, seed=5)])
    V = brapprox.idfact_batched(blocks[:1], rtol=1e-12, sketch="sprn", rank=5, sketchfact_adap=False, seed=2, ctx=ctx)
    assert V[0].k == 5 and V[0].T.shape == (5, n - 5)
    V = brapprox.idfact_batched(blocks, rtol=1e-6, sketch="sprn", seed=2, ld_t=128, ctx=ctx)
    assert V[0].k < 32
    assert 32 <= V[1].k <= 128
    for b in range(2):
        assert o.id_error(blocks[b], V[b]) <= 1e-3


def test_batched_fast_mode_error(ctx):
    import brapprox
    m = n = 512
    blocks = np.stack([_cauchy_block(m, n, 40 + b) for b in range(32)])
    V = brapprox.idfact_batched(blocks, rtol=1e-12, sketch="sprn", seed=11, ctx=ctx)
    ks = [v.k for v in V]
    assert 8 <= min(ks) and max(ks) <= 31
    errs = [o.id_error(blocks[b], V[b]) for b in range(0, 32, 8)]
    assert max(errs) <= 1e-8, errs


def _splitmix_at(s0, i):
    M = (1 << 64) - 1
    z = (s0 + (i + 1) * 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return z ^ (z >> 31)


def _fast_mode_draws(seed, b, m):
    """Host emulation of the per-block random inputs the fused kernel draws in fast mode (csrc/batched.cu): the
    permutation = argsort of 21 random bits | 11 index bits, the weights = Box-Muller on counter-based uniforms."""
    M = (1 << 64) - 1
    bkey = (seed * 0x9E3779B97F4A7C15 + (b + 1) * 0xD1B54A32D192ED03) & M
    keys = np.array([(((_splitmix_at(bkey, r) >> 32) & ~0x7FF) & 0xFFFFFFFF) | r for r in range(m)], dtype=np.uint64)
    perm = np.argsort(keys, kind="stable").astype(np.int64) + 1
    wkey = bkey ^ 0xA5A5A5A5A5A5A5A5
    s = np.empty(m)
    for r in range(m):
        z1, z2 = _splitmix_at(wkey, 2 * r), _splitmix_at(wkey, 2 * r + 1)
        u1 = ((z1 >> 11) + 1.0) / 9007199254740992.0
        u2 = ((z2 >> 11) + 0.5) / 9007199254740992.0
        s[r] = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    return {"perm": perm, "s": s}


@pytest.mark.parametrize("m", [512, 300, 200, 700])
def test_batched_fast_mode_draws_a_permutation_per_block(ctx, m):
    """Fast mode: every block gets its OWN randperm(m) and weights (the reference draws per sketch call, i.e. per block:
    src/sketch.jl:575-579), generated inside the kernel.  Replaying the same draws through the parity interface must give
    the same k, p and T -- which also proves the in-kernel sort yields a true permutation."""
    import brapprox
    n, nb, seed = 384, 5, 23
    blocks = np.stack([_cauchy_block(m, n, 60 + b) for b in range(nb)])
    Vf = brapprox.idfact_batched(blocks, rtol=1e-11, sketch="sprn", seed=seed, ctx=ctx)
    rands = [_fast_mode_draws(seed, b, m) for b in range(nb)]
    for r in rands:
        assert sorted(r["perm"]) == list(range(1, m + 1))
    assert any(not np.array_equal(rands[0]["perm"], r["perm"]) for r in rands[1:])
    Vp = brapprox.idfact_batched(blocks, brapprox.LRAOptions(rtol=1e-11, sketch="sprn"), rand=rands, ctx=ctx)
    for b in range(nb):
        assert Vf[b].k == Vp[b].k
        np.testing.assert_array_equal(Vf[b].p[:Vf[b].k], Vp[b].p[:Vp[b].k])
        C = blocks[b][:, Vp[b].sk - 1]
        assert np.max(np.abs(C @ Vf[b].T - C @ Vp[b].T)) <= 1e-9 * np.linalg.norm(blocks[b], 2)
