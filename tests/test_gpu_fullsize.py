"""GPU parity at the FULL sizes of BASELINE.json's single-GPU configs, against the oracle on identical random inputs
(caller-supplied Omega_t / (d, idx)), through the C ABI:
  C2  idfact + psvdfact, 8192 x 8192, sigma_j = 10^(-12 j / 500), rtol = 1e-12, sketch = :randn
  C3  idfact, 16384 x 16384, same spectrum, sketch = :srft
Criteria (BASELINE.json north_star): rounds, k and p exact; C*T within 1e-10 ||A|| entrywise (the attainable
end-to-end factor check, SURVEY.md section 7 hard part 3); spectral error snormdiff(A, F)/snorm(A) within 2x of the
oracle's; psvd: rank exact, |dsigma| <= 1e-10 sigma_1, U*S and S*Vt entrywise 1e-10 sigma_1 after sign fixing.
||A||_2 = sigma_1 = 1 by construction.
  C4  pqrfact 1 048 576 x 4096 rank = 256 and C5 batched idfact of 16 384 blocks 512 x 512 exist only on the device: checked
      through size-independent properties (rank cap and round trace, orthonormality, reconstruction error, determinism, exact
      linearity under a power-of-two scaling, replay of sampled blocks through the parity interface).  They need ~100 GB /
      ~60 GB of free device memory and are skipped on a smaller or busier GPU."""
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


def test_c2_fast_mode_full_size_replayed_through_the_oracle(ctx):
    """The benchmark's own configuration -- fast mode, the library's nested device Omega, A host-resident (pipelined
    upload with the stacked speculative sketch) -- replayed through the oracle on the random numbers read back from the
    library (tests/test_gpu_fastmode_replay.py): rounds, k, p exact, C T within 1e-10, psvd values within 1e-10."""
    import brapprox
    from test_gpu_fastmode_replay import DeviceDraws
    A = o.decaying_matrix(8192, 8192, RANK_GEN, DECADES, JDIV, seed=1)
    seed = 2024
    Vg = brapprox.idfact(A, rtol=RTOL, seed=seed, ctx=ctx)
    assert brapprox.lib.bra_debug_sketch_rows(ctx.handle) == 520          # nested: max, not sum, of the orders
    Vo = o.idfact(A, o.LRAOptions(rtol=RTOL), DeviceDraws(ctx, seed))
    assert [l for l, _ in Vo.rounds] == [40, 72, 136, 264, 520]
    _check_id(A, Vo, Vg)
    Fg = brapprox.psvdfact(A, rtol=RTOL, seed=seed, ctx=ctx)
    Fo = o.psvdfact(A, o.LRAOptions(rtol=RTOL), DeviceDraws(ctx, seed))
    assert len(Fg.S) == len(Fo.S)
    assert np.max(np.abs(Fg.S - Fo.S)) <= 1e-10 * Fo.S[0]
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


# ---- C4 / C5 at full size: the matrices exist only on the device (34 GB each), so the checks are the size-independent
# ---- properties the domain offers: exact rank cap and round trace, orthonormality, reconstruction at the level of the
# ---- first discarded singular value, determinism, exact linearity under a power-of-two scaling, replay of sampled blocks

def _need_gb(gb):
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < gb * (1 << 30):
        pytest.skip(f"needs {gb} GB of free device memory")


def test_c4_tall_pqrfact_full_size_properties(ctx):
    """C4: pqrfact of 1 048 576 x 4096, rank = 256 cap, sketch = :randn, fast mode (one GPU holds all rows here)."""
    import ctypes as C
    import torch
    from brapprox import _binding as B
    from brapprox._binding import DeviceMatrix
    from brapprox._frontend import pqrfact_device
    _need_gb(100)
    dev = torch.device("cuda", 0)
    m, n, r = 1 << 20, 4096, 512
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    Y, _ = torch.linalg.qr(torch.randn(n, r, dtype=torch.float64, device=dev, generator=g))
    sig = 10.0 ** (-6.0 * torch.arange(r, dtype=torch.float64, device=dev) / 256.0)
    Ys = (Y * sig).T.contiguous()
    At = torch.empty((n, m), dtype=torch.float64, device=dev)            # column-major m x n
    for c0 in range(0, m, 65536):
        X = torch.randn((65536, r), dtype=torch.float64, device=dev, generator=g) / (m ** 0.5)
        At[:, c0:c0 + 65536] = (X @ Ys).T
    del Y, Ys, X
    A = DeviceMatrix(At.data_ptr(), m, n, m, keep=At)
    torch.cuda.synchronize()

    def run():
        inf = pqrfact_device(A, rtol=1e-12, rank=256, seed=11, ctx=ctx)
        k = int(inf.k)
        p = ctx.fetch(B.F_P, (n,), np.int64)
        R = ctx.fetch(B.F_R, (k, n))
        return inf, k, p, R

    inf, k, p, R = run()
    assert k == 256                                                       # the rank cap binds (src/pqr.jl:352)
    assert [int(inf.orders[i]) for i in range(inf.rounds)] == [40, 72, 136, 264, 520]
    assert sorted(p) == list(range(1, n + 1))
    # Q (m x k) stays on the device: fetch into a torch buffer, check orthonormality and the reconstruction there
    Q = torch.empty((k, m), dtype=torch.float64, device=dev)              # column-major m x k
    ctx.check(B.lib.bra_fetch(ctx.handle, B.F_Q, C.c_void_p(Q.data_ptr()), m))
    torch.cuda.synchronize()
    G = Q @ Q.T
    assert float(torch.linalg.norm(G - torch.eye(k, dtype=torch.float64, device=dev))) <= 1e-12 * np.sqrt(k)
    Rt = torch.from_numpy(np.ascontiguousarray(R.T)).to(dev)              # n x k  (R' rows = pivoted columns)
    pidx = torch.from_numpy(p - 1).to(dev)
    err2 = nrm2 = 0.0
    for c0 in range(0, m, 131072):
        Ac = At[pidx, c0:c0 + 131072]                                     # (A P)' chunk: n x rows
        D = Ac - Rt @ Q[:, c0:c0 + 131072]
        err2 += float(torch.sum(D * D))
        nrm2 += float(torch.sum(Ac * Ac))
    rel = (err2 / nrm2) ** 0.5
    assert rel <= 5e-6, rel                                               # sigma_257 / sigma_1 = 1e-6: the cap's own error
    del Q, G, D, Ac
    # determinism: the same call gives the same bits
    _, k2, p2, R2 = run()
    assert k2 == k and np.array_equal(p2, p) and np.array_equal(R2, R)
    # exact linearity under a power-of-two scaling: same pivots, R scaled exactly
    At.mul_(4.0)
    torch.cuda.synchronize()
    _, k3, p3, R3 = run()
    assert k3 == k and np.array_equal(p3, p) and np.array_equal(R3, 4.0 * R)


def test_c5_batched_idfact_full_size_properties(ctx):
    """C5: 16 384 independent 512 x 512 Cauchy blocks, sketch = :sprn, rtol = 1e-12, fast mode."""
    import torch
    import brapprox
    from test_gpu_batched import _fast_mode_draws
    from brapprox._frontend import idfact_batched_device
    _need_gb(60)
    dev = torch.device("cuda", 0)
    nb, m = 16384, 512
    n = m
    At = torch.empty((nb, n, m), dtype=torch.float64, device=dev)         # block b column-major, lda = m
    g = torch.Generator(device=dev)
    g.manual_seed(50)
    for c0 in range(0, nb, 1024):
        x = torch.sort(torch.rand((1024, m), dtype=torch.float64, device=dev, generator=g), dim=1).values
        y = torch.sort(torch.rand((1024, n), dtype=torch.float64, device=dev, generator=g), dim=1).values + 1.02
        At[c0:c0 + 1024] = 1.0 / (x[:, None, :] - y[:, :, None])
    ld_t = 32

    def run(seed):
        kd = torch.zeros(nb, dtype=torch.int64, device=dev)
        pd = torch.zeros((nb, n), dtype=torch.int64, device=dev)
        Td = torch.zeros((nb, n, ld_t), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        unfinished = idfact_batched_device(At.data_ptr(), nb, m, n, m, m * n, kd.data_ptr(), pd.data_ptr(), Td.data_ptr(),
                                           ld_t, ld_t * n, None, ctx=ctx, rtol=1e-12, sketch="sprn", seed=seed)
        torch.cuda.synchronize()
        return unfinished, kd, pd, Td

    unf, kd, pd, Td = run(7)
    ks = kd.cpu().numpy()
    assert unf == 0 and ks.min() >= 8 and ks.max() <= 31                  # every block finishes in the fused first round
    ps = pd.cpu().numpy()
    assert np.all(np.sort(ps, axis=1) == np.arange(1, n + 1)[None, :])    # every p is a permutation
    # determinism
    _, kd2, pd2, Td2 = run(7)
    assert torch.equal(kd, kd2) and torch.equal(pd, pd2) and torch.equal(Td, Td2)
    # sampled blocks: ID error, and the same k / p / T when the block's own draws are replayed through the parity interface
    for b in (0, 1, 4095, 9000, 16383):
        Ab = np.asfortranarray(At[b].cpu().numpy().T)                     # m x n
        k = int(ks[b])
        sk, rd = ps[b, :k] - 1, ps[b, k:] - 1
        T = Td[b, :n - k, :k].cpu().numpy().T
        nrm = np.linalg.norm(Ab, 2)
        assert np.linalg.norm(Ab[:, rd] - Ab[:, sk] @ T, 2) <= 1e-8 * nrm
        V = brapprox.idfact(Ab, brapprox.LRAOptions(rtol=1e-12, sketch="sprn"), rand=[_fast_mode_draws(7, b, m)], ctx=ctx)
        assert V.k == k
        np.testing.assert_array_equal(V.p[:k], ps[b, :k])
        assert np.max(np.abs(Ab[:, sk] @ (V.T - T))) <= 1e-9 * nrm
