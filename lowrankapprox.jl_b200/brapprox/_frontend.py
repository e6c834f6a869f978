"""Front-ends with the reference's names and argument meaning, over the C ABI."""
from __future__ import annotations

import ctypes as C
import dataclasses
import functools
import inspect
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _binding as B
from ._binding import BraError, Context, DeviceMatrix, IDPackedV, LRAOptions, lib, mat_arg

_default_ctx: Dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def _trans(trans: str) -> bytes:
    t = str(trans).lstrip(":").lower()
    if t not in ("n", "c"):
        raise ValueError("trans")                      # chktrans (src/LowRankApprox.jl:150)
    return t.encode()


def _is_f32(A) -> bool:
    return isinstance(A, np.ndarray) and A.dtype == np.float32


def _narrow(x):
    """Rounds the FP64 arrays of a result (dataclass, tuple, list) to Float32; index sets and bookkeeping stay."""
    if isinstance(x, np.ndarray):
        return x.astype(np.float32) if x.dtype == np.float64 else x
    if dataclasses.is_dataclass(x) and not isinstance(x, type):
        for f in dataclasses.fields(x):
            if f.name not in ("rounds", "steps"):
                setattr(x, f.name, _narrow(getattr(x, f.name)))
        return x
    if isinstance(x, tuple):
        return tuple(_narrow(v) for v in x)
    return x


def _eltype(fn):
    """Element-type dispatch of the reference's front-ends (f(A::AbstractMatOrLinOp{T}, opts = LRAOptions(T); ...)):
    a Float32 A gets the T = Float32 default options when none are passed, is widened on the device (bra_widen_f32),
    factored by the FP64 kernels, and the factors are rounded to Float32."""
    takes_opts = "opts" in inspect.signature(fn).parameters

    @functools.wraps(fn)
    def wrapper(A, *args, **kw):
        if not _is_f32(A):
            return fn(A, *args, **kw)
        if kw.get("out") is not None:
            raise TypeError("out= buffers are FP64; not available for a Float32 A")
        args = list(args)
        if not takes_opts:
            pass
        elif args and (args[0] is None or isinstance(args[0], LRAOptions)):
            if args[0] is None:
                args[0] = LRAOptions.for_eltype(np.float32)
        elif kw.get("opts") is None:
            kw["opts"] = LRAOptions.for_eltype(np.float32)
        ctx = kw.get("ctx") or default_context()
        kw["ctx"] = ctx
        return _narrow(fn(ctx.widen_f32(A), *args, **kw))
    return wrapper


def _opts(opts: Optional[LRAOptions], kw) -> LRAOptions:
    o = (opts or LRAOptions()).copy(**kw)              # copy(opts; args...) -- kwargs win
    o.chk()
    return o


class _RandPack:
    """Marshals per-round random inputs into a bra_rand (keeps the arrays alive)."""

    def __init__(self, rounds: Optional[Sequence[dict]], opts: Optional[LRAOptions] = None,
                 contracted: Optional[int] = None):
        self.keep = []
        self.c = B.bra_rand()
        self.c.n_rounds = 0
        if not rounds:
            return
        n = len(rounds)
        self.c.n_rounds = n
        if opts is not None and contracted is not None:
            self._check_shapes(rounds, opts, int(contracted))

        def column(key, dtype):
            arr = (C.c_void_p * n)()
            any_ = False
            for t, r in enumerate(rounds):
                v = r.get(key)
                if v is None:
                    arr[t] = None
                    continue
                any_ = True
                if isinstance(v, DeviceMatrix):
                    self.keep.append(v)
                    arr[t] = v.ptr
                else:
                    a = np.asfortranarray(v, dtype=dtype)
                    self.keep.append(a)
                    arr[t] = a.ctypes.data
            self.keep.append(arr)
            return C.cast(arr, B._pp_d) if any_ else None

        self.c.omega = column("Omega", np.float64)
        self.c.d = column("d", np.float64)
        self.c.idx = column("idx", np.int64)
        self.c.perm = column("perm", np.int64)
        self.c.s = column("s", np.float64)
        self.c.r = column("r", np.int64)


    @staticmethod
    def _round_order(o: LRAOptions, t: int) -> int:
        """Sketch order of adaptive round t / of the single non-adaptive round (src/sketch.jl:226-236, 680-686)."""
        adaptive = o.sketchfact_adap or o.rank < 0
        nn = (o.nb << t) if adaptive else o.rank
        if o.sketch == "sprn":
            return nn
        a, b = o._samp_affine()
        if a == 0 and b == 0:
            a, b = (4, 8) if o.sketch == "sub" else (1, 8)
        return a * nn + b

    @classmethod
    def _check_shapes(cls, rounds, o: LRAOptions, contracted: int) -> None:
        """The C side reads Omega as order x contracted (ld = order) and the vectors at their nominal lengths: a
        wrongly shaped input would be read past its end, so it is refused here (DimensionMismatch in the reference)."""
        want = {"Omega": lambda l: (l, contracted), "d": lambda l: (contracted,), "idx": lambda l: (l,),
                "perm": lambda l: (contracted,), "s": lambda l: (contracted,), "r": lambda l: (l,)}
        for t, r in enumerate(rounds):
            l = cls._round_order(o, t)
            for key, shp in want.items():
                v = r.get(key)
                if v is None:
                    continue
                got = v.shape if isinstance(v, DeviceMatrix) else np.shape(v)
                if isinstance(v, DeviceMatrix) and v.ld != max(v.m, 1):
                    raise ValueError(f"DimensionMismatch: round {t} {key} must be stored with ld == rows")
                if tuple(got) != shp(l):
                    raise ValueError(f"DimensionMismatch: round {t} {key} has shape {tuple(got)}, expected {shp(l)}")


def sketch(A, order: int, opts: Optional[LRAOptions] = None, side: str = "left", trans: str = "n",
           rand: Optional[dict] = None, ctx: Optional[Context] = None, **kw) -> np.ndarray:
    """sketch(side, trans, A, order, opts; kw...) (src/sketch.jl:35-50).  `rand` carries the random
    inputs the reference would draw (Omega for :randn)."""
    if str(side).lstrip(":") not in ("left", "right"):
        raise ValueError("side")                       # sketchfact_chkargs (src/sketch.jl:80-84)
    tr = _trans(trans)
    if order < 0:
        raise ValueError("order")
    o = _opts(opts, kw)
    if str(side).lstrip(":") == "right":
        # B = op(A) S is the transpose of the left sketch of op(A)' on the same random inputs (real arithmetic; the
        # mul! forms at src/sketch.jl:91-110, 248-293, 474-522, 571-653 are transposes of each other)
        flipped = "c" if tr == b"n" else "n"
        return np.asfortranarray(sketch(A, order, o, "left", flipped, rand, ctx).T)
    ctx = ctx or default_context()
    pA, m, n, lda, keepA = mat_arg(A)
    mA, nA = (m, n) if tr == b"n" else (n, m)
    out = np.zeros((order, nA), order="F")
    if o.sketch == "randn" or o.sketch == "none":
        if rand is None or "Omega" not in rand:
            raise ValueError("stage-wise sketch needs rand['Omega'] (order x contracted-dim)")
        pO, lo, mo, ldo, keepO = mat_arg(rand["Omega"])
        if (lo, mo) != (order, mA):
            raise ValueError("DimensionMismatch: Omega")
        ctx.check(lib.bra_sketch_randn_f64(ctx.handle, tr, m, n, pA, lda, order, pO, ldo,
                                           C.c_void_p(out.ctypes.data), max(order, 1)))
        return out
    outp, ldo = C.c_void_p(out.ctypes.data), max(order, 1)

    def vec(key, dtype, length):
        if rand is None or key not in rand:
            raise ValueError(f"stage-wise sketch = :{o.sketch} needs rand[{key!r}]")
        a = np.ascontiguousarray(rand[key], dtype=dtype)
        if a.shape != (length,):
            raise ValueError(f"DimensionMismatch: {key}")
        return a

    if o.sketch == "sub":
        r = vec("r", np.int64, order)
        ctx.check(lib.bra_sketch_sub_f64(ctx.handle, tr, m, n, pA, lda, order, C.c_void_p(r.ctypes.data), outp, ldo))
        return out
    if o.sketch == "sprn":
        perm, sv = vec("perm", np.int64, mA), vec("s", np.float64, mA)
        ctx.check(lib.bra_sketch_sprn_f64(ctx.handle, tr, m, n, pA, lda, order, C.c_void_p(perm.ctypes.data),
                                          C.c_void_p(sv.ctypes.data), outp, ldo))
        return out
    if o.sketch == "srft":
        d, idx = vec("d", np.float64, mA), vec("idx", np.int64, order)
        ctx.check(lib.bra_sketch_srft_f64(ctx.handle, tr, m, n, pA, lda, order, C.c_void_p(d.ctypes.data),
                                          C.c_void_p(idx.ctypes.data), outp, ldo))
        return out
    raise ValueError("sketch")                          # chkopts! already rejects unknown kinds


def geqp3_adap(Bm: np.ndarray, opts: Optional[LRAOptions] = None, ctx: Optional[Context] = None, **kw):
    """geqp3_adap!(B, opts) (src/pqr.jl:348-359): returns (B_out, jpvt 1-based, tau, k, trace) where
    B_out is the in-place result in LAPACK layout and trace = {"kb": [...], "steps": n}."""
    o = _opts(opts, kw)
    ctx = ctx or default_context()
    Bw = np.array(Bm, dtype=np.float64, order="F", copy=True)
    l, n = Bw.shape
    lmin = min(l, n)
    kcap = lmin if (o.rank < 0 or o.rank > lmin) else o.rank
    jpvt = np.zeros(n, dtype=np.int64)
    tau = np.zeros(max(kcap, 1))
    kb = np.zeros(max(kcap + 1, 1), dtype=np.int32)
    k, ns, nbk = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    co = o.to_c()
    ctx.check(lib.bra_geqp3_adap_f64(ctx.handle, l, n, C.c_void_p(Bw.ctypes.data), max(l, 1), C.byref(co),
                                     C.c_void_p(jpvt.ctypes.data), C.c_void_p(tau.ctypes.data), C.byref(k),
                                     C.byref(ns), C.c_void_p(kb.ctypes.data), kb.size, C.byref(nbk)))
    return Bw, jpvt, tau[:ns.value], int(k.value), {"kb": kb[:nbk.value].tolist(), "steps": int(ns.value)}


def trsolve_T(R: np.ndarray, ctx: Optional[Context] = None) -> np.ndarray:
    """maxdet_t(R) = R[:, :k] \\ R[:, k:] (src/pqr.jl:438-442)."""
    ctx = ctx or default_context()
    Rf = np.asfortranarray(R, dtype=np.float64)
    k, n = Rf.shape
    T = np.zeros((k, n - k), order="F")
    ctx.check(lib.bra_trsolve_T_f64(ctx.handle, k, n, C.c_void_p(Rf.ctypes.data), max(k, 1),
                                    C.c_void_p(T.ctypes.data), max(k, 1)))
    return T


def _rounds(ctx: Context):
    inf = ctx.info()
    r = [(int(inf.orders[t]), int(inf.ks[t])) for t in range(inf.rounds)]
    s = [int(inf.steps[t]) for t in range(inf.rounds)]
    return inf, r, s


def idfact_device(A, opts: Optional[LRAOptions] = None, trans: str = "n", rand=None,
                  ctx: Optional[Context] = None, **kw):
    """Runs idfact and leaves every result on the device; returns the bra_info (k, rounds...)."""
    o = _opts(opts, kw)
    o.pqrfact_retval = "t"                              # src/id.jl:438
    ctx = ctx or default_context()
    pA, m, n, lda, keepA = mat_arg(A)
    rp = _RandPack(rand, o, m if _trans(trans) == b"n" else n)
    co = o.to_c()
    ctx.check(lib.bra_idfact_f64(ctx.handle, _trans(trans), m, n, pA, lda, C.byref(co), C.byref(rp.c)))
    return ctx.info()


@_eltype
def idfact(A, opts: Optional[LRAOptions] = None, trans: str = "n", rand=None,
           ctx: Optional[Context] = None, **kw) -> IDPackedV:
    """idfact(trans, A, opts; kw...) -> IDPackedV(sk, rd, T) (src/id.jl:434-447)."""
    _trans(trans)                                       # chktrans first, like the reference (src/id.jl:436)
    ctx = ctx or default_context()
    idfact_device(A, opts, trans, rand, ctx, **kw)
    inf, rounds, steps = _rounds(ctx)
    k, n = int(inf.k), int(inf.n)
    p = ctx.fetch(B.F_P, (n,), np.int64)
    T = ctx.fetch(B.F_T, (k, n - k))
    return IDPackedV(p[:k].copy(), p[k:].copy(), T, rounds, steps)


def _skeleton(A, opts, trans, rand, ctx, kw):
    """sketchfact(:left, trans, A, opts)[:p][1:k] (the index set only; src/cur.jl:538-556 asks for retval "")."""
    if _is_f32(A):
        A = ctx.widen_f32(A)
    idfact_device(A, opts, trans, rand, ctx, **kw)
    inf, _, _ = _rounds(ctx)
    k, n = int(inf.k), int(inf.n)
    return ctx.fetch(B.F_P, (n,), np.int64)[:k].copy()


def curfact(A, opts: Optional[LRAOptions] = None, rand=None, ctx: Optional[Context] = None, **kw):
    """curfact(A, opts; kw...) -> CURPackedU(rows, cols) / HermCURPackedU(cols) (src/cur.jl:532-566).

    The control flow of the reference -- which side is sketched first, the second sketch on the selected rows or
    columns only, the truncation to k = min(kr, kc) -- with both sketch-and-pivot passes on the device.  `rand` is an
    optional pair (first pass, second pass) of per-round random inputs (parity mode).  A Hermitian (real: exactly
    symmetric) A takes one pass and returns the column set for both sides, like HermCURPackedU."""
    ctx = ctx or default_context()
    A = np.asarray(A)
    if A.ndim != 2:
        raise ValueError("A")
    if _is_f32(A) and opts is None:
        opts = LRAOptions.for_eltype(np.float32)
    r1, r2 = (rand if rand is not None else (None, None))
    m, n = A.shape
    if m == n and np.array_equal(A, A.T):
        cols = _skeleton(A, opts, "n", r1, ctx, kw)
        return B.CURPackedU(cols.copy(), cols, hermitian=True)
    if m >= n:
        rows = _skeleton(A, opts, "c", r1, ctx, kw)
        cols = _skeleton(np.asfortranarray(A[rows - 1, :]), opts, "n", r2, ctx, kw)
    else:
        cols = _skeleton(A, opts, "n", r1, ctx, kw)
        rows = _skeleton(np.asfortranarray(A[:, cols - 1]), opts, "c", r2, ctx, kw)
    k = min(len(rows), len(cols))
    return B.CURPackedU(rows[:k], cols[:k])


@_eltype
def CUR(A, rows, cols=None, ctx: Optional[Context] = None):
    """CUR(A, U::CURPackedU) / CUR(A, rows, cols) / HermCUR(A, cols) (src/cur.jl:85-109): the factors of the CUR
    decomposition named by the index sets curfact returned -- C, R and the pseudo-inverse of the k x k core
    (svd! / eigen!(Hermitian(.)) in the reference; the one-sided Jacobi on the device here)."""
    herm = False
    if isinstance(rows, B.CURPackedU):
        U = rows
        rows, cols, herm = U.rows, U.cols, U.hermitian
    elif cols is None:
        cols, herm = rows, True                         # HermCUR(A, cols)
    ctx = ctx or default_context()
    pA, m, n, lda, keepA = mat_arg(A)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    cols = np.ascontiguousarray(cols, dtype=np.int64)
    k = len(cols)
    if len(rows) != k:
        raise ValueError("DimensionMismatch: rows and cols")
    if herm and m != n:
        raise ValueError("matrix is not square")
    for idx, lim in ((rows, m), (cols, n)):
        if k and (idx.min() < 1 or idx.max() > lim):
            raise IndexError("BoundsError")
    ctx.check(lib.bra_cur_f64(ctx.handle, m, n, pA, lda, k, C.c_void_p(rows.ctypes.data), C.c_void_p(cols.ctypes.data),
                              int(herm)))
    Cm = ctx.fetch(B.F_Q, (m, k)) if k else np.zeros((m, 0), order="F")
    s = ctx.fetch(B.F_S, (k,)) if k else np.zeros(0)
    V = ctx.fetch(B.F_U, (k, k)) if k else np.zeros((0, 0), order="F")
    if herm:
        return B.CUR(rows, cols, Cm, B.PartialHermEigen(s, V), None)
    Ut = ctx.fetch(B.F_VT, (k, k)) if k else np.zeros((0, 0), order="F")
    Rm = ctx.fetch(B.F_R, (k, n)) if k else np.zeros((0, n), order="F")
    return B.CUR(rows, cols, Cm, B.PartialSVD(V, s, Ut), Rm)


def cur(A, *args, **kw):
    """cur(A, ...) -> (rows, cols) (src/cur.jl:568-571)."""
    U = curfact(A, *args, **kw)
    return U.rows, U.cols


def idfact_batched_device(a_ptr: int, nblocks: int, m: int, n: int, lda: int, stride_a: int, k_ptr: int, p_ptr: int,
                          t_ptr: int, ld_t: int, stride_t: int, opts: Optional[LRAOptions] = None,
                          perm_ptr: int = 0, perm_stride: int = 0, s_ptr: int = 0, s_stride: int = 0,
                          ctx: Optional[Context] = None, **kw) -> int:
    """Batched idfact of `nblocks` independent device-resident m x n blocks (the reference would loop idfact,
    src/id.jl:434-447): raw device pointers in, results left on the device.  Returns the number of blocks that
    needed adaptive rounds beyond the fused first one."""
    o = _opts(opts, kw)
    o.pqrfact_retval = "t"
    ctx = ctx or default_context()
    co = o.to_c()
    ctx.check(lib.bra_idfact_batched_f64(ctx.handle, nblocks, m, n, C.c_void_p(a_ptr), lda, stride_a, C.byref(co),
                                         C.c_void_p(perm_ptr or None), perm_stride, C.c_void_p(s_ptr or None), s_stride,
                                         C.c_void_p(k_ptr), C.c_void_p(p_ptr), C.c_void_p(t_ptr), ld_t, stride_t))
    return int(lib.bra_batched_unfinished(ctx.handle))


def idfact_batched(blocks, opts: Optional[LRAOptions] = None, rand: Optional[Sequence[dict]] = None,
                   ctx: Optional[Context] = None, device: int = 0, ld_t: Optional[int] = None, **kw) -> List[IDPackedV]:
    """[idfact(A_b, opts) for A_b in blocks] through the fused batched kernel.  `blocks` is an array of shape
    (nblocks, m, n); `rand[b]` = {"perm": ..., "s": ...} carries block b's reference-order random inputs (omit for
    the fast mode).  torch is used here only to hold the device buffers."""
    import torch
    o = _opts(opts, kw)
    ctx = ctx or default_context(device)
    dev = torch.device("cuda", device)
    blk = np.asarray(blocks, dtype=np.float64)
    nb, m, n = blk.shape
    At = torch.from_numpy(np.ascontiguousarray(blk.transpose(0, 2, 1))).to(dev)       # block b column-major, lda = m
    ld_t = int(ld_t or min(o.nb, m, n))
    kd = torch.zeros(nb, dtype=torch.int64, device=dev)
    pd = torch.zeros((nb, n), dtype=torch.int64, device=dev)
    Td = torch.zeros((nb, n, ld_t), dtype=torch.float64, device=dev)
    perm_ptr = s_ptr = 0
    keep = None
    if rand is not None:
        permd = torch.from_numpy(np.stack([np.asarray(r["perm"], dtype=np.int64) for r in rand])).to(dev)
        sd = torch.from_numpy(np.stack([np.asarray(r["s"], dtype=np.float64) for r in rand])).to(dev)
        keep = (permd, sd)
        perm_ptr, s_ptr = permd.data_ptr(), sd.data_ptr()
    torch.cuda.synchronize(dev)
    idfact_batched_device(At.data_ptr(), nb, m, n, m, m * n, kd.data_ptr(), pd.data_ptr(), Td.data_ptr(), ld_t,
                          ld_t * n, o, perm_ptr, m if rand is not None else 0, s_ptr, m if rand is not None else 0, ctx)
    ks, ps, Ts = kd.cpu().numpy(), pd.cpu().numpy(), Td.cpu().numpy()
    del keep
    out = []
    for b in range(nb):
        k = int(ks[b])
        T = np.asfortranarray(Ts[b, : n - k, :k].T)
        out.append(IDPackedV(ps[b, :k].copy(), ps[b, k:].copy(), T))
    return out


def id(A, *args, **kw):
    """id(...) -> (sk, rd, T) (src/id.jl:452-456)."""
    V = idfact(A, *args, **kw)
    return V.sk, V.rd, V.T


def pqrfact_device(A, opts: Optional[LRAOptions] = None, trans: str = "n", rand=None,
                   ctx: Optional[Context] = None, **kw):
    o = _opts(opts, kw)
    ctx = ctx or default_context()
    pA, m, n, lda, keepA = mat_arg(A)
    rp = _RandPack(rand, o, m if _trans(trans) == b"n" else n)
    co = o.to_c()
    ctx.check(lib.bra_pqrfact_f64(ctx.handle, _trans(trans), m, n, pA, lda, C.byref(co), C.byref(rp.c)))
    return ctx.info()


@_eltype
def pqrfact(A, opts: Optional[LRAOptions] = None, trans: str = "n", rand=None,
            ctx: Optional[Context] = None, **kw):
    """pqrfact(trans, A, opts; kw...) (src/pqr.jl:290-307), sketched path.  Returns PartialQR(Q, R, p) when
    pqrfact_retval has q and r and not t (the default "qr"), like the reference (src/pqr.jl:305)."""
    ctx = ctx or default_context()
    o = _opts(opts, kw)
    pqrfact_device(A, o, trans, rand, ctx)
    inf, rounds, steps = _rounds(ctx)
    k, n, m = int(inf.k), int(inf.n), int(inf.m)
    p = ctx.fetch(B.F_P, (n,), np.int64)
    Q = ctx.fetch(B.F_Q, (m, k)) if k > 0 else np.zeros((m, 0), order="F")
    R = ctx.fetch(B.F_R, (k, n)) if k > 0 else np.zeros((0, n), order="F")
    F = B.PartialQR(Q, R, p, rounds)
    if "t" in o.pqrfact_retval:
        F.T = ctx.fetch(B.F_T, (k, n - k))
    return F


def _orgqr(Bl: np.ndarray, tau: np.ndarray, k: int) -> np.ndarray:
    """LAPACK.orgqr!(A[:, 1:k], tau, k) (src/pqr.jl:427) on the host: the first k columns of H_1 ... H_k from the
    reflectors below the diagonal of the factored sketch (result-type arithmetic: order x k, a few hundred rows)."""
    l = Bl.shape[0]
    Q = np.zeros((l, k), order="F")
    Q[np.arange(k), np.arange(k)] = 1.0
    for j in range(k - 1, -1, -1):
        v = np.zeros(l)
        v[j] = 1.0
        v[j + 1:] = Bl[j + 1:, j]
        Q[j:, j:] -= tau[j] * np.outer(v[j:], v[j:] @ Q[j:, j:])
    return Q


@_eltype
def sketchfact(A, opts: Optional[LRAOptions] = None, side: str = "left", trans: str = "n", rand=None,
               ctx: Optional[Context] = None, **kw):
    """sketchfact(side, trans, A, opts; kw...) (src/sketch.jl:52-66): the early-terminating pivoted QR of the SKETCH of
    op(A) -- side "left": B = S op(A) (order x n_op), side "right": B = op(A) S (m_op x order).  Returns PartialQR(Q, R, p)
    when pqrfact_retval is "qr", otherwise PQRFactors(Q, R, p, k, T) with the pieces retval names (src/pqr.jl:434-435).
    Left side: Q = orgqr of the sketch's reflectors (formed on the host from BRA_F_BSKETCH / BRA_F_TAU; R = triu(B[1:k, :])).
    Right side: Q is the device's CholeskyQR2 factor of B[:, p[1:k]] (the Householder Q with diag(R) > 0), R follows the same
    sign convention.  With maxdet swaps, q / r of the swapped factorization are not maintained on the device (only p and
    T are): asking for them then raises."""
    sd = str(side).lstrip(":")
    if sd not in ("left", "right"):
        raise ValueError("side")                       # sketchfact_chkargs (src/sketch.jl:80-84)
    tr = _trans(trans)
    o = _opts(opts, kw)
    if o.sketch == "none":
        raise ValueError("sketch")                     # src/sketch.jl:62
    ctx = ctx or default_context()
    pA, m, n, lda, keepA = mat_arg(A)
    contracted = (m if tr == b"n" else n) if sd == "left" else (n if tr == b"n" else m)
    rp = _RandPack(rand, o, contracted)
    co = o.to_c()
    ctx.check(lib.bra_sketchfact_f64(ctx.handle, sd[0].encode(), tr, m, n, pA, lda, C.byref(co), C.byref(rp.c)))
    inf, rounds, steps = _rounds(ctx)
    k, nB, mB = int(inf.k), int(inf.n), int(inf.m)
    rows = int(inf.orders[inf.rounds - 1]) if sd == "left" else mB
    p = ctx.fetch(B.F_P, (nB,), np.int64)
    rv = o.pqrfact_retval
    want_q, want_r, want_t = "q" in rv, "r" in rv, "t" in rv
    maxdet = 0 < k < nB and o.maxdet_tol >= 0
    T = ctx.fetch(B.F_T, (k, nB - k)) if (want_t or maxdet) and k > 0 else (np.zeros((0, nB), order="F") if want_t else None)
    if (want_q or want_r) and maxdet and ctx.maxdet_swaps() > 0:
        raise ValueError("sketchfact: Q / R after maxdet column swaps are not maintained on the device (p and T are)")
    Q = R = None
    if want_q or want_r:
        Bl = ctx.fetch(B.F_BSKETCH, (rows, nB)) if rows > 0 and nB > 0 else np.zeros((rows, nB), order="F")
        R = np.asfortranarray(np.triu(Bl[:k, :]))
        if sd == "left":
            tau = ctx.fetch(B.F_TAU, (steps[-1],)) if steps[-1] > 0 else np.zeros(0)
            Q = _orgqr(Bl, tau, k) if want_q else None
        else:
            d = np.sign(np.diag(R[:, :k]))
            d[d == 0] = 1.0
            R = np.asfortranarray(R * d[:, None])      # the device Q has diag(R) > 0
            Q = (ctx.fetch(B.F_Q, (mB, k)) if k > 0 else np.zeros((mB, 0), order="F")) if want_q else None
        if not want_r:
            R = None
    if want_q and want_r and not want_t:
        return B.PartialQR(Q, R, p, rounds)
    return B.PQRFactors(Q, R, p, k, T if want_t else None, rounds)


@_eltype
def prange(A, opts: Optional[LRAOptions] = None, trans: str = "n", rand=None, rand2=None,
           ctx: Optional[Context] = None, **kw):
    """prange(trans, A, opts; kw...) -> Q (src/prange.jl:14-62): an orthonormal basis of the range of A (trans "n"), of
    A' ("c") or of both ("b").  rand / rand2: the per-round random inputs of the (first / second) right-hand sketch."""
    if trans not in ("n", "c", "b"):
        raise ValueError("trans")                                   # prange_chktrans
    ctx = ctx or default_context()
    o = _opts(opts, kw)
    pA, m, n, lda, keepA = mat_arg(A)
    rp, rp2 = _RandPack(rand), _RandPack(rand2)
    co = o.to_c()
    rc = lib.bra_prange_f64(ctx.handle, trans.encode(), m, n, pA, lda, C.byref(co), C.byref(rp.c), C.byref(rp2.c))
    if rc == -3:
        raise ValueError(lib.bra_last_error(ctx.handle).decode())   # checksquare -> DimensionMismatch
    ctx.check(rc)
    inf, rounds, steps = _rounds(ctx)
    k, M = int(inf.k), int(inf.m)
    return ctx.fetch(B.F_Q, (M, k)) if k > 0 else np.zeros((M, 0), order="F")


def pqr(A, *args, **kw):
    """pqr(...) -> (Q, R, p) (src/pqr.jl:312-318)."""
    kw.setdefault("pqrfact_retval", "qr")
    F = pqrfact(A, *args, **kw)
    return F.Q, F.R, F.p


def psvdfact_device(A, opts: Optional[LRAOptions] = None, rand=None, ctx: Optional[Context] = None, **kw):
    o = _opts(opts, kw)
    ctx = ctx or default_context()
    pA, m, n, lda, keepA = mat_arg(A)
    rp = _RandPack(rand, o, max(m, n))                  # trans = :n if m >= n else :c (src/psvd.jl:242,256)
    co = o.to_c()
    ctx.check(lib.bra_psvdfact_f64(ctx.handle, m, n, pA, lda, C.byref(co), C.byref(rp.c)))
    return ctx.info()


@_eltype
def psvdfact(A, opts: Optional[LRAOptions] = None, rand=None, ctx: Optional[Context] = None, out=None, **kw):
    """psvdfact(A, opts; kw...) -> PartialSVD(U, S, Vt) (src/psvd.jl:238-272).

    `out` (optional) = (U, S, Vt) caller-owned column-major host buffers at least m x k, k and k x n large (the C ABI's
    ownership model: results go into caller memory); with pinned buffers the read-back runs at the PCIe rate and the
    returned factors are views of them."""
    ctx = ctx or default_context()
    pA, m, n, lda, keepA = mat_arg(A)
    oU, oS, oV = out if out is not None else (None, None, None)
    if out is not None:
        # register the caller's buffers: each factor is copied out inside the call, as soon as it exists
        # (bra_psvd_set_outputs); what did not fit or was not registered is fetched below
        def _ok(a, nd):
            return (a is not None and a.dtype == np.float64 and a.ndim == nd and a.strides[0] == a.itemsize
                    and a.flags.writeable)
        uo = oU if _ok(oU, 2) and oU.shape[0] >= m else None
        so = oS if _ok(oS, 1) else None
        vo = oV if _ok(oV, 2) and oV.shape[1] >= n else None
        lib.bra_psvd_set_outputs(
            ctx.handle, C.c_void_p(uo.ctypes.data if uo is not None else None),
            (uo.strides[1] // 8 if uo is not None and uo.shape[1] > 1 else max(m, 1)), (uo.shape[1] if uo is not None else 0),
            C.c_void_p(so.ctypes.data if so is not None else None), (so.shape[0] if so is not None else 0),
            C.c_void_p(vo.ctypes.data if vo is not None else None),
            (vo.strides[1] // 8 if vo is not None and vo.shape[1] > 1 else (vo.shape[0] if vo is not None else 0)))
    psvdfact_device(A, opts, rand, ctx, **kw)
    done = int(lib.bra_psvd_outputs_done(ctx.handle)) if out is not None else 0
    inf, rounds, steps = _rounds(ctx)
    ks = int(inf.ksvd)
    if ks == 0:
        return B.PartialSVD(np.zeros((m, 0)), np.zeros(0), np.zeros((0, n)), int(inf.k), rounds)
    U = oU[:m, :ks] if done & 1 else ctx.fetch(B.F_U, (m, ks), out=oU)
    S = oS[:ks] if done & 2 else ctx.fetch(B.F_S, (ks,), out=oS)
    Vt = oV[:ks, :n] if done & 4 else ctx.fetch(B.F_VT, (ks, n), out=oV)
    return B.PartialSVD(U, S, Vt, int(inf.k), rounds)


@_eltype
def pheigfact(A, opts: Optional[LRAOptions] = None, rand=None, ctx: Optional[Context] = None, **kw):
    """pheigfact(A, opts; kw...) -> PartialHermEigen(values, vectors) (src/pheig.jl:276-296); A real symmetric,
    otherwise ValueError("matrix must be Hermitian") (:279)."""
    o = _opts(opts, kw)
    ctx = ctx or default_context()
    pA, m, n, lda, keepA = mat_arg(A)
    if m != n:
        raise ValueError("matrix is not square")            # checksquare
    rp = _RandPack(rand, o, n)
    co = o.to_c()
    rc = lib.bra_pheigfact_f64(ctx.handle, n, pA, lda, C.byref(co), C.byref(rp.c))
    if rc == -3:
        raise ValueError("matrix must be Hermitian")          # error("matrix must be Hermitian"), src/pheig.jl:279
    ctx.check(rc)
    inf, rounds, steps = _rounds(ctx)
    kk = int(inf.ksvd)
    if kk == 0:
        return B.PartialHermEigen(np.zeros(0), np.zeros((n, 0)), int(inf.k), rounds)
    return B.PartialHermEigen(ctx.fetch(B.F_S, (kk,)), ctx.fetch(B.F_U, (n, kk)), int(inf.k), rounds)


def pheig(A, *args, **kw):
    """pheig(A, ...) -> (values, vectors) (src/pheig.jl:316-319)."""
    F = pheigfact(A, *args, **kw)
    return F.values, F.vectors


def pheigvals(A, *args, **kw):
    """pheigvals(A, ...) (src/pheig.jl:298-311)."""
    return pheigfact(A, *args, **kw).values


def psvd(A, *args, **kw):
    """psvd(A, ...) -> (U, S, V) with V = Vt' (src/psvd.jl:296-299)."""
    F = psvdfact(A, *args, **kw)
    return F.U, F.S, F.Vt.T


@_eltype
def psvdvals(A, opts: Optional[LRAOptions] = None, rand=None, ctx: Optional[Context] = None, **kw):
    """psvdvals(A, opts; kw...) (src/psvd.jl:274-290): the singular values of psvdfact without its vectors
    (bra_psvdvals_f64: no explicit Q, no U / Vt products)."""
    o = _opts(opts, kw)
    ctx = ctx or default_context()
    pA, m, n, lda, keepA = mat_arg(A)
    rp = _RandPack(rand, o, max(m, n))
    co = o.to_c()
    ctx.check(lib.bra_psvdvals_f64(ctx.handle, m, n, pA, lda, C.byref(co), C.byref(rp.c)))
    ks = int(ctx.info().ksvd)
    return ctx.fetch(B.F_S, (ks,)) if ks else np.zeros(0)


def _dev(a):
    """Column-major device copy of a host array (torch is the device allocator of this host mirror); cuda tensors and
    DeviceMatrix pass through.  Returns (DeviceMatrix, keepalive)."""
    import torch
    if isinstance(a, DeviceMatrix):
        return a, a
    if hasattr(a, "is_cuda"):
        return DeviceMatrix.from_torch(a), a
    h = np.asarray(a, dtype=np.float64)
    if h.ndim == 1:
        h = h.reshape(-1, 1)
    t = torch.from_numpy(np.ascontiguousarray(h.T)).cuda()         # row-major (n x m) == column-major m x n
    return DeviceMatrix(t.data_ptr(), h.shape[0], h.shape[1], max(h.shape[0], 1), keep=t), t


def snormdiff(A, left=None, right=None, opts: Optional[LRAOptions] = None, x0=None, ctx: Optional[Context] = None, **kw):
    """snormdiff(A, F) = snorm(A - F) for F = left @ right (src/snorm.jl:14-53) on the device; left = right = None gives
    snorm(A).  `left` may also be a factorization of this package: PartialSVD, PartialQR (with its permutation), or
    PartialHermEigen.  x0: the reference's crandn(n) start vector (parity); default: device Philox keyed by opts.seed."""
    o = _opts(opts, kw)
    ctx = ctx or default_context()
    if isinstance(left, B.PartialSVD):
        left, right = left.U * left.S, left.Vt
    elif isinstance(left, B.PartialHermEigen):
        left, right = left.vectors * left.values, left.vectors.T
    elif isinstance(left, B.PartialQR):
        R = np.zeros_like(left.R)
        R[:, left.p - 1] = left.R
        left, right = left.Q, R
    dA, kA = _dev(A)
    k = 0
    pL = pR = C.c_void_p(0)
    ldl = ldr = 1
    keep = []
    if left is not None:
        dL, kL = _dev(left)
        dR, kR = _dev(right)
        if dL.m != dA.m or dR.n != dA.n or dL.n != dR.m:
            raise ValueError("DimensionMismatch")
        k, pL, pR, ldl, ldr = dL.n, C.c_void_p(dL.ptr), C.c_void_p(dR.ptr), dL.ld, dR.ld
        keep += [kL, kR]
    px = C.c_void_p(0)
    if x0 is not None:
        if np.size(x0) != dA.n:
            raise ValueError("DimensionMismatch: x0")
        dx, kx = _dev(np.asarray(x0, dtype=np.float64).reshape(-1, 1))
        px = C.c_void_p(dx.ptr)
        keep.append(kx)
    co = o.to_c()
    res, nit = C.c_double(0.0), C.c_int64(0)
    ctx.check(lib.bra_snorm_f64(ctx.handle, dA.m, dA.n, C.c_void_p(dA.ptr), dA.ld, k, pL, ldl, pR, ldr, C.byref(co),
                                int(o.snorm_niter), px, C.byref(res), C.byref(nit)))
    return float(res.value)


def snorm(A, opts: Optional[LRAOptions] = None, x0=None, ctx: Optional[Context] = None, **kw):
    """snorm(A, opts; kw...) (src/snorm.jl:14-44)."""
    return snormdiff(A, None, None, opts, x0, ctx, **kw)


def probe_fp64_peak(ctx: Optional[Context] = None) -> dict:
    ctx = ctx or default_context()
    out = (C.c_double * 8)()
    ctx.check(lib.bra_probe_fp64_peak(ctx.handle, out))
    return {"dmma_m8n8k4": out[0], "dfma": out[1], "dmma_m16n8k4": out[2], "dmma_m16n8k8": out[3],
            "dmma_m16n8k16": out[4]}


def probe_exchange_latency(ctas: int = 148, iters: int = 2000, ctx: Optional[Context] = None) -> float:
    ctx = ctx or default_context()
    us = C.c_double(0)
    ctx.check(lib.bra_probe_exchange_latency(ctx.handle, ctas, iters, C.byref(us)))
    return us.value
