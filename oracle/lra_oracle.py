"""CPU oracle for the sketch -> early-terminating QRCP -> ID solve -> psvd path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it, and only as the checker / baseline.

What it is: a restatement, in numpy, of the reference's Julia driver code
(all citations are into /root/reference/), which calls -- through ctypes --
the *same LAPACK/BLAS routines* the reference reaches through ``ccall``:
``dlaqps`` (src/lapack.jl:117-139), ``dgemm`` (src/sketch.jl:93), ``dtrsm``
(src/pqr.jl:441), ``dgeqrf/dorgqr`` (src/pqr.jl:298,302), ``dgesdd``
(src/psvd.jl:245).  The library is scipy's bundled OpenBLAS
(``scipy.libs/libscipy_openblas-*.so``, LP64, symbols ``scipy_<name>_``).

Parity status: **unpinned against the Julia reference itself** -- the image
has no Julia, the reference ships no golden vectors, and
its tests only pin residual inequalities.  What *is* pinned (tests/test_oracle.py):
  (i)   the reference's own test inequalities on 128x64 Fourier(real part)
        matrices for every sketch kind (test/id.jl:27-32, test/pqr.jl:27-30,
        test/psvd.jl:28-32);
  (ii)  the README's known answers: Hilbert-1024 rank 26/27 at default rtol
        and ~22 at rtol=1e-12 (README.md:107,122);
  (iii) the from-scratch ``dlaqps_restated`` below == the real ``dlaqps``
        (pivots, block lengths, R, tau);
  (iv)  the SRFT restatement == a plain DFT.

Randomness: the reference draws from Julia's task-local RNG.  Every function
here takes the random inputs explicitly (Omega / (d, idx) / (perm, s) / r), in
the order the reference draws them, so that the GPU path and the oracle can be
fed identical inputs (SURVEY.md section 8b "Randomness").
"""
from __future__ import annotations

import ctypes
import glob
import os
from dataclasses import dataclass, field, replace
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------
# LAPACK / BLAS binding (scipy's OpenBLAS, LP64)
# --------------------------------------------------------------------------

_c_int = ctypes.c_int
_c_dbl = ctypes.c_double
_pi = ctypes.POINTER(_c_int)
_pd = ctypes.POINTER(_c_dbl)


def _find_openblas() -> str:
    import scipy

    root = os.path.join(os.path.dirname(scipy.__file__), os.pardir, "scipy.libs")
    hits = sorted(glob.glob(os.path.join(root, "libscipy_openblas*.so")))
    if not hits:
        raise RuntimeError("scipy's bundled OpenBLAS not found")
    return hits[0]


_lib = ctypes.CDLL(_find_openblas())


def set_blas_threads(n: int) -> None:
    _lib.scipy_openblas_set_num_threads(int(n))


def get_blas_threads() -> int:
    return int(_lib.scipy_openblas_get_num_threads())


def _ref(x: int):
    return ctypes.byref(_c_int(int(x)))


def _dptr(a: np.ndarray, offset_elems: int = 0):
    return ctypes.cast(a.ctypes.data + 8 * offset_elems, _pd)


def _iptr(a: np.ndarray, offset_elems: int = 0):
    return ctypes.cast(a.ctypes.data + 4 * offset_elems, _pi)


def _fortran(a: np.ndarray) -> np.ndarray:
    return np.asfortranarray(a, dtype=np.float64)


def dgemm(A: np.ndarray, B: np.ndarray, transa: bool = False, transb: bool = False) -> np.ndarray:
    """C = op(A) op(B) through the real BLAS dgemm (src/sketch.jl:93 -> mul!)."""
    A = _fortran(A)
    B = _fortran(B)
    m = A.shape[1] if transa else A.shape[0]
    k = A.shape[0] if transa else A.shape[1]
    n = B.shape[0] if transb else B.shape[1]
    if k != (B.shape[1] if transb else B.shape[0]):
        raise ValueError(f"DimensionMismatch: op(A) is {m}x{k}, op(B) has {B.shape[1] if transb else B.shape[0]} rows")
    C = np.zeros((m, n), order="F")
    one, zero = _c_dbl(1.0), _c_dbl(0.0)
    _lib.scipy_dgemm_(
        ctypes.c_char_p(b"T" if transa else b"N"), ctypes.c_char_p(b"T" if transb else b"N"),
        _ref(m), _ref(n), _ref(k), ctypes.byref(one), _dptr(A), _ref(max(1, A.shape[0])),
        _dptr(B), _ref(max(1, B.shape[0])), ctypes.byref(zero), _dptr(C), _ref(max(1, m)),
        ctypes.c_size_t(1), ctypes.c_size_t(1))
    return C


def dnrm2(x: np.ndarray) -> float:
    x = np.ascontiguousarray(x, dtype=np.float64)
    _lib.scipy_dnrm2_.restype = _c_dbl
    return float(_lib.scipy_dnrm2_(_ref(x.size), _dptr(x), _ref(1)))


def dlaqps_real(offset: int, nb: int, A: np.ndarray, col0: int, jpvt: np.ndarray,
                tau: np.ndarray, vn1: np.ndarray, vn2: np.ndarray) -> int:
    """One call of the real LAPACK ``dlaqps`` exactly as src/pqr.jl:397-400 +
    src/lapack.jl:129-136 pass it: sub-matrix A[:, col0:], N = n - col0,
    OFFSET = col0, JPVT/TAU/VN1/VN2 offset by col0, LDF = N.  Returns KB."""
    m, n = A.shape
    N = n - col0
    kb = _c_int(0)
    auxv = np.zeros(max(nb, 1))
    F = np.zeros(max(N * nb, 1))
    _lib.scipy_dlaqps_(
        _ref(m), _ref(N), _ref(offset), _ref(nb), ctypes.byref(kb),
        _dptr(A, col0 * m), _ref(max(1, m)), _iptr(jpvt, col0), _dptr(tau, col0),
        _dptr(vn1, col0), _dptr(vn2, col0), _dptr(auxv), _dptr(F), _ref(max(1, N)))
    return int(kb.value)


# --------------------------------------------------------------------------
# dlaqps restated from the published LAPACK algorithm (SURVEY.md section 8 a5')
# --------------------------------------------------------------------------

TOL3Z = float(np.sqrt(2.0 ** -53))  # sqrt(DLAMCH('Epsilon')), LAPACK's eps is the rounding unit


def _dlarfg(alpha: float, x: np.ndarray) -> Tuple[float, float]:
    """LAPACK dlarfg: returns (beta, tau) and scales x in place (Appendix C)."""
    if x.size == 0:
        return alpha, 0.0
    xnorm = float(np.sqrt(np.dot(x, x)))
    if xnorm == 0.0:
        return alpha, 0.0
    beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
    tau = (beta - alpha) / beta
    x *= 1.0 / (alpha - beta)
    return float(beta), float(tau)


def dlaqps_restated(offset: int, nb: int, A: np.ndarray, col0: int, jpvt: np.ndarray,
                    tau: np.ndarray, vn1: np.ndarray, vn2: np.ndarray) -> int:
    """From-scratch restatement of LAPACK ``dlaqps`` (same argument convention
    as :func:`dlaqps_real`).  Used to pin the semantics the CUDA kernel must
    reproduce: first-max pivot, norm hand-over on swap, LAWN-176 downdate with
    tol3z = sqrt(2^-53), early block end on the first flagged column, norm
    recomputation after the trailing update."""
    m, n = A.shape
    N = n - col0
    S = A[:, col0:]          # view
    jp = jpvt[col0:]
    v1 = vn1[col0:]
    v2 = vn2[col0:]
    tv = tau[col0:]
    lastrk = min(m, N + offset)
    F = np.zeros((N, nb))
    flagged: List[int] = []
    k = 0
    while k < nb and not flagged:
        rk = offset + k
        pvt = k + int(np.argmax(v1[k:]))          # idamax: first maximum
        if pvt != k:
            S[:, [pvt, k]] = S[:, [k, pvt]]
            F[[pvt, k], :k] = F[[k, pvt], :k]
            jp[pvt], jp[k] = jp[k], jp[pvt]
            v1[pvt] = v1[k]
            v2[pvt] = v2[k]
        if k > 0:
            S[rk:, k] -= S[rk:, :k] @ F[k, :k]
        if rk < m - 1:
            beta, t = _dlarfg(S[rk, k], S[rk + 1:, k])
        else:
            beta, t = S[rk, k], 0.0
        tv[k] = t
        akk = beta
        S[rk, k] = 1.0
        if k < N - 1:
            F[k + 1:, k] = t * (S[rk:, k + 1:].T @ S[rk:, k])
        F[:k + 1, k] = 0.0
        if k > 0:
            aux = -t * (S[rk:, :k].T @ S[rk:, k])
            F[:, k] += F[:, :k] @ aux
        if k < N - 1:
            S[rk, k + 1:] -= F[k + 1:, :k + 1] @ S[rk, :k + 1]
        if rk < lastrk - 1:
            for j in range(k + 1, N):
                if v1[j] != 0.0:
                    temp = abs(S[rk, j]) / v1[j]
                    temp = max(0.0, (1.0 + temp) * (1.0 - temp))
                    temp2 = temp * (v1[j] / v2[j]) ** 2
                    if temp2 <= TOL3Z:
                        flagged.append(j)
                    else:
                        v1[j] *= np.sqrt(temp)
        S[rk, k] = akk
        k += 1
    kb = k
    rk = offset + kb
    if kb < min(N, m - offset):
        S[rk:, kb:] -= S[rk:, :kb] @ F[kb:, :kb].T
    for j in flagged:
        v1[j] = float(np.sqrt(np.dot(S[rk:, j], S[rk:, j])))
        v2[j] = v1[j]
    return kb


# --------------------------------------------------------------------------
# Options (src/LowRankApprox.jl:77-148)
# --------------------------------------------------------------------------

EPS = float(np.finfo(np.float64).eps)


@dataclass
class LRAOptions:
    atol: float = 0.0
    maxdet_niter: int = -1
    maxdet_tol: float = -1.0
    nb: int = 32
    pheig_orthtol: float = float(np.sqrt(EPS))
    pqrfact_retval: str = "qr"
    rank: int = -1
    rtol: float = 5 * EPS
    sketch: str = "randn"
    sketch_randn_niter: int = 0
    sketchfact_adap: bool = True
    sketchfact_randn_samp: Callable[[int], int] = field(default=lambda n: n + 8)
    sketchfact_srft_samp: Callable[[int], int] = field(default=lambda n: n + 8)
    sketchfact_sub_samp: Callable[[int], int] = field(default=lambda n: 4 * n + 8)
    snorm_niter: int = 32
    verb: bool = True

    def copy(self, **kw) -> "LRAOptions":
        return replace(self, **kw)


def chkopts(opts: LRAOptions) -> None:
    """src/LowRankApprox.jl:133-141."""
    if not opts.atol >= 0:
        raise ValueError("atol")
    if not opts.nb > 0:
        raise ValueError("nb")
    if not opts.rtol >= 0:
        raise ValueError("rtol")
    if opts.sketch not in ("none", "randn", "sprn", "srft", "sub"):
        raise ValueError("sketch")
    opts.pqrfact_retval = opts.pqrfact_retval.lower()


# --------------------------------------------------------------------------
# Test matrices (src/matrixlib.jl)
# --------------------------------------------------------------------------

def matrixlib_hilb(m: int, n: Optional[int] = None) -> np.ndarray:
    """A[i,j] = 1/(i+j-1), 1-based (src/matrixlib.jl:39-47)."""
    n = m if n is None else n
    i = np.arange(1, m + 1)[:, None]
    j = np.arange(1, n + 1)[None, :]
    return np.asfortranarray(1.0 / (i + j - 1))


def matrixlib_cauchy(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """A[i,j] = 1/(x_i - y_j) (src/matrixlib.jl:13-23)."""
    return np.asfortranarray(1.0 / (np.asarray(x)[:, None] - np.asarray(y)[None, :]))


def matrixlib_fourier(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """A[i,j] = exp(-2 pi i x_i y_j) (src/matrixlib.jl:25-37); complex."""
    return np.asfortranarray(np.exp(-2j * np.pi * np.asarray(x)[:, None] * np.asarray(y)[None, :]))


def decaying_matrix(m: int, n: int, r: int, decades: float, jdiv: float, seed: int) -> np.ndarray:
    """BASELINE config 2/3 matrix: U diag(sigma) V^T, sigma_j = 10^(-decades*j/jdiv),
    U, V = thin QR of seeded Gaussians."""
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((m, r)))
    V, _ = np.linalg.qr(rng.standard_normal((n, r)))
    s = 10.0 ** (-decades * np.arange(r) / jdiv)
    return np.asfortranarray((U * s) @ V.T)


# --------------------------------------------------------------------------
# Sketches, (:left, :n) and (:left, :c) forms, explicit random inputs
# --------------------------------------------------------------------------

def sketch_randn(A: np.ndarray, Omega: np.ndarray, trans: str = "n") -> np.ndarray:
    """B = Omega * A (src/sketch.jl:91-94,129-139) or Omega * A' (:101-104).
    Omega is order x size(op(A),1), drawn by crandn (src/util.jl:4)."""
    return dgemm(Omega, A, transb=(trans == "c"))


def orthrows(B: np.ndarray) -> np.ndarray:
    """orthrows!(A; thin=false) (src/util.jl:87-97): LAPACK dgelqf + dorglq, rows beyond min(m, n) zeroed."""
    B = _fortran(B).copy(order="F")
    m, n = B.shape
    k = min(m, n)
    if k == 0:
        return B
    tau = np.zeros(k)
    lwork = max(1, 64 * max(m, 1))
    work = np.zeros(lwork)
    info = _c_int(0)
    _lib.scipy_dgelqf_(_ref(m), _ref(n), _dptr(B), _ref(max(1, m)), _dptr(tau), _dptr(work), _ref(lwork), ctypes.byref(info))
    _lib.scipy_dorglq_(_ref(k), _ref(n), _ref(k), _dptr(B), _ref(max(1, m)), _dptr(tau), _dptr(work), _ref(lwork),
                       ctypes.byref(info))
    B[k:, :] = 0.0
    return B


def sketch_randn_power(A: np.ndarray, Omega: np.ndarray, trans: str, niter: int) -> np.ndarray:
    """sketch_randn_ln / sketch_randn_lc with sketch_randn_niter power steps (src/sketch.jl:129-151, 152-173):
    Bp = Omega op(A); repeat: orthonormalise the rows, multiply by op(A)', (unless A is Hermitian) orthonormalise
    again and multiply by op(A)."""
    Bp = sketch_randn(A, Omega, trans)
    isherm = A.shape[0] == A.shape[1] and np.array_equal(A, A.T)
    other = "c" if trans == "n" else "n"
    for _ in range(niter):
        Bp = orthrows(Bp)
        Bq = sketch_randn(A, Bp, other)          # Bp * op(A)'
        if isherm:
            Bp = Bq
        else:
            Bq = orthrows(Bq)
            Bp = sketch_randn(A, Bq, trans)      # Bq * op(A)
    return Bp


def sketch_sub(A: np.ndarray, r: np.ndarray, trans: str = "n") -> np.ndarray:
    """B[i,:] = op(A)[r_i,:], r 1-based with replacement (src/sketch.jl:248-257,268-279)."""
    r0 = np.asarray(r, dtype=np.int64) - 1
    if trans == "n":
        return np.asfortranarray(A[r0, :])
    return np.asfortranarray(A[:, r0].T)


def sprn_counts(m: int, order: int) -> np.ndarray:
    """p_i = fld(m - i, order) + 1, i = 1..order (src/sketch.jl:577)."""
    i = np.arange(1, order + 1)
    return (m - i) // order + 1


def sketch_sprn(A: np.ndarray, order: int, perm: np.ndarray, s: np.ndarray, trans: str = "n") -> np.ndarray:
    """Row i of B = sum_{l<=p_i} s_l * op(A)[perm[idx+l], :] (src/sketch.jl:571-589).
    perm is randperm(m) (1-based); s is the concatenation of the per-row
    crandn(p_i) draws (sum p_i = m entries)."""
    Aop = A if trans == "n" else A.T
    m, n = Aop.shape
    p = sprn_counts(m, order)
    perm0 = np.asarray(perm, dtype=np.int64) - 1
    B = np.zeros((order, n), order="F")
    idx = 0
    for i in range(order):
        rows = perm0[idx:idx + p[i]]
        # the reference accumulates l = 1..p sequentially per entry
        acc = np.zeros(n)
        for l in range(p[i]):
            acc += s[idx + l] * Aop[rows[l], :]
        B[i, :] = acc
        idx += p[i]
    return B


def srft_l(m: int, order: int) -> int:
    """Largest l <= order dividing m (src/sketch.jl:354-357)."""
    l = order
    while m % l > 0:
        l -= 1
    return l


def _r2hc(X: np.ndarray) -> np.ndarray:
    """FFTW R2HC along dim 0: [r0, r1, ..., r_{l/2}, i_{ceil(l/2)-1}, ..., i_1].
    FFTW is not in the image; numpy's pocketfft computes the same DFT."""
    l = X.shape[0]
    Z = np.fft.rfft(X, axis=0)
    H = np.empty_like(X)
    H[: l // 2 + 1] = Z.real
    for c in range(1, (l + 1) // 2):
        H[l - c] = Z[c].imag
    return H


def sketch_srft(A: np.ndarray, order: int, d: np.ndarray, idx: np.ndarray, trans: str = "n") -> np.ndarray:
    """Real SRFT (src/sketch.jl:353-450,474-484), restated including its quirks:
    (Re, Im) pairs consume every other idx entry, and the `in == 0 || in == nnyq`
    test (:425) is always false so only `i == k` yields a lone real row."""
    Aop = A if trans == "n" else A.T
    m, n = Aop.shape
    k = order
    l = srft_l(m, k)
    mp = m // l
    idx = np.asarray(idx, dtype=np.int64)
    B = np.zeros((k, n), order="F")
    cnyq = l // 2
    # X[j,kk] = d[i] x[i], i = j*mp + kk  (srft_reshape!, :378-385)
    X = (Aop * np.asarray(d, dtype=np.float64)[:, None]).reshape(l, mp, n)
    H = _r2hc(X)                           # l x mp x n
    wn = np.exp(-2j * np.pi / m)
    wm = np.exp(-2j * np.pi / mp)
    i = 0
    while i < k:
        f = int(idx[i]) - 1
        row = f // l
        col_ = f % l
        w = wm ** row * wn ** col_
        cswap = col_ > cnyq
        ia = l - col_ if cswap else col_
        ib = (l - ia) if col_ > 0 else 0
        a = H[ia]                          # mp x n
        b = np.zeros_like(a) if (ib == 0 or ib == ia) else H[ib]
        if cswap:
            b = -b
        # s = w^j by repeated multiplication, as the reference does
        s = np.ones(mp, dtype=np.complex128)
        for j in range(1, mp):
            s[j] = s[j - 1] * w
        z = (s[:, None] * (a + 1j * b)).sum(axis=0)
        if i == k - 1:
            B[i, :] = z.real
        else:
            B[i, :] = z.real
            B[i + 1, :] = z.imag
            i += 1
        i += 1
    return B


# --------------------------------------------------------------------------
# Early-terminating QRCP (src/pqr.jl:348-418) and post-processing (:420-442)
# --------------------------------------------------------------------------

@dataclass
class QRCPTrace:
    kb: List[int] = field(default_factory=list)      # block lengths laqps returned
    ptol: float = 0.0
    steps: int = 0                                   # pivot steps actually executed


def geqp3_adap(B: np.ndarray, opts: LRAOptions, laqps=dlaqps_real,
               trace: Optional[QRCPTrace] = None) -> Tuple[np.ndarray, np.ndarray, int]:
    """In-place on B (must be F-ordered float64).  Returns (jpvt 1-based, tau, k).
    Follows geqp3_adap! / geqp3_adap_main! (src/pqr.jl:348-418) line by line."""
    assert B.flags.f_contiguous and B.dtype == np.float64
    m, n = B.shape
    jpvt = np.arange(1, n + 1, dtype=np.int32)
    l = min(m, n)
    k = l if (opts.rank < 0 or opts.rank > l) else opts.rank
    tau = np.zeros(max(k, 1))
    if k == 0:
        return jpvt.astype(np.int64), tau[:0], 0
    nb = min(opts.nb, k)
    vn1 = np.array([dnrm2(B[:, j]) for j in range(n)]) if m >= 32 else \
        np.sqrt(np.einsum("ij,ij->j", B, B))
    # Julia's norm(view) uses BLAS nrm2 for length >= 32 and a generic loop below
    vn2 = vn1.copy()
    maxnrm = float(vn1.max()) if n else 0.0
    ptol = max(opts.atol, opts.rtol * maxnrm)
    if trace is not None:
        trace.ptol = ptol
    j = 0                                # 0-based
    while j < k:
        jb = min(nb, k - j)
        fjb = laqps(j, jb, B, j, jpvt, tau, vn1, vn2)
        if trace is not None:
            trace.kb.append(fjb)
            trace.steps += fjb
        jn = j + fjb
        if abs(B[jn - 1, jn - 1]) <= ptol:
            for i in range(j, jn):
                if abs(B[i, i]) <= ptol:
                    return jpvt.astype(np.int64), tau[:k], i
        j = jn
    return jpvt.astype(np.int64), tau[:k], k


def dtrsm_upper(R11: np.ndarray, R12: np.ndarray) -> np.ndarray:
    """T = R11^{-1} R12 through the real BLAS dtrsm (L,U,N,N) (src/pqr.jl:438-442)."""
    k = R11.shape[0]
    T = _fortran(R12).copy(order="F")
    R11 = _fortran(R11)
    if k == 0 or T.shape[1] == 0:
        return T
    one = _c_dbl(1.0)
    _lib.scipy_dtrsm_(ctypes.c_char_p(b"L"), ctypes.c_char_p(b"U"), ctypes.c_char_p(b"N"),
                      ctypes.c_char_p(b"N"), _ref(k), _ref(T.shape[1]), ctypes.byref(one),
                      _dptr(R11), _ref(max(1, k)), _dptr(T), _ref(max(1, k)),
                      ctypes.c_size_t(1), ctypes.c_size_t(1), ctypes.c_size_t(1), ctypes.c_size_t(1))
    return T


@dataclass
class PQRFactors:
    """PartialQRFactors (src/pqr.jl:14-20)."""
    Q: Optional[np.ndarray]
    R: Optional[np.ndarray]
    p: np.ndarray
    k: int
    T: Optional[np.ndarray]
    rounds: List[Tuple[int, int]] = field(default_factory=list)   # (order, k_t) per adaptive round
    traces: List[QRCPTrace] = field(default_factory=list)


def dorgqr(Bk: np.ndarray, tau: np.ndarray) -> np.ndarray:
    Q = _fortran(Bk).copy(order="F")
    m, k = Q.shape
    if k == 0:
        return Q
    info = _c_int(0)
    lwork = max(1, 64 * k)
    work = np.zeros(lwork)
    _lib.scipy_dorgqr_(_ref(m), _ref(k), _ref(k), _dptr(Q), _ref(max(1, m)), _dptr(np.ascontiguousarray(tau)),
                       _dptr(work), _ref(lwork), ctypes.byref(info))
    return Q


def findmaxabs(x: np.ndarray) -> Tuple[float, int]:
    """src/util.jl:13-23: largest |x[i]| in column-major order; among equal values the LAST index wins
    (`t < m && continue`).  Returns (value, 0-based linear index)."""
    a = np.abs(np.asarray(x)).ravel(order="F")
    if a.size == 0:
        return 0.0, -1
    m = a.max()
    return float(m), int(np.flatnonzero(a == m)[-1])


def maxdet_swapcols(R: Optional[np.ndarray], p: np.ndarray, T: np.ndarray, opts: LRAOptions, retr: bool) -> int:
    """src/pqr.jl:444-501 for the factors this path returns (p, T and, if requested, R1): while max|T| > 1 + tol,
    swap skeleton column i with redundant column j and update T by Sherman-Morrison (maxdet_update!, :481-501).
    The final re-triangularisation of R1 / update of Q (:467-476) is done by the caller (pqrback_postproc).
    Returns the number of swaps."""
    k, nk = T.shape
    niter = 0
    while True:
        tmax, idx = findmaxabs(T)
        if tmax <= 1 + opts.maxdet_tol:
            break
        if niter == opts.maxdet_niter:
            break
        niter += 1
        i, j = idx % k, idx // k
        p[i], p[k + j] = p[k + j], p[i]
        w1 = T[:, j].copy()
        T[:, j] = 0.0
        w1[i] -= 1.0
        T[i, j] = 1.0
        w2 = T[i, :].copy()
        alpha = -1.0 / (1.0 + w1[i])
        T += alpha * np.outer(w1, w2)              # BLAS.ger!
        if retr and R is not None:
            R[:, i] += R[:, :k] @ w1
    return niter


def pqrback_postproc(B: np.ndarray, p: np.ndarray, tau: np.ndarray, k: int, opts: LRAOptions) -> PQRFactors:
    """src/pqr.jl:420-436."""
    retq = "q" in opts.pqrfact_retval
    retr = "r" in opts.pqrfact_retval
    rett = "t" in opts.pqrfact_retval
    maxdet = 0 < k < B.shape[1] and opts.maxdet_tol >= 0
    Q = dorgqr(B[:, :k], tau[:k]) if retq else None
    R = np.triu(B[:k, :]) if (retr or rett or maxdet) else None
    T = dtrsm_upper(R[:, :k], R[:, k:]) if (rett or maxdet) else None
    if maxdet:
        nsw = maxdet_swapcols(R, p, T, opts, retr)
        if nsw > 0 and (retq or retr):
            # src/pqr.jl:467-476: re-triangularise R1 (only retr keeps it updated: with "q" alone R1 is still
            # triangular, the QR below is the identity and Q ignores the swaps)
            Qf, Rf = qr_thin(R[:, :k])
            if retq:
                Q = Q @ Qf
            if retr:
                R[:, :k] = Rf
                R[:, k:] = Rf @ T
    return PQRFactors(Q, R if retr else None, p, k, T)


# --------------------------------------------------------------------------
# Adaptive drivers (src/sketch.jl:223-240, 313-330, 545-562, 674-690)
# --------------------------------------------------------------------------

def sketch_order(kind: str, n: int, opts: LRAOptions) -> int:
    if kind == "randn":
        return opts.sketchfact_randn_samp(n)
    if kind == "srft":
        return opts.sketchfact_srft_samp(n)
    if kind == "sub":
        return opts.sketchfact_sub_samp(n)
    if kind == "sprn":
        return n
    raise ValueError(kind)


class RandomInputs:
    """Supplies, per adaptive round, the random inputs the reference would draw.
    Default: numpy Generator keyed by (seed, round).  Tests pass the very same
    object state to the GPU path by recording what was drawn (``self.drawn``)."""

    def __init__(self, seed: int = 0):
        self.seed = seed
        self.drawn: List[dict] = []

    def draw(self, kind: str, rnd: int, order: int, m: int) -> dict:
        rng = np.random.default_rng([self.seed, rnd])
        if kind == "randn":
            out = {"Omega": np.asfortranarray(rng.standard_normal((order, m)))}
        elif kind == "sub":
            out = {"r": rng.integers(1, m + 1, size=order)}
        elif kind == "srft":
            d = 2.0 * (rng.random(m) > 0.5) - 1.0
            out = {"d": d, "idx": rng.integers(1, m + 1, size=order)}
        elif kind == "sprn":
            out = {"perm": rng.permutation(m) + 1, "s": rng.standard_normal(m)}
        else:
            raise ValueError(kind)
        self.drawn.append({"kind": kind, "round": rnd, "order": order, **out})
        return out


def apply_sketch(kind: str, A: np.ndarray, order: int, rin: dict, trans: str, niter: int = 0) -> np.ndarray:
    if kind == "randn":
        if niter > 0:
            return sketch_randn_power(A, rin["Omega"], trans, niter)
        return sketch_randn(A, rin["Omega"], trans)
    if kind == "sub":
        return sketch_sub(A, rin["r"], trans)
    if kind == "srft":
        return sketch_srft(A, order, rin["d"], rin["idx"], trans)
    if kind == "sprn":
        return sketch_sprn(A, order, rin["perm"], rin["s"], trans)
    raise ValueError(kind)


def sketchfact(A: np.ndarray, opts: LRAOptions, rand: RandomInputs, trans: str = "n",
               laqps=dlaqps_real, side: str = "left") -> PQRFactors:
    """sketchfact(side, trans, A, opts) (src/sketch.jl:52-66 + the four drivers).  side = "right": B = op(A) S, which in
    real arithmetic is the transpose of the left sketch of op(A)' on the same draws (the mul! forms at
    src/sketch.jl:91-110, 248-293, 474-522, 571-653 are transposes of each other), so the random inputs are recorded in
    the left-sketch format with the contracted dimension size(op(A), 2)."""
    chkopts(opts)
    kind = opts.sketch
    if side == "right":
        ft = "c" if trans == "n" else "n"

        def apply_right(kind_, A_, order_, rin_, trans_, niter_=0):
            return np.asfortranarray(apply_sketch(kind_, A_, order_, rin_, ft, niter_).T)
        sk = apply_right
        m = A.shape[1] if trans == "n" else A.shape[0]
    else:
        sk = apply_sketch
        m = A.shape[0] if trans == "n" else A.shape[1]
    rounds: List[Tuple[int, int]] = []
    traces: List[QRCPTrace] = []
    if opts.sketchfact_adap or opts.rank < 0:
        n = opts.nb
        opts_ = opts.copy(maxdet_tol=-1.0)
        rnd = 0
        while True:
            order = sketch_order(kind, n, opts)
            B = sk(kind, A, order, rand.draw(kind, rnd, order, m), trans, opts.sketch_randn_niter)
            tr = QRCPTrace()
            p, tau, k = geqp3_adap(B, opts_, laqps, tr)
            rounds.append((order, k))
            traces.append(tr)
            if k < n:
                F = pqrback_postproc(B, p, tau, k, opts)
                F.rounds, F.traces = rounds, traces
                return F
            n *= 2
            rnd += 1
    order = sketch_order(kind, opts.rank, opts) if kind != "sprn" else opts.rank
    B = sk(kind, A, order, rand.draw(kind, 0, order, m), trans, opts.sketch_randn_niter)
    tr = QRCPTrace()
    p, tau, k = geqp3_adap(B, opts, laqps, tr)
    F = pqrback_postproc(B, p, tau, k, opts)
    F.rounds, F.traces = [(order, k)], [tr]
    return F


def pqrfact_none(A: np.ndarray, opts: LRAOptions, trans: str = "n", laqps=dlaqps_real) -> PQRFactors:
    """sketch = :none (src/pqr.jl:323-327,343-346): QRCP on a copy of op(A)."""
    B = np.array(A if trans == "n" else A.T, order="F", dtype=np.float64)
    tr = QRCPTrace()
    p, tau, k = geqp3_adap(B, opts, laqps, tr)
    F = pqrback_postproc(B, p, tau, k, opts)
    F.rounds, F.traces = [(B.shape[0], k)], [tr]
    return F


# --------------------------------------------------------------------------
# Front-ends (src/id.jl:429-458, src/pqr.jl:285-340, src/psvd.jl:238-308)
# --------------------------------------------------------------------------

@dataclass
class IDPackedV:
    sk: np.ndarray        # 1-based
    rd: np.ndarray        # 1-based
    T: np.ndarray
    rounds: List[Tuple[int, int]] = field(default_factory=list)
    traces: List[QRCPTrace] = field(default_factory=list)

    @property
    def k(self) -> int:
        return len(self.sk)

    @property
    def p(self) -> np.ndarray:
        return np.concatenate([self.sk, self.rd])

    def matrix(self) -> np.ndarray:
        """Matrix(V) = [I T] P' (src/id.jl:32-46)."""
        k, n = self.k, len(self.sk) + len(self.rd)
        M = np.zeros((k, n))
        M[:, self.p - 1] = np.hstack([np.eye(k), self.T])
        return M


def idfact(A: np.ndarray, opts: LRAOptions, rand: Optional[RandomInputs] = None, trans: str = "n",
           laqps=dlaqps_real) -> IDPackedV:
    opts = opts.copy(pqrfact_retval="t")
    chkopts(opts)
    if opts.sketch == "none":
        F = pqrfact_none(A, opts, trans, laqps)
    else:
        F = sketchfact(A, opts, rand or RandomInputs(), trans, laqps)
    k = F.k
    return IDPackedV(F.p[:k].copy(), F.p[k:].copy(), F.T, F.rounds, F.traces)


def getcols(A: np.ndarray, cols1: np.ndarray, trans: str = "n") -> np.ndarray:
    c = np.asarray(cols1) - 1
    return np.asfortranarray(A[:, c] if trans == "n" else A[c, :].T)


def qr_thin(C: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Unpivoted Householder QR through real LAPACK dgeqrf + dorgqr (Julia's qr!
    uses geqrt; same reflectors, different blocking)."""
    C = _fortran(C).copy(order="F")
    m, k = C.shape
    if k == 0:
        return C, np.zeros((0, 0))
    tau = np.zeros(min(m, k))
    lwork = max(1, 64 * k)
    work = np.zeros(lwork)
    info = _c_int(0)
    _lib.scipy_dgeqrf_(_ref(m), _ref(k), _dptr(C), _ref(max(1, m)), _dptr(tau), _dptr(work), _ref(lwork),
                       ctypes.byref(info))
    R = np.triu(C[:min(m, k), :])
    Q = dorgqr(C[:, :min(m, k)], tau)
    return Q, R


@dataclass
class PartialQR:
    Q: np.ndarray
    R: np.ndarray
    p: np.ndarray
    T: Optional[np.ndarray] = None
    rounds: List[Tuple[int, int]] = field(default_factory=list)

    @property
    def k(self) -> int:
        return self.Q.shape[1]

    def matrix(self) -> np.ndarray:
        M = np.zeros((self.Q.shape[0], self.R.shape[1]))
        M[:, self.p - 1] = self.Q @ self.R
        return M


def pqrfact(A: np.ndarray, opts: LRAOptions, rand: Optional[RandomInputs] = None, trans: str = "n",
            laqps=dlaqps_real) -> PartialQR:
    """src/pqr.jl:290-307 with the default retval "qr"."""
    chkopts(opts)
    if opts.sketch == "none":
        F = pqrfact_none(A, opts.copy(pqrfact_retval="qr"), trans, laqps)
        return PartialQR(F.Q, F.R, F.p, None, F.rounds)
    V = idfact(A, opts, rand, trans, laqps)
    Q, R1 = qr_thin(getcols(A, V.sk, trans))
    R = np.hstack([R1, R1 @ V.T])            # pqrr: trmm (src/pqr.jl:330-340)
    return PartialQR(Q, R, V.p, V.T, V.rounds)


@dataclass
class PartialSVD:
    U: np.ndarray
    S: np.ndarray
    Vt: np.ndarray
    k_id: int = 0
    rounds: List[Tuple[int, int]] = field(default_factory=list)

    def matrix(self) -> np.ndarray:
        return (self.U * self.S) @ self.Vt


def psvdrank(s: np.ndarray, opts: LRAOptions) -> int:
    """src/psvd.jl:301-308."""
    k = len(s)
    if k == 0:
        return 0
    ptol = max(opts.atol, opts.rtol * s[0])
    for i in range(1, k):
        if s[i] <= ptol:
            return i
    return k


def psvdfact(A: np.ndarray, opts: LRAOptions, rand: Optional[RandomInputs] = None,
             laqps=dlaqps_real) -> PartialSVD:
    """src/psvd.jl:238-272."""
    import scipy.linalg as sla

    m, n = A.shape
    if m >= n:
        V = idfact(A, opts, rand, "n", laqps)
        Q, R = qr_thin(getcols(A, V.sk, "n"))
        W = R @ V.matrix()                                  # R*V (src/id.jl:354-360)
        Ut, s, Vt = sla.svd(W, full_matrices=False, lapack_driver="gesdd")
        k = psvdrank(s, opts)
        return PartialSVD(Q @ Ut[:, :k], s[:k], Vt[:k, :], V.k, V.rounds)
    V = idfact(A, opts, rand, "c", laqps)
    Q, R = qr_thin(getcols(A, V.sk, "c"))
    W = V.matrix().T @ R.T
    Ut, s, Vt = sla.svd(W, full_matrices=False, lapack_driver="gesdd")
    k = psvdrank(s, opts)
    return PartialSVD(Ut[:, :k], s[:k], Vt[:k, :] @ Q.T, V.k, V.rounds)


# --------------------------------------------------------------------------
# Error metric (src/snorm.jl:14-53)
# --------------------------------------------------------------------------

def snorm(matvec, rmatvec, n: int, opts: Optional[LRAOptions] = None, seed: int = 0, x0=None,
          herm: bool = False) -> float:
    """Randomised power iteration (src/snorm.jl:14-44).  x0: the start vector in place of crandn(n); herm: the
    reference's ishermitian branch (one product per iteration, s = ||x|| instead of sqrt)."""
    opts = opts or LRAOptions()
    if x0 is None:
        xn = np.random.default_rng(seed).standard_normal(n)
    else:
        xn = np.array(x0, dtype=np.float64).reshape(n)
    xnrm = np.linalg.norm(xn)
    s, t, niter = 1.0, 0.0, 0
    while s > 0 and abs(s - t) > max(opts.atol, t * opts.rtol):
        if niter == opts.snorm_niter:
            break
        niter += 1
        xn = xn / xnrm
        xn = matvec(xn) if herm else rmatvec(matvec(xn))
        xnrm = np.linalg.norm(xn)
        t = s
        s = float(xnrm) if herm else float(np.sqrt(xnrm))
    return s


def snorm_dense(A: np.ndarray, opts: Optional[LRAOptions] = None, seed: int = 0, x0=None) -> float:
    herm = A.shape[0] == A.shape[1] and np.array_equal(A, A.T)
    return snorm(lambda x: A @ x, lambda y: A.T @ y, A.shape[1], opts, seed, x0, herm)


def snormdiff_lowrank(A: np.ndarray, left: np.ndarray, right: np.ndarray,
                      opts: Optional[LRAOptions] = None, seed: int = 0, x0=None) -> float:
    """snormdiff(A, F) for F = left @ right without forming F (src/snorm.jl:49-53)."""
    return snorm(lambda x: A @ x - left @ (right @ x),
                 lambda y: A.T @ y - right.T @ (left.T @ y), A.shape[1], opts, seed, x0)


def id_error(A: np.ndarray, V: IDPackedV, trans: str = "n", seed: int = 0) -> float:
    """snormdiff(A, ID(A, V)) / snorm(A) -- the BASELINE error metric."""
    Aop = A if trans == "n" else A.T
    C = Aop[:, V.sk - 1]
    return snormdiff_lowrank(Aop, C, V.matrix(), seed=seed) / snorm_dense(Aop, seed=seed)



def prange(A: np.ndarray, opts: LRAOptions, rand: Optional[RandomInputs] = None, trans: str = "n",
           rand2: Optional[RandomInputs] = None) -> np.ndarray:
    """prange(trans, A, opts) (src/prange.jl:14-62) -> Q.  trans "b": rand drives the sketch of A' (factored first),
    rand2 the sketch of A."""
    chkopts(opts)
    if trans not in ("n", "c", "b"):
        raise ValueError("trans")
    if trans == "b":
        if A.shape[0] != A.shape[1]:
            raise ValueError("DimensionMismatch: matrix is not square")
        if np.array_equal(A, A.T):
            return prange(A, opts, rand, "n")
        o = opts.copy(pqrfact_retval="qr")
        if o.sketch == "none":
            Fr, Fc = pqrfact_none(A, o, "c"), pqrfact_none(A, o, "n")
        else:
            Fr = sketchfact(A, o, rand, "c", side="right")
            Fc = sketchfact(A, o, rand2, "n", side="right")
        B = np.asfortranarray(np.hstack([Fr.Q @ Fr.R[:, :Fr.k], Fc.Q @ Fc.R[:, :Fc.k]]))
        return pqrfact_none(B, opts.copy(pqrfact_retval="q"), "n").Q
    o = opts.copy(pqrfact_retval="q")
    if o.sketch == "none":
        return pqrfact_none(A, o, trans).Q
    if o.sketch == "sub":                                  # prange_sub (src/prange.jl:64-77)
        F = sketchfact(A, o.copy(pqrfact_retval="t"), rand, trans)
        C = getcols(A, F.p[:F.k], trans)
        return qr_thin(C)[0]
    return sketchfact(A, o, rand, trans, side="right").Q


def curfact(A: np.ndarray, opts: LRAOptions, rand1: Optional[RandomInputs] = None,
            rand2: Optional[RandomInputs] = None):
    """curfact (src/cur.jl:532-566): (rows, cols), 1-based; Hermitian A -> (cols, cols)."""
    chkopts(opts)
    o1 = opts.copy(pqrfact_retval="t")
    m, n = A.shape
    if m == n and np.array_equal(A, A.T):
        V = idfact(A, o1, rand1, "n")
        return V.sk.copy(), V.sk.copy()
    if m >= n:
        rows = idfact(A, o1, rand1, "c").sk
        cols = idfact(np.asfortranarray(A[rows - 1, :]), o1, rand2, "n").sk
    else:
        cols = idfact(A, o1, rand1, "n").sk
        rows = idfact(np.asfortranarray(A[:, cols - 1]), o1, rand2, "c").sk
    k = min(len(rows), len(cols))
    return rows[:k].copy(), cols[:k].copy()


def pheigrank(w: np.ndarray, opts: LRAOptions) -> Tuple[int, int]:
    """src/pheig.jl:322-341 on ascending w: (kn, kp) kept at the negative / positive end."""
    n = len(w)
    wmax = max(abs(w[0]), abs(w[-1]))
    lo = int(np.searchsorted(w, 0.0, side="left"))
    hi = int(np.searchsorted(w, 0.0, side="right"))
    ptol = max(opts.atol, opts.rtol * wmax)

    def rank1(v):
        k = len(v)
        k = min(opts.rank, k) if opts.rank >= 0 else k
        for i in range(1, k):
            if abs(v[i]) <= ptol:
                return i
        return k
    return rank1(w[:lo]), rank1(w[::-1][:n - hi])


def pheigfact(A: np.ndarray, opts: LRAOptions, rand: Optional[RandomInputs] = None):
    """pheigfact (src/pheig.jl:276-296): (values, vectors); the eigen cluster re-orthonormalisation (pheigorth!,
    :344-364) is a no-op for numpy's eigh, whose vectors are orthonormal to machine precision."""
    if A.shape[0] != A.shape[1] or not np.array_equal(A, A.T):
        raise ValueError("matrix must be Hermitian")
    V = idfact(A, opts, rand, "n")
    k, n = V.k, A.shape[0]
    if k == 0:
        return np.zeros(0), np.zeros((n, 0)), V
    Z = np.zeros((n, k))
    Z[V.p - 1, :] = np.vstack([np.eye(k), V.T.T])              # Matrix(:c, V) = P [I; T']
    Q, R = qr_thin(Z)
    Bm = R @ (A[np.ix_(V.sk - 1, V.sk - 1)] @ R.T)
    w, X = np.linalg.eigh((Bm + Bm.T) / 2)
    kn, kp = pheigrank(w, opts)
    if kn + kp < k:
        idx = np.r_[0:kn, k - kp:k]
        w, X = w[idx], X[:, idx]
    X = np.array(X, order="F")
    pheigorth(w, X, opts)
    return w, Q @ X, V


def pheigorth(values: np.ndarray, vectors: np.ndarray, opts: LRAOptions) -> None:
    """pheigorth! (src/pheig.jl:342-364): runs of adjacent eigenvalues with symrelerr(va, vb) = 2|va - vb| / |va + vb|
    (src/util.jl:118) <= pheig_orthtol are re-orthogonalised in place, v_j -= <v_i, v_j> v_i for i < j in the run (no
    normalisation)."""
    n = len(values)
    a = 0
    while a < n:
        va = values[a]
        b = a + 1
        while b < n:
            vb = values[b]
            with np.errstate(divide="ignore", invalid="ignore"):
                err = 2 * abs((va - vb) / (va + vb))
            if err > opts.pheig_orthtol:
                break
            b += 1
        b -= 1
        for i in range(a, b + 1):
            for j in range(i + 1, b + 1):
                vectors[:, j] -= (vectors[:, i] @ vectors[:, j]) * vectors[:, i]
        a = b + 1


def cur_core(A: np.ndarray, rows: np.ndarray, cols: np.ndarray):
    """CUR(A, U::CURPackedU) (src/cur.jl:85-93): C = A[:, cols], R = A[rows, :], U, s, V = svd!(C[rows, :]) and
    U2 = PartialSVD(V, 1 ./ s, U') -> (C, (V, 1/s, U'), R).  rows / cols 1-based."""
    C = np.asfortranarray(A[:, cols - 1])
    R = np.asfortranarray(A[rows - 1, :])
    import scipy.linalg as sla
    U, s, Vt = sla.svd(np.asfortranarray(C[rows - 1, :]), full_matrices=False, lapack_driver="gesdd")
    with np.errstate(divide="ignore"):
        return C, (np.asfortranarray(Vt.T), 1.0 / s, np.asfortranarray(U.T)), R


def hermcur_core(A: np.ndarray, cols: np.ndarray):
    """HermCUR(A, U) (src/cur.jl:99-105): C = A[:, cols], F = eigen!(Hermitian(C[cols, :])),
    U = PartialHermEigen(1 ./ F.values, F.vectors) -> (C, (1/values, vectors))."""
    C = np.asfortranarray(A[:, cols - 1])
    W = C[cols - 1, :]
    w, X = np.linalg.eigh(np.triu(W) + np.triu(W, 1).T)      # Hermitian(.) reads the upper triangle
    with np.errstate(divide="ignore"):
        return C, (1.0 / w, X)
