"""ctypes binding of include/brapprox.h (no torch types cross the boundary)."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import numpy as np

BRA_MAX_ROUNDS = 24
SKETCH_CODES = {"none": 0, "randn": 1, "sprn": 2, "srft": 3, "sub": 4}
RET_Q, RET_R, RET_T = 1, 2, 4
F_P, F_T, F_Q, F_R, F_U, F_S, F_VT, F_TAU, F_BSKETCH = 1, 2, 3, 4, 5, 6, 7, 8, 9

lib_path = os.environ.get("BRA_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libbrapprox.so")


class BraError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libbrapprox status {code}: {msg}")
        self.code = code


class bra_opts(C.Structure):
    _fields_ = [
        ("atol", C.c_double), ("rtol", C.c_double), ("rank", C.c_int64), ("nb", C.c_int64),
        ("sketch", C.c_int32), ("sketch_randn_niter", C.c_int32), ("sketchfact_adap", C.c_int32),
        ("retval_mask", C.c_int32), ("maxdet_tol", C.c_double), ("maxdet_niter", C.c_int64),
        ("samp_a", C.c_int64), ("samp_b", C.c_int64), ("seed", C.c_uint64), ("verb", C.c_int32),
        ("flags", C.c_int32), ("pheig_orthtol", C.c_double),
    ]


_pp_d = C.POINTER(C.c_void_p)


class bra_rand(C.Structure):
    _fields_ = [
        ("n_rounds", C.c_int32), ("reserved", C.c_int32),
        ("omega", _pp_d), ("d", _pp_d), ("idx", _pp_d), ("perm", _pp_d), ("s", _pp_d), ("r", _pp_d),
    ]


class bra_info(C.Structure):
    _fields_ = [
        ("m", C.c_int64), ("n", C.c_int64), ("k", C.c_int64), ("ksvd", C.c_int64),
        ("rounds", C.c_int32), ("reserved", C.c_int32),
        ("orders", C.c_int64 * BRA_MAX_ROUNDS), ("ks", C.c_int64 * BRA_MAX_ROUNDS),
        ("steps", C.c_int64 * BRA_MAX_ROUNDS),
    ]


def _load() -> C.CDLL:
    if not os.path.exists(lib_path):
        raise ImportError(
            f"{lib_path} is missing: build it with `python lowrankapprox.jl_b200/build.py` "
            "(there is no CPU fallback)")
    return C.CDLL(lib_path)


lib = _load()
_vp, _i64, _d = C.c_void_p, C.c_int64, C.c_double
lib.bra_version.restype = C.c_int
lib.bra_create.argtypes = [C.POINTER(_vp), C.c_int]
lib.bra_destroy.argtypes = [_vp]
lib.bra_last_error.argtypes = [_vp]
lib.bra_last_error.restype = C.c_char_p
lib.bra_opts_default.argtypes = [C.POINTER(bra_opts)]
lib.bra_opts_default.restype = None
lib.bra_chkopts.argtypes = [_vp, C.POINTER(bra_opts)]
lib.bra_launch_count.argtypes = [_vp]
lib.bra_launch_count.restype = C.c_uint64
lib.bra_sync.argtypes = [_vp]
lib.bra_debug_maxdet_swaps.argtypes = [_vp]
lib.bra_debug_maxdet_swaps.restype = C.c_int64
lib.bra_stream.argtypes = [_vp]
lib.bra_stream.restype = _vp
lib.bra_sketch_randn_f64.argtypes = [_vp, C.c_char, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _vp, _i64]
lib.bra_sketch_sub_f64.argtypes = [_vp, C.c_char, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _i64]
lib.bra_sketch_sprn_f64.argtypes = [_vp, C.c_char, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _i64]
lib.bra_sketch_srft_f64.argtypes = [_vp, C.c_char, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _i64]
lib.bra_idfact_batched_f64.argtypes = [_vp, _i64, _i64, _i64, _vp, _i64, _i64, C.POINTER(bra_opts), _vp, _i64, _vp, _i64,
                                       _vp, _vp, _vp, _i64, _i64]
lib.bra_batched_unfinished.argtypes = [_vp]
lib.bra_batched_unfinished.restype = C.c_int64
lib.bra_comm_unique_id.argtypes = [_vp]
lib.bra_comm_init.argtypes = [_vp, _vp, C.c_int, C.c_int]
lib.bra_comm_destroy.argtypes = [_vp]
lib.bra_set_row_shard.argtypes = [_vp, _i64, _i64]
lib.bra_collective_count.argtypes = [_vp]
lib.bra_collective_count.restype = C.c_uint64
lib.bra_geqp3_adap_f64.argtypes = [_vp, _i64, _i64, _vp, _i64, C.POINTER(bra_opts), _vp, _vp,
                                   C.POINTER(_i64), C.POINTER(_i64), _vp, _i64, C.POINTER(_i64)]
lib.bra_trsolve_T_f64.argtypes = [_vp, _i64, _i64, _vp, _i64, _vp, _i64]
lib.bra_idfact_f64.argtypes = [_vp, C.c_char, _i64, _i64, _vp, _i64, C.POINTER(bra_opts), C.POINTER(bra_rand)]
lib.bra_pqrfact_f64.argtypes = [_vp, C.c_char, _i64, _i64, _vp, _i64, C.POINTER(bra_opts), C.POINTER(bra_rand)]
lib.bra_pheigfact_f64.argtypes = [_vp, _i64, _vp, _i64, C.POINTER(bra_opts), C.POINTER(bra_rand)]
lib.bra_sketchfact_f64.argtypes = [_vp, C.c_char, C.c_char, _i64, _i64, _vp, _i64, C.POINTER(bra_opts),
                                   C.POINTER(bra_rand)]
lib.bra_cur_f64.argtypes = [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, C.c_int]
lib.bra_widen_f32.argtypes = [_vp, _i64, _i64, _vp, _i64, _vp, _vp]
lib.bra_widen_f32.restype = C.c_int
lib.bra_debug_sketch_rows.argtypes = [_vp]
lib.bra_debug_sketch_rows.restype = C.c_int64
lib.bra_debug_randn.argtypes = [_vp, _vp, _i64, C.c_uint64, C.c_uint64]
lib.bra_debug_meta.argtypes = [_vp, C.c_int, _vp, _i64, _i64, C.c_uint64, C.c_uint64]
lib.bra_prange_f64.argtypes = [_vp, C.c_char, _i64, _i64, _vp, _i64, C.POINTER(bra_opts), C.POINTER(bra_rand),
                               C.POINTER(bra_rand)]
lib.bra_snorm_f64.argtypes = [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _vp, _i64, C.POINTER(bra_opts), _i64, _vp,
                              C.POINTER(C.c_double), C.POINTER(_i64)]
lib.bra_psvdfact_f64.argtypes = [_vp, _i64, _i64, _vp, _i64, C.POINTER(bra_opts), C.POINTER(bra_rand)]
lib.bra_psvdvals_f64.argtypes = [_vp, _i64, _i64, _vp, _i64, C.POINTER(bra_opts), C.POINTER(bra_rand)]
lib.bra_psvd_set_outputs.argtypes = [_vp, _vp, _i64, _i64, _vp, _i64, _vp, _i64]
lib.bra_psvd_outputs_done.argtypes = [_vp]
lib.bra_get_info.argtypes = [_vp, C.POINTER(bra_info)]
lib.bra_fetch.argtypes = [_vp, C.c_int, _vp, _i64]
lib.bra_profile_enable.argtypes = [_vp, C.c_int]
lib.bra_profile_read.argtypes = [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
PROF_TAGS = ["omega", "gemm", "splitk", "qrcp", "gather", "trsolve", "tail", "sketch_other", "svd", "qr", "tail_gemm", "comm", "batched"]
lib.bra_debug_qrcp_phases.argtypes = [_vp, C.POINTER(C.c_int32)]
lib.bra_probe_fp64_peak.argtypes = [_vp, C.POINTER(C.c_double)]
lib.bra_probe_exchange_latency.argtypes = [_vp, C.c_int, C.c_int, C.POINTER(C.c_double)]

EPS = float(np.finfo(np.float64).eps)


@dataclass
class LRAOptions:
    """Mirror of LRAOptions(Float64) (src/LowRankApprox.jl:77-119), same defaults."""
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
    seed: int = 0            # fast-mode device RNG key (no reference counterpart: Julia's global RNG)
    sketch_fresh: bool = False   # fast mode: independent Gaussian Omega per adaptive round (BRA_OPT_FRESH_SKETCH) instead of nested rounds

    @classmethod
    def for_eltype(cls, dtype, **kw) -> "LRAOptions":
        """LRAOptions(T; args...) (src/LowRankApprox.jl:96-118): pheig_orthtol = sqrt(eps(T)), rtol = 5 eps(T)."""
        e = float(np.finfo(np.dtype(dtype)).eps)
        return cls(pheig_orthtol=float(np.sqrt(e)), rtol=5 * e).copy(**kw)

    def copy(self, **kw) -> "LRAOptions":
        """copy(opts; args...) (src/LowRankApprox.jl:122-131): never mutates the caller's object."""
        import dataclasses
        for k in kw:
            if k not in self.__dataclass_fields__:
                raise TypeError(f"LRAOptions has no field {k}")       # Julia: setfield! error
        return dataclasses.replace(self, **kw)

    def chk(self) -> None:
        """chkopts! (src/LowRankApprox.jl:133-141); ArgumentError -> ValueError."""
        if not self.atol >= 0:
            raise ValueError("atol")
        if not self.nb > 0:
            raise ValueError("nb")
        if not self.pheig_orthtol >= 0:
            raise ValueError("pheig_orthtol")
        if not self.rtol >= 0:
            raise ValueError("rtol")
        if self.sketch not in SKETCH_CODES:
            raise ValueError("sketch")
        self.pqrfact_retval = self.pqrfact_retval.lower()

    def _samp_affine(self) -> Tuple[int, int]:
        """The three *_samp closures cannot cross the C ABI: evaluate the active one into (a, b)
        with order = a*n + b, rejecting non-affine closures."""
        f = {"randn": self.sketchfact_randn_samp, "srft": self.sketchfact_srft_samp,
             "sub": self.sketchfact_sub_samp}.get(self.sketch)
        if f is None:
            return 0, 0
        b = int(f(0))
        a = int(f(1)) - b
        for n in (2, 32, 64, 1000):
            if int(f(n)) != a * n + b:
                raise ValueError("sketchfact_*_samp must be affine to cross the C ABI")
        return a, b

    def to_c(self) -> bra_opts:
        o = bra_opts()
        lib.bra_opts_default(C.byref(o))
        o.atol, o.rtol, o.rank, o.nb = self.atol, self.rtol, self.rank, self.nb
        o.sketch = SKETCH_CODES[self.sketch]
        o.sketch_randn_niter = self.sketch_randn_niter
        o.sketchfact_adap = int(bool(self.sketchfact_adap))
        rv = self.pqrfact_retval.lower()
        o.retval_mask = (RET_Q if "q" in rv else 0) | (RET_R if "r" in rv else 0) | (RET_T if "t" in rv else 0)
        o.maxdet_tol, o.maxdet_niter = self.maxdet_tol, self.maxdet_niter
        o.samp_a, o.samp_b = self._samp_affine()
        o.seed = self.seed
        o.verb = int(bool(self.verb))
        o.pheig_orthtol = self.pheig_orthtol
        o.flags = 1 if self.sketch_fresh else 0
        return o


class DeviceMatrix:
    """A column-major FP64 matrix already resident on the GPU (pointer + dims + ld).
    `keep` holds whatever owns the memory (e.g. a torch tensor)."""

    def __init__(self, ptr: int, m: int, n: int, ld: Optional[int] = None, keep=None):
        self.ptr, self.m, self.n, self.ld, self.keep = int(ptr), int(m), int(n), int(ld or max(m, 1)), keep

    @property
    def shape(self):
        return (self.m, self.n)

    @staticmethod
    def from_torch(t) -> "DeviceMatrix":
        """t: 2-D CUDA float64 tensor whose memory is column-major, i.e. stride(0) == 1
        (for instance `x.t()` of a contiguous n x m tensor)."""
        assert t.is_cuda and t.dim() == 2 and str(t.dtype) == "torch.float64"
        m, n = t.shape
        assert t.stride(0) == 1 or m == 1, "column-major storage required"
        ld = t.stride(1) if n > 1 else max(m, 1)
        return DeviceMatrix(t.data_ptr(), m, n, ld, keep=t)


def mat_arg(A):
    """-> (pointer, m, n, ld, keepalive) for a numpy array (host) or a DeviceMatrix."""
    if isinstance(A, DeviceMatrix):
        return C.c_void_p(A.ptr), A.m, A.n, A.ld, A
    if hasattr(A, "is_cuda"):
        return mat_arg(DeviceMatrix.from_torch(A))
    a = np.asarray(A)
    if a.dtype != np.float64:
        raise TypeError("Float64 (and, through Context.widen_f32, Float32) matrices only; got " + str(a.dtype))
    if a.ndim != 2:
        raise ValueError("matrix expected")
    if not a.flags.f_contiguous:
        a = np.asfortranarray(a)
    return C.c_void_p(a.ctypes.data), a.shape[0], a.shape[1], max(a.shape[0], 1), a


class Context:
    """One device + one stream + all workspaces (bra_create / bra_destroy)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        rc = lib.bra_create(C.byref(self._h), device)
        if rc != 0:
            msg = lib.bra_last_error(self._h).decode() if self._h else "bra_create failed"
            if self._h:
                lib.bra_destroy(self._h)
                self._h = C.c_void_p()
            raise BraError(rc, msg)
        self.device = device

    def close(self):
        if self._h:
            lib.bra_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != 0:
            raise BraError(rc, lib.bra_last_error(self._h).decode())

    @property
    def handle(self):
        return self._h

    def launch_count(self) -> int:
        return int(lib.bra_launch_count(self._h))

    def maxdet_swaps(self) -> int:
        """Column swaps made by the last maxdet post-processing (maxdet_swapcols!, src/pqr.jl:444-478)."""
        return int(lib.bra_debug_maxdet_swaps(self._h))

    def collective_count(self) -> int:
        return int(lib.bra_collective_count(self._h))

    def set_row_shard(self, row0: int, m_global: int):
        """This ctx holds rows [row0, row0 + m_local) of an m_global-row matrix (0, 0 = not sharded)."""
        self.check(lib.bra_set_row_shard(self._h, row0, m_global))

    def profile_enable(self, on: bool = True):
        self.check(lib.bra_profile_enable(self._h, int(on)))

    def profile_read(self) -> dict:
        """{tag: (milliseconds, spans)} accumulated since profile_enable (CUDA events on the ctx stream)."""
        ms = (C.c_double * len(PROF_TAGS))()
        calls = (C.c_int64 * len(PROF_TAGS))()
        self.check(lib.bra_profile_read(self._h, ms, calls))
        return {t: (ms[i], int(calls[i])) for i, t in enumerate(PROF_TAGS)}

    def qrcp_phases(self):
        out = (C.c_int32 * 6)()
        lib.bra_debug_qrcp_phases(self._h, out)
        return dict(zip(["merge_hdr", "dlarfg_rec", "gather", "fetch", "update", "unused"],
                        [int(x) for x in out]))

    def sync(self):
        self.check(lib.bra_sync(self._h))

    def info(self) -> bra_info:
        inf = bra_info()
        self.check(lib.bra_get_info(self._h, C.byref(inf)))
        return inf

    def widen_f32(self, A: np.ndarray) -> "DeviceMatrix":
        """bra_widen_f32: uploads a Float32 matrix and widens it on the device; the returned DeviceMatrix (context-owned
        memory, valid until the next widen_f32 on this context) is accepted wherever an FP64 matrix is."""
        a = np.asarray(A)
        if a.dtype != np.float32 or a.ndim != 2:
            raise TypeError("widen_f32: 2-D float32 array expected")
        if not a.flags.f_contiguous:
            a = np.asfortranarray(a)
        dA, ld = C.c_void_p(), C.c_int64()
        self.check(lib.bra_widen_f32(self._h, a.shape[0], a.shape[1], C.c_void_p(a.ctypes.data), max(a.shape[0], 1),
                                     C.byref(dA), C.byref(ld)))
        return DeviceMatrix(dA.value or 0, a.shape[0], a.shape[1], ld.value, keep=self)

    def fetch(self, which: int, shape, dtype=np.float64, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Copies one factor of the last call into host memory.  `out` (optional) is a caller-owned column-major
        buffer of at least `shape` (e.g. pinned memory: the copy then runs at the full PCIe rate and nothing is
        allocated); the returned array is the leading `shape` block of it."""
        if out is None:
            out = np.empty(shape, dtype=dtype, order="F")
            ld = max(shape[0], 1) if len(shape) == 2 else 1
        else:
            if out.dtype != dtype or out.ndim != len(shape) or any(o < s_ for o, s_ in zip(out.shape, shape)):
                raise ValueError("out: wrong dtype or too small")
            if len(shape) == 2:
                if out.strides[0] != out.itemsize:
                    raise ValueError("out: must be column-major (unit stride along rows)")
                ld = max(out.strides[1] // out.itemsize, 1)
                out = out[:shape[0], :shape[1]]
            else:
                if out.strides[0] != out.itemsize:
                    raise ValueError("out: must be contiguous")
                ld = 1
                out = out[:shape[0]]
        if out.size == 0:
            return out
        self.check(lib.bra_fetch(self._h, which, C.c_void_p(out.ctypes.data), ld))
        return out


@dataclass
class PartialHermEigen:
    """PartialHermEigen (src/pheig.jl:4-8): values ascending (negative part, then positive part), vectors n x k."""
    values: np.ndarray
    vectors: np.ndarray
    k_id: int = 0
    rounds: List[Tuple[int, int]] = field(default_factory=list)

    def __getitem__(self, key):
        if key == "values":
            return self.values
        if key == "vectors":
            return self.vectors
        if key == "k":
            return len(self.values)
        raise KeyError(key)

    def matrix(self) -> np.ndarray:
        return (self.vectors * self.values) @ self.vectors.T


@dataclass
class CURPackedU:
    """CURPackedU / HermCURPackedU (src/cur.jl:14-26): 1-based row and column index sets."""
    rows: np.ndarray
    cols: np.ndarray
    hermitian: bool = False

    def __getitem__(self, key):
        if key == "rows":
            return self.rows
        if key == "cols":
            return self.cols
        if key == "k":
            return len(self.cols)
        raise KeyError(key)


@dataclass
class CUR:
    """CUR / HermCUR (src/cur.jl:60-72): A ~ C U R with C = A[:, cols], R = A[rows, :] and U the pseudo-inverse of the
    k x k core A[rows, cols], kept factored: PartialSVD(V, 1 ./ s, U') (general) or PartialHermEigen(1 ./ values,
    vectors) (Hermitian; then R = C')."""
    rows: np.ndarray
    cols: np.ndarray
    C: np.ndarray
    U: object
    R: Optional[np.ndarray]

    def __getitem__(self, key):
        if key in ("rows", "cols", "C", "U", "R"):
            return getattr(self, key) if not (key == "R" and self.R is None) else self.C.T
        if key == "k":
            return len(self.cols)
        raise KeyError(key)

    def matrix(self) -> np.ndarray:
        R = self.R if self.R is not None else self.C.T
        return self.C @ (self.U.matrix() @ R)


@dataclass
class IDPackedV:
    """IDPackedV (src/id.jl:15-19): sk, rd are 1-based like the reference."""
    sk: np.ndarray
    rd: np.ndarray
    T: np.ndarray
    rounds: List[Tuple[int, int]] = field(default_factory=list)
    steps: List[int] = field(default_factory=list)

    def __getitem__(self, key):
        # getindex (src/id.jl:72-81); unknown key -> KeyError like the reference
        if key == "sk":
            return self.sk
        if key == "rd":
            return self.rd
        if key == "T":
            return self.T
        if key == "p":
            return self.p
        if key == "k":
            return self.k
        raise KeyError(key)

    @property
    def k(self) -> int:
        return len(self.sk)

    @property
    def p(self) -> np.ndarray:
        return np.concatenate([self.sk, self.rd])

    @property
    def shape(self):
        return (self.k, len(self.sk) + len(self.rd))

    def matrix(self) -> np.ndarray:
        """Matrix(V) = [I T] P' (src/id.jl:32-46)."""
        k, n = self.shape
        M = np.zeros((k, n))
        M[:, self.p - 1] = np.hstack([np.eye(k), self.T])
        return M


@dataclass
class PartialQR:
    """PartialQR (src/pqr.jl:36-40)."""
    Q: np.ndarray
    R: np.ndarray
    p: np.ndarray
    rounds: List[Tuple[int, int]] = field(default_factory=list)
    T: Optional[np.ndarray] = None

    @property
    def k(self) -> int:
        return self.Q.shape[1]

    def __getitem__(self, key):
        if key in ("Q", "R", "p", "k"):
            return getattr(self, key)
        raise KeyError(key)

    def matrix(self) -> np.ndarray:
        M = np.zeros((self.Q.shape[0], self.R.shape[1]))
        M[:, self.p - 1] = self.Q @ self.R
        return M


@dataclass
class PQRFactors:
    """PartialQRFactors (src/pqr.jl:44-50): what pqrback_postproc returns unless retval is exactly "qr"; the result of
    sketchfact.  Q / R / T are None when pqrfact_retval does not ask for them."""
    Q: Optional[np.ndarray]
    R: Optional[np.ndarray]
    p: np.ndarray
    k: int
    T: Optional[np.ndarray]
    rounds: List[Tuple[int, int]] = field(default_factory=list)

    def __getitem__(self, key):
        if key in ("Q", "R", "p", "k", "T"):
            return getattr(self, key)
        raise KeyError(key)


@dataclass
class PartialSVD:
    """PartialSVD (src/psvd.jl:4-16)."""
    U: np.ndarray
    S: np.ndarray
    Vt: np.ndarray
    k_id: int = 0
    rounds: List[Tuple[int, int]] = field(default_factory=list)

    def __getitem__(self, key):
        if key in ("U", "S", "Vt"):
            return getattr(self, key)
        if key == "V":
            return self.Vt.T
        if key == "k":
            return len(self.S)
        raise KeyError(key)

    def matrix(self) -> np.ndarray:
        return (self.U * self.S) @ self.Vt
