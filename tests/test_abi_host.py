"""CPU tests of the boundary and the host logic (no compute calls: there is no GPU here).

* the C-ABI library loads and exports EVERY function include/brapprox.h declares;
* the POD structs of the ctypes binding match the header's layout (sizes);
* LRAOptions mirrors the reference's defaults, copy semantics and validation
  (src/LowRankApprox.jl:96-148);
* without a usable B200 the product path fails LOUDLY (no CPU fallback, nothing routes through oracle/).
"""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "brapprox.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bra_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import brapprox
    names = _declared_functions()
    assert len(names) >= 15
    for nm in names:
        assert hasattr(brapprox.lib, nm), f"libbrapprox.so does not export {nm}"
    assert brapprox.lib.bra_version() >= 100


def test_struct_layouts_match_header():
    """Compile a tiny C program against the header and compare sizeof with the ctypes mirrors."""
    from brapprox import _binding as B
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "sz.c")
        open(c, "w").write('#include <stdio.h>\n#include "brapprox.h"\nint main(){printf("%zu %zu %zu\\n",'
                           "sizeof(bra_opts),sizeof(bra_rand),sizeof(bra_info));return 0;}\n")
        exe = os.path.join(d, "sz")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        so, sr, si = (int(x) for x in subprocess.check_output([exe]).split())
    assert ctypes.sizeof(B.bra_opts) == so
    assert ctypes.sizeof(B.bra_rand) == sr
    assert ctypes.sizeof(B.bra_info) == si


def test_opts_defaults_match_reference():
    import brapprox
    from brapprox import _binding as B
    o = brapprox.LRAOptions()
    eps = np.finfo(np.float64).eps
    assert (o.atol, o.rtol, o.rank, o.nb, o.sketch) == (0.0, 5 * eps, -1, 32, "randn")
    assert (o.sketch_randn_niter, o.sketchfact_adap, o.maxdet_tol, o.maxdet_niter) == (0, True, -1.0, -1)
    assert (o.pqrfact_retval, o.snorm_niter, o.verb) == ("qr", 32, True)
    assert [o.sketchfact_randn_samp(32), o.sketchfact_srft_samp(32), o.sketchfact_sub_samp(32)] == [40, 40, 136]
    c = B.bra_opts()
    B.lib.bra_opts_default(ctypes.byref(c))
    assert (c.atol, c.rtol, c.rank, c.nb, c.sketch, c.sketchfact_adap) == (0.0, 5 * eps, -1, 32, 1, 1)
    assert c.maxdet_tol < 0 and c.retval_mask == 3


def test_opts_copy_and_validation():
    import brapprox
    o = brapprox.LRAOptions(rtol=1e-8)
    o2 = o.copy(rank=5, sketch="srft")
    assert (o.rank, o.sketch) == (-1, "randn") and (o2.rank, o2.sketch, o2.rtol) == (5, "srft", 1e-8)
    with pytest.raises(TypeError):
        o.copy(not_a_field=1)
    for bad in (dict(atol=-1.0), dict(nb=0), dict(rtol=-1e-3), dict(sketch="fft")):
        with pytest.raises(ValueError):
            brapprox.LRAOptions(**bad).chk()
    o3 = brapprox.LRAOptions(pqrfact_retval="QRT")
    o3.chk()
    assert o3.pqrfact_retval == "qrt"
    assert o3.to_c().retval_mask == 7


def test_samp_closures_cross_as_affine():
    import brapprox
    o = brapprox.LRAOptions(sketchfact_randn_samp=lambda n: 2 * n + 4)
    c = o.to_c()
    assert (c.samp_a, c.samp_b) == (2, 4)
    with pytest.raises(ValueError):
        brapprox.LRAOptions(sketchfact_randn_samp=lambda n: n * n).to_c()
    o = brapprox.LRAOptions(sketch="sub")
    c = o.to_c()
    assert (c.samp_a, c.samp_b) == (4, 8)


def test_result_type_accessors():
    import brapprox
    V = brapprox.IDPackedV(np.array([2, 1]), np.array([3]), np.array([[0.5], [0.25]]))
    assert V["k"] == 2 and V.shape == (2, 3)
    np.testing.assert_array_equal(V["p"], [2, 1, 3])
    M = V.matrix()
    np.testing.assert_allclose(M[:, [1, 0, 2]], np.hstack([np.eye(2), V.T]))
    with pytest.raises(KeyError):
        V["nope"]


def test_no_cpu_fallback():
    """On a box without a B200 the context cannot be created and says why."""
    import brapprox
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present; the loud-failure path is exercised on CPU boxes")
    except ImportError:
        pass
    with pytest.raises(brapprox.BraError) as e:
        brapprox.Context(0)
    assert e.value.code != 0
    with pytest.raises(brapprox.BraError):
        brapprox.idfact(np.eye(8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lowrankapprox.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".jl")):
                txt = open(os.path.join(dp, f)).read()
                bad = re.findall(r"^\s*(?:import|from)\s+(?:lra_oracle|oracle)\b|#include\s+[\"<][^\">]*oracle",
                                 txt, flags=re.M)
                assert not bad, (f, bad)


def test_frontend_argument_checks_without_gpu():
    """The reference's argument checks fire before any device work (prange_chktrans src/prange.jl:79-80,
    sketchfact_chkargs src/sketch.jl:80-84, chktrans src/LowRankApprox.jl:150)."""
    import brapprox
    A = np.zeros((4, 4))
    with pytest.raises(ValueError):
        brapprox.prange(A, trans="x")
    with pytest.raises(ValueError):
        brapprox.sketch(A, 2, side="middle")
    with pytest.raises(ValueError):
        brapprox.sketch(A, -1)
    with pytest.raises(ValueError):
        brapprox.idfact(A, trans="t")


def test_random_input_shapes_are_checked_on_the_host():
    """Caller-supplied per-round random inputs (bra_rand) are validated against the round's sketch order and the
    contracted dimension before they cross the C ABI (the C side reads them at their nominal sizes)."""
    import brapprox
    from brapprox._frontend import _RandPack
    o = brapprox.LRAOptions(rtol=1e-9)                      # randn, adaptive: orders 40, 72, ...
    good = [{"Omega": np.zeros((40, 100))}, {"Omega": np.zeros((72, 100))}]
    _RandPack(good, o, 100)
    for bad in ([{"Omega": np.zeros((40, 99))}], [{"Omega": np.zeros((41, 100))}],
                [{"Omega": np.zeros((40, 100))}, {"Omega": np.zeros((40, 100))}]):
        with pytest.raises(ValueError, match="DimensionMismatch"):
            _RandPack(bad, o, 100)
    o = brapprox.LRAOptions(sketch="srft", rank=10, sketchfact_adap=False)      # one round of order 18
    _RandPack([{"d": np.ones(64), "idx": np.ones(18, dtype=np.int64)}], o, 64)
    with pytest.raises(ValueError, match="DimensionMismatch"):
        _RandPack([{"d": np.ones(63), "idx": np.ones(18, dtype=np.int64)}], o, 64)
    with pytest.raises(ValueError, match="DimensionMismatch"):
        _RandPack([{"d": np.ones(64), "idx": np.ones(10, dtype=np.int64)}], o, 64)
    o = brapprox.LRAOptions(sketch="sprn")                  # order 32 in round 0, no oversampling
    _RandPack([{"perm": np.arange(50), "s": np.ones(50)}], o, 50)
    with pytest.raises(ValueError, match="DimensionMismatch"):
        _RandPack([{"perm": np.arange(49), "s": np.ones(50)}], o, 50)
    o = brapprox.LRAOptions(sketch="sub")                   # order 4*32 + 8
    _RandPack([{"r": np.ones(136, dtype=np.int64)}], o, 50)
    with pytest.raises(ValueError, match="DimensionMismatch"):
        _RandPack([{"r": np.ones(40, dtype=np.int64)}], o, 50)


def test_sketchfact_argument_checks_without_gpu():
    """sketchfact_chkargs (src/sketch.jl:80-84) and chktrans fire before any device work."""
    import brapprox
    A = np.zeros((4, 4))
    with pytest.raises(ValueError):
        brapprox.sketchfact(A, side="up")
    with pytest.raises(ValueError):
        brapprox.sketchfact(A, trans="x")
    with pytest.raises(ValueError):
        brapprox.sketchfact(A, sketch="none")


def test_float32_default_options_and_rounding():
    """LRAOptions(Float32) (src/LowRankApprox.jl:96-118): rtol = 5 eps(Float32), pheig_orthtol = sqrt(eps(Float32));
    the mirror rounds FP64 factors to Float32 and leaves index sets and bookkeeping alone."""
    import brapprox
    from brapprox import _frontend as fe
    e = float(np.finfo(np.float32).eps)
    o32 = brapprox.LRAOptions.for_eltype(np.float32, sketch="srft")
    assert o32.rtol == 5 * e and o32.pheig_orthtol == float(np.sqrt(e)) and o32.sketch == "srft"
    o64 = brapprox.LRAOptions.for_eltype(np.float64)
    assert o64.rtol == brapprox.LRAOptions().rtol
    V = brapprox.IDPackedV(np.array([2, 1]), np.array([3]), np.ones((2, 1)), [(40, 2)], [2])
    W = fe._narrow(V)
    assert W.T.dtype == np.float32 and W.sk.dtype == V.sk.dtype and W.rounds == [(40, 2)]
    F = fe._narrow(brapprox.PQRFactors(None, np.eye(2), np.array([1, 2]), 2, None))
    assert F.Q is None and F.R.dtype == np.float32 and F.k == 2
    assert fe._is_f32(np.zeros((2, 2), np.float32)) and not fe._is_f32(np.zeros((2, 2)))


def test_eltype_wrapper_dispatch_without_gpu():
    """The element-type wrapper of the front-ends (f(A::AbstractMatOrLinOp{T}, opts = LRAOptions(T); ...)): a Float32 A
    gets the LRAOptions(Float32) defaults only when no options are passed, keyword arguments still win, the matrix goes
    through Context.widen_f32, FP64 results are rounded; a Float64 A passes through untouched."""
    import brapprox
    from brapprox import _frontend as fe
    seen = {}

    class FakeCtx:
        def widen_f32(self, A):
            seen["widened"] = A.dtype
            return "device-matrix"

    @fe._eltype
    def f(A, opts=None, ctx=None, **kw):
        o = fe._opts(opts, kw)
        seen["A"], seen["rtol"], seen["orth"] = A, o.rtol, o.pheig_orthtol
        return brapprox.IDPackedV(np.array([1]), np.array([2]), np.ones((1, 1)))

    @fe._eltype
    def g(A, rows, ctx=None):                      # no `opts` parameter (CUR)
        seen["A"] = A
        return np.ones(2)

    e32, e64 = float(np.finfo(np.float32).eps), float(np.finfo(np.float64).eps)
    A32, A64 = np.zeros((3, 2), np.float32), np.zeros((3, 2))
    out = f(A32, ctx=FakeCtx())
    assert seen["A"] == "device-matrix" and seen["widened"] == np.float32 and out.T.dtype == np.float32
    assert seen["rtol"] == 5 * e32 and seen["orth"] == float(np.sqrt(e32))
    f(A32, ctx=FakeCtx(), rtol=1e-3)
    assert seen["rtol"] == 1e-3 and seen["orth"] == float(np.sqrt(e32))
    f(A32, brapprox.LRAOptions(rtol=1e-9), ctx=FakeCtx())
    assert seen["rtol"] == 1e-9 and seen["orth"] == float(np.sqrt(e64))            # explicit options are taken as given
    f(A32, None, ctx=FakeCtx())
    assert seen["rtol"] == 5 * e32
    out = f(A64, ctx=FakeCtx())
    assert seen["A"] is A64 and seen["rtol"] == 5 * e64 and out.T.dtype == np.float64
    assert g(A32, [1], ctx=FakeCtx()).dtype == np.float32 and seen["A"] == "device-matrix"
    with pytest.raises(TypeError):
        f(A32, ctx=FakeCtx(), out=(None, None, None))
