#= BRApprox.jl -- Julia shim over libbrapprox.so (the B200-native sketch-then-factor path).

   NOT EXECUTABLE IN THE BUILD IMAGE (no Julia there): this file documents the reference-side binding a
   maintainer adds; the same C ABI is exercised by the Python host in ../brapprox/ and by tests/.

   It re-defines the hot-path front-ends with the reference's signatures and return types
   (src/id.jl:434-456, src/pqr.jl:285-320, src/psvd.jl:238-299, src/pheig.jl:276-320, src/cur.jl:85-109,527-572,
   src/prange.jl:14-62, src/sketch.jl:52-66) and leaves everything else (result-type arithmetic, LinearOperator)
   to LowRankApprox.jl itself.

   Random inputs, two modes (BRApprox.PARITY[]):
     false (default): n_rounds = 0 -- the library draws with the device Philox generator keyed by BraOpts.seed, which
                      is taken from Julia's task-local RNG, so `Random.seed!` still makes a run reproducible;
     true:            the shim draws what the reference would draw, WHEN the reference would draw it: one round's
                      inputs per adaptive round, lazily (the call is retried with one more round on BRA_ERR_ROUNDS;
                      the rounds already drawn are kept, so Julia's RNG advances exactly as in the reference).
=#
module BRApprox

using LowRankApprox, Random
using LowRankApprox: LRAOptions, IDPackedV, PartialQR, PartialQRFactors, PartialSVD, PartialHermEigen,
                     CURPackedU, HermCURPackedU, chkopts!, chktrans
import LowRankApprox: idfact, id, pqrfact, pqr, psvdfact, psvdvals, psvd, pheigfact, pheigvals, prange,
                      curfact, sketchfact, CUR, HermCUR
using LinearAlgebra: checksquare, ishermitian

const libbra = "libbrapprox.so"
const BRA_MAX_ROUNDS = 24
const BRA_ERR_UNSUPPORTED, BRA_ERR_ROUNDS = 2, 3
const SKETCH = Dict(:none => 0, :randn => 1, :sprn => 2, :srft => 3, :sub => 4)
const F_P, F_T, F_Q, F_R, F_U, F_S, F_VT, F_TAU, F_BSKETCH = 1, 2, 3, 4, 5, 6, 7, 8, 9
const PARITY = Ref(false)

struct BraOpts                      # mirrors `bra_opts` (include/brapprox.h), field for field
  atol::Cdouble; rtol::Cdouble; rank::Int64; nb::Int64
  sketch::Int32; sketch_randn_niter::Int32; sketchfact_adap::Int32; retval_mask::Int32
  maxdet_tol::Cdouble; maxdet_niter::Int64; samp_a::Int64; samp_b::Int64
  seed::UInt64; verb::Int32; flags::Int32    # flags: BRA_OPT_FRESH_SKETCH = 1 (0: nested Gaussian sketches)
  pheig_orthtol::Cdouble
end

struct BraRand
  n_rounds::Int32; reserved::Int32
  omega::Ptr{Ptr{Float64}}; d::Ptr{Ptr{Float64}}; idx::Ptr{Ptr{Int64}}
  perm::Ptr{Ptr{Int64}}; s::Ptr{Ptr{Float64}}; r::Ptr{Ptr{Int64}}
end
const NORAND = BraRand(0, 0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL)

struct BraInfo
  m::Int64; n::Int64; k::Int64; ksvd::Int64; rounds::Int32; reserved::Int32
  orders::NTuple{BRA_MAX_ROUNDS,Int64}; ks::NTuple{BRA_MAX_ROUNDS,Int64}; steps::NTuple{BRA_MAX_ROUNDS,Int64}
end

const CTX = Ref{Ptr{Cvoid}}(C_NULL)

function __init__()
  rc = ccall((:bra_create, libbra), Cint, (Ref{Ptr{Cvoid}}, Cint), CTX, 0)
  rc == 0 || error("bra_create: ", lasterr())
end
lasterr() = unsafe_string(ccall((:bra_last_error, libbra), Cstring, (Ptr{Cvoid},), CTX[]))

# the *_samp closures cannot cross the ABI: evaluate into (a, b), order = a*n + b
function affine(f::Function)
  b = f(0); a = f(1) - b
  all(f(n) == a*n + b for n in (2, 32, 64, 1000)) || throw(ArgumentError("sketchfact_*_samp must be affine"))
  a, b
end
samp(o::LRAOptions) = o.sketch == :srft ? o.sketchfact_srft_samp : o.sketch == :sub ? o.sketchfact_sub_samp :
                      o.sketchfact_randn_samp

function BraOpts(o::LRAOptions)
  a, b = o.sketch in (:sprn, :none) ? (0, 0) : affine(samp(o))
  rv = lowercase(o.pqrfact_retval)
  mask = (occursin("q", rv) ? 1 : 0) | (occursin("r", rv) ? 2 : 0) | (occursin("t", rv) ? 4 : 0)
  BraOpts(o.atol, o.rtol, o.rank, o.nb, SKETCH[o.sketch], o.sketch_randn_niter, o.sketchfact_adap, mask,
          o.maxdet_tol, o.maxdet_niter, a, b, rand(UInt64), o.verb, 0, o.pheig_orthtol)
end

# status -> Julia exception; BRA_ERR_UNSUPPORTED is returned to the caller, which routes to the reference body
function check(rc)
  rc == 0 && return
  rc < 0 && throw(ArgumentError("libbrapprox: argument $(-rc): " * lasterr()))
  error("libbrapprox status $rc: ", lasterr())
end

# ---- parity mode: the random inputs of ONE adaptive round, drawn in the reference's order -------------------
# round t (0-based) has order l = samp(nb * 2^t) (sprn: nb * 2^t); non-adaptive: one round, l = samp(rank) / rank
#   :randn  crandn(T, l, mA)                                   src/util.jl:4, src/sketch.jl:91
#   :srft   d = 2(rand(mA) .> 0.5) .- 1, idx = rand(1:mA, l)    src/sketch.jl:339-361
#   :sprn   perm = randperm(mA); s = vcat((randn(p_i) for i = 1:l)...), p_i = fld(mA - i, l) + 1   src/sketch.jl:575-579
#   :sub    r = rand(1:mA, l)                                  src/sketch.jl:252
struct RoundDraw
  omega::Matrix{Float64}; d::Vector{Float64}; idx::Vector{Int64}
  perm::Vector{Int64}; s::Vector{Float64}; r::Vector{Int64}
end
function round_order(o::LRAOptions, t::Integer)
  n = (o.sketchfact_adap || o.rank < 0) ? o.nb << t : o.rank
  o.sketch == :sprn ? n : samp(o)(n)
end
function draw_round(o::LRAOptions, t::Integer, mA::Integer)
  l = round_order(o, t)
  e, ei, em = Float64[], Int64[], Matrix{Float64}(undef, 0, 0)
  if o.sketch == :randn
    RoundDraw(randn(l, mA), e, ei, ei, e, ei)
  elseif o.sketch == :srft
    x = rand(mA); d = [2.0*(xi > 0.5) - 1.0 for xi in x]
    RoundDraw(em, d, rand(1:mA, l), ei, e, ei)
  elseif o.sketch == :sprn
    perm = randperm(mA)
    s = reduce(vcat, [randn(fld(mA - i, l) + 1) for i = 1:l]; init=Float64[])
    RoundDraw(em, e, ei, perm, s, ei)
  else # :sub
    RoundDraw(em, e, ei, ei, e, rand(1:mA, l))
  end
end

# Runs `call(rnd::BraRand)` (a ccall returning the status).  Fast mode: once, with NORAND.  Parity mode: with the rounds
# drawn so far, one more each time the library answers BRA_ERR_ROUNDS (sketch = :none draws nothing).
function with_random_inputs(call::Function, o::LRAOptions, mA::Integer)
  (!PARITY[] || o.sketch == :none) && return call(NORAND)
  draws = RoundDraw[]
  while true
    push!(draws, draw_round(o, length(draws), mA))
    po = [pointer(x.omega) for x in draws]; pd = [pointer(x.d) for x in draws]; pi = [pointer(x.idx) for x in draws]
    pp = [pointer(x.perm) for x in draws]; ps = [pointer(x.s) for x in draws]; pr = [pointer(x.r) for x in draws]
    rc = GC.@preserve draws po pd pi pp ps pr begin
      k = o.sketch
      call(BraRand(length(draws), 0,
                   k == :randn ? pointer(po) : C_NULL, k == :srft ? pointer(pd) : C_NULL, k == :srft ? pointer(pi) : C_NULL,
                   k == :sprn ? pointer(pp) : C_NULL, k == :sprn ? pointer(ps) : C_NULL, k == :sub ? pointer(pr) : C_NULL))
    end
    rc == BRA_ERR_ROUNDS && length(draws) < BRA_MAX_ROUNDS && continue
    return rc
  end
end

getinfo() = (i = Ref{BraInfo}(); ccall((:bra_get_info, libbra), Cint, (Ptr{Cvoid}, Ref{BraInfo}), CTX[], i); i[])
function fetch!(which::Integer, dst::Array, ld::Integer)
  isempty(dst) && return dst
  check(ccall((:bra_fetch, libbra), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64), CTX[], which, dst, max(ld, 1)))
  dst
end
tchar(trans::Symbol) = trans == :n ? 'n' : trans == :c ? 'c' : 'b'

# ---- idfact / id (src/id.jl:434-456) -------------------------------------------------------------------------
function idfact(trans::Symbol, A::Matrix{Float64}, opts::LRAOptions=LRAOptions(); args...)
  chktrans(trans)
  opts = copy(opts; args...)
  opts.pqrfact_retval = "t"
  chkopts!(opts, A)
  m, n = size(A)
  rc = with_random_inputs(opts, trans == :n ? m : n) do rnd
    ccall((:bra_idfact_f64, libbra), Cint,
          (Ptr{Cvoid}, Cchar, Int64, Int64, Ptr{Float64}, Int64, Ref{BraOpts}, Ref{BraRand}),
          CTX[], tchar(trans), m, n, A, stride(A, 2), BraOpts(opts), rnd)
  end
  rc == BRA_ERR_UNSUPPORTED && return invoke(idfact, Tuple{Symbol,AbstractMatrix,LRAOptions}, trans, A, opts)
  check(rc)
  i = getinfo()
  p = fetch!(F_P, Vector{Int}(undef, i.n), i.n)
  T = fetch!(F_T, Matrix{Float64}(undef, i.k, i.n - i.k), i.k)
  IDPackedV(p[1:i.k], p[i.k+1:end], T)                    # src/id.jl:445-446
end
idfact(A::Matrix{Float64}, args...; kwargs...) = idfact(:n, A, args...; kwargs...)
id(trans::Symbol, A::Matrix{Float64}, args...; kwargs...) = (V = idfact(trans, A, args...; kwargs...); (V[:sk], V[:rd], V[:T]))

# ---- pqrfact / pqr (src/pqr.jl:285-320): PartialQR for retval "qr", PartialQRFactors otherwise (:434-435) ------
function pqrfact(trans::Symbol, A::Matrix{Float64}, opts::LRAOptions=LRAOptions(); args...)
  chktrans(trans)
  opts = copy(opts; args...)
  chkopts!(opts, A)
  m, n = size(A)
  rc = with_random_inputs(opts, trans == :n ? m : n) do rnd
    ccall((:bra_pqrfact_f64, libbra), Cint,
          (Ptr{Cvoid}, Cchar, Int64, Int64, Ptr{Float64}, Int64, Ref{BraOpts}, Ref{BraRand}),
          CTX[], tchar(trans), m, n, A, stride(A, 2), BraOpts(opts), rnd)
  end
  rc == BRA_ERR_UNSUPPORTED && return invoke(pqrfact, Tuple{Symbol,AbstractMatrix,LRAOptions}, trans, A, opts)
  check(rc)
  i = getinfo()
  rv = lowercase(opts.pqrfact_retval)
  retq, retr, rett = occursin("q", rv), occursin("r", rv), occursin("t", rv)
  p = fetch!(F_P, Vector{Int}(undef, i.n), i.n)
  Q = retq ? fetch!(F_Q, Matrix{Float64}(undef, i.m, i.k), i.m) : nothing
  R = retr ? fetch!(F_R, Matrix{Float64}(undef, i.k, i.n), i.k) : nothing
  T = rett ? fetch!(F_T, Matrix{Float64}(undef, i.k, i.n - i.k), i.k) : nothing
  retq && retr && !rett && return PartialQR(Q, R, p)
  PartialQRFactors(Q, R, p, Int(i.k), T)
end
pqrfact(A::Matrix{Float64}, args...; kwargs...) = pqrfact(:n, A, args...; kwargs...)
function pqr(trans::Symbol, A::Matrix{Float64}, args...; kwargs...)
  F = pqrfact(trans, A, args...; kwargs...)
  F[:Q], F[:R], F[:p]
end

# ---- sketchfact (src/sketch.jl:52-66): the factors of the sketch itself ------------------------------------------
function sketchfact(side::Symbol, trans::Symbol, A::Matrix{Float64}, opts::LRAOptions=LRAOptions(); args...)
  side in (:left, :right) || throw(ArgumentError("side"))
  chktrans(trans)
  opts = copy(opts; args...)
  chkopts!(opts, A)
  m, n = size(A)
  mA = side == :left ? (trans == :n ? m : n) : (trans == :n ? n : m)
  rc = with_random_inputs(opts, mA) do rnd
    ccall((:bra_sketchfact_f64, libbra), Cint,
          (Ptr{Cvoid}, Cchar, Cchar, Int64, Int64, Ptr{Float64}, Int64, Ref{BraOpts}, Ref{BraRand}),
          CTX[], side == :left ? 'l' : 'r', tchar(trans), m, n, A, stride(A, 2), BraOpts(opts), rnd)
  end
  check(rc)
  i = getinfo()
  rows = side == :left ? i.orders[i.rounds] : i.m
  p = fetch!(F_P, Vector{Int}(undef, i.n), i.n)
  B = fetch!(F_BSKETCH, Matrix{Float64}(undef, rows, i.n), rows)       # LAPACK layout: R above, reflectors below
  tau = fetch!(F_TAU, Vector{Float64}(undef, i.steps[i.rounds]), 1)
  LowRankApprox.pqrback_postproc(B, p, tau, Int(i.k), opts)             # orgqr / triu / maxdet_t on the small sketch
end

# ---- psvdfact / psvdvals / psvd (src/psvd.jl:238-299) ---------------------------------------------------------
function psvd_call(A::Matrix{Float64}, opts::LRAOptions; vals_only::Bool=false)
  m, n = size(A)
  with_random_inputs(opts, max(m, n)) do rnd          # trans = :n if m >= n else :c (src/psvd.jl:242,256)
    if vals_only                                      # psvdvals: no Q, no singular vectors
      ccall((:bra_psvdvals_f64, libbra), Cint,
            (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Ref{BraOpts}, Ref{BraRand}),
            CTX[], m, n, A, stride(A, 2), BraOpts(opts), rnd)
    else
      ccall((:bra_psvdfact_f64, libbra), Cint,
            (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Ref{BraOpts}, Ref{BraRand}),
            CTX[], m, n, A, stride(A, 2), BraOpts(opts), rnd)
    end
  end
end
function psvdfact(A::Matrix{Float64}, opts::LRAOptions=LRAOptions(); args...)
  opts = copy(opts; args...)
  chkopts!(opts, A)
  rc = psvd_call(A, opts)
  rc == BRA_ERR_UNSUPPORTED && return invoke(psvdfact, Tuple{AbstractMatrix,LRAOptions}, A, opts)
  check(rc)
  m, n = size(A); k = getinfo().ksvd
  PartialSVD(fetch!(F_U, Matrix{Float64}(undef, m, k), m), fetch!(F_S, Vector{Float64}(undef, k), k),
             fetch!(F_VT, Matrix{Float64}(undef, k, n), k))
end
function psvdvals(A::Matrix{Float64}, opts::LRAOptions=LRAOptions(); args...)
  opts = copy(opts; args...)
  chkopts!(opts, A)
  check(psvd_call(A, opts; vals_only=true))
  k = getinfo().ksvd
  fetch!(F_S, Vector{Float64}(undef, k), k)
end
psvd(A::Matrix{Float64}, args...; kwargs...) = (F = psvdfact(A, args...; kwargs...); (F.U, F.S, F.Vt'))

# ---- pheigfact / pheigvals (src/pheig.jl:276-311) --------------------------------------------------------------
function pheig_call(A::Matrix{Float64}, opts::LRAOptions)
  n = checksquare(A)
  rc = with_random_inputs(opts, n) do rnd
    ccall((:bra_pheigfact_f64, libbra), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Int64, Ref{BraOpts}, Ref{BraRand}),
          CTX[], n, A, stride(A, 2), BraOpts(opts), rnd)
  end
  rc == -3 && error("matrix must be Hermitian")           # src/pheig.jl:279
  check(rc)
  n, getinfo().ksvd
end
function pheigfact(A::Matrix{Float64}, opts::LRAOptions=LRAOptions(); args...)
  opts = isempty(args) ? opts : copy(opts; args...)
  n, k = pheig_call(A, opts)
  PartialHermEigen(fetch!(F_S, Vector{Float64}(undef, k), k), fetch!(F_U, Matrix{Float64}(undef, n, k), n))
end
function pheigvals(A::Matrix{Float64}, opts::LRAOptions=LRAOptions(); args...)
  opts = isempty(args) ? opts : copy(opts; args...)
  _, k = pheig_call(A, opts)
  fetch!(F_S, Vector{Float64}(undef, k), k)
end

# ---- curfact (src/cur.jl:527-572): the reference's control flow over the device idfact; CUR / HermCUR (:85-109) --
function curfact(A::Matrix{Float64}, opts::LRAOptions=LRAOptions(); args...)
  opts = copy(opts; args...)
  opts.pqrfact_retval = ""
  chkopts!(opts, A)
  if ishermitian(A)
    return HermCURPackedU(idfact(:n, A, opts)[:sk])
  end
  m, n = size(A)
  if m >= n
    rows = idfact(:c, A, opts)[:sk]
    cols = idfact(:n, A[rows,:], opts)[:sk]
  else
    cols = idfact(:n, A, opts)[:sk]
    rows = idfact(:c, A[:,cols], opts)[:sk]
  end
  k = min(length(rows), length(cols))
  CURPackedU(rows[1:k], cols[1:k])
end
function cur_call(A::Matrix{Float64}, rows::Vector{Int}, cols::Vector{Int}, herm::Bool)
  m, n = size(A)
  check(ccall((:bra_cur_f64, libbra), Cint,
              (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Cint),
              CTX[], m, n, A, stride(A, 2), length(cols), rows, cols, herm))
  k = length(cols)
  C = fetch!(F_Q, Matrix{Float64}(undef, m, k), m)
  s = fetch!(F_S, Vector{Float64}(undef, k), k)
  V = fetch!(F_U, Matrix{Float64}(undef, k, k), k)
  C, s, V
end
function CUR(A::Matrix{Float64}, U::CURPackedU)
  C, sinv, V = cur_call(A, U[:rows], U[:cols], false)
  k, n = length(sinv), size(A, 2)
  Ut = fetch!(F_VT, Matrix{Float64}(undef, k, k), k)
  R = fetch!(F_R, Matrix{Float64}(undef, k, n), k)
  LowRankApprox.CUR(U[:rows], U[:cols], C, PartialSVD(V, sinv, Ut), R)     # src/cur.jl:91-92
end
function HermCUR(A::Matrix{Float64}, U::HermCURPackedU)
  C, winv, X = cur_call(A, U[:cols], U[:cols], true)
  LowRankApprox.HermCUR(U[:cols], C, PartialHermEigen(winv, X))            # src/cur.jl:103-104
end

# ---- prange (src/prange.jl:14-62) ------------------------------------------------------------------------------
function prange(trans::Symbol, A::Matrix{Float64}, opts::LRAOptions=LRAOptions(); args...)
  trans in (:n, :c, :b) || throw(ArgumentError("trans"))
  opts = copy(opts; args...)
  chkopts!(opts, A)
  m, n = size(A)
  # fast mode only: trans = :b takes two independent sets of draws (rnd for A', rnd2 for A)
  rc = ccall((:bra_prange_f64, libbra), Cint,
             (Ptr{Cvoid}, Cchar, Int64, Int64, Ptr{Float64}, Int64, Ref{BraOpts}, Ref{BraRand}, Ref{BraRand}),
             CTX[], tchar(trans), m, n, A, stride(A, 2), BraOpts(opts), NORAND, NORAND)
  rc == -3 && throw(DimensionMismatch("matrix is not square"))          # checksquare, src/prange.jl:25
  check(rc)
  i = getinfo()
  fetch!(F_Q, Matrix{Float64}(undef, i.m, i.k), i.m)
end

# snormdiff(A, L*R) with A, L, R already on the device (CuArray pointers): src/snorm.jl:14-53
function snormdiff_device(m, n, dA::Ptr{Float64}, lda, k, dL::Ptr{Float64}, ldl, dR::Ptr{Float64}, ldr, opts::LRAOptions)
  res = Ref{Cdouble}(0); nit = Ref{Int64}(0)
  check(ccall((:bra_snorm_f64, libbra), Cint,
              (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64,
               Ref{BraOpts}, Int64, Ptr{Float64}, Ref{Cdouble}, Ref{Int64}),
              CTX[], m, n, dA, lda, k, dL, ldl, dR, ldr, BraOpts(opts), opts.snorm_niter, C_NULL, res, nit))
  res[]
end

end # module
