#= BRApprox.jl -- Julia shim over libbrapprox.so (the B200-native sketch-then-factor path).

   NOT EXECUTABLE IN THE BUILD IMAGE (no Julia there): this file documents the reference-side binding a
   maintainer adds; the same C ABI is exercised by the Python host in ../brapprox/ and by tests/.

   It re-defines the hot-path front-ends with the reference's signatures and return types
   (src/id.jl:434-456, src/pqr.jl:285-320, src/psvd.jl:238-299) and leaves everything else
   (result-type arithmetic, LinearOperator, CUR/pheig bodies) to LowRankApprox.jl itself.
=#
module BRApprox

using LowRankApprox
using LowRankApprox: LRAOptions, IDPackedV, chkopts!, chktrans
import LowRankApprox: idfact, pqrfact, psvdfact, prange

const libbra = "libbrapprox.so"
const BRA_MAX_ROUNDS = 24
const SKETCH = Dict(:none => 0, :randn => 1, :sprn => 2, :srft => 3, :sub => 4)

struct BraOpts
  atol::Cdouble; rtol::Cdouble; rank::Int64; nb::Int64
  sketch::Int32; sketch_randn_niter::Int32; sketchfact_adap::Int32; retval_mask::Int32
  maxdet_tol::Cdouble; maxdet_niter::Int64; samp_a::Int64; samp_b::Int64
  seed::UInt64; verb::Int32; reserved::Int32
end

struct BraRand
  n_rounds::Int32; reserved::Int32
  omega::Ptr{Ptr{Float64}}; d::Ptr{Ptr{Float64}}; idx::Ptr{Ptr{Int64}}
  perm::Ptr{Ptr{Int64}}; s::Ptr{Ptr{Float64}}; r::Ptr{Ptr{Int64}}
end

struct BraInfo
  m::Int64; n::Int64; k::Int64; ksvd::Int64; rounds::Int32; reserved::Int32
  orders::NTuple{BRA_MAX_ROUNDS,Int64}; ks::NTuple{BRA_MAX_ROUNDS,Int64}; steps::NTuple{BRA_MAX_ROUNDS,Int64}
end

const CTX = Ref{Ptr{Cvoid}}(C_NULL)

function __init__()
  rc = ccall((:bra_create, libbra), Cint, (Ref{Ptr{Cvoid}}, Cint), CTX, 0)
  rc == 0 || error("bra_create: ", unsafe_string(ccall((:bra_last_error, libbra), Cstring, (Ptr{Cvoid},), CTX[])))
end

# the *_samp closures cannot cross the ABI: evaluate into (a, b), order = a*n + b
function affine(f::Function)
  b = f(0); a = f(1) - b
  all(f(n) == a*n + b for n in (2, 32, 64, 1000)) || throw(ArgumentError("sketchfact_*_samp must be affine"))
  a, b
end

function BraOpts(o::LRAOptions)
  f = o.sketch == :srft ? o.sketchfact_srft_samp : o.sketch == :sub ? o.sketchfact_sub_samp : o.sketchfact_randn_samp
  a, b = o.sketch == :sprn ? (0, 0) : affine(f)
  rv = lowercase(o.pqrfact_retval)
  mask = (occursin("q", rv) ? 1 : 0) | (occursin("r", rv) ? 2 : 0) | (occursin("t", rv) ? 4 : 0)
  BraOpts(o.atol, o.rtol, o.rank, o.nb, SKETCH[o.sketch], o.sketch_randn_niter, o.sketchfact_adap, mask,
          o.maxdet_tol, o.maxdet_niter, a, b, rand(UInt64), o.verb, 0)
end

throw_bra(rc) = rc < 0 ? throw(ArgumentError("libbrapprox: argument $(-rc)")) :
  error("libbrapprox status $rc: ", unsafe_string(ccall((:bra_last_error, libbra), Cstring, (Ptr{Cvoid},), CTX[])))

# Draw the Omegas the reference would draw (crandn, src/util.jl:4), one per adaptive round.
function draw_omegas(o::LRAOptions, mA::Integer, maxrounds::Integer=8)
  n = o.nb
  Ωs = Matrix{Float64}[]
  for _ = 1:maxrounds
    push!(Ωs, randn(o.sketchfact_randn_samp(n), mA)); n *= 2
  end
  Ωs
end

function idfact(trans::Symbol, A::Matrix{Float64}, opts::LRAOptions=LRAOptions(); args...)
  chktrans(trans)
  opts = copy(opts; args...)
  opts.pqrfact_retval = "t"
  chkopts!(opts, A)
  # every LRAOptions combination of the Float64 dense path runs on the device: all five sketches (:none included),
  # maxdet_tol / maxdet_niter (src/pqr.jl:444-501) and sketch_randn_niter (src/sketch.jl:140-149)
  m, n = size(A)
  Ωs = draw_omegas(opts, trans == :n ? m : n)
  ptrs = [pointer(Ω) for Ω in Ωs]
  GC.@preserve Ωs ptrs begin
    rnd = BraRand(length(Ωs), 0, pointer(ptrs), C_NULL, C_NULL, C_NULL, C_NULL, C_NULL)
    rc = ccall((:bra_idfact_f64, libbra), Cint,
               (Ptr{Cvoid}, Cchar, Int64, Int64, Ptr{Float64}, Int64, Ref{BraOpts}, Ref{BraRand}),
               CTX[], trans == :n ? 'n' : 'c', m, n, A, stride(A, 2), BraOpts(opts), rnd)
  end
  rc == 0 || throw_bra(rc)
  info = Ref{BraInfo}()
  ccall((:bra_get_info, libbra), Cint, (Ptr{Cvoid}, Ref{BraInfo}), CTX[], info)
  k, nn = info[].k, info[].n
  p = Vector{Int}(undef, nn)
  ccall((:bra_fetch, libbra), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64), CTX[], 1, p, nn)
  T = Matrix{Float64}(undef, k, nn - k)
  k > 0 && nn > k && ccall((:bra_fetch, libbra), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64), CTX[], 2, T, k)
  IDPackedV(p[1:k], p[k+1:end], T)                        # src/id.jl:445-446
end

# ---- the other front-ends, in FAST mode (n_rounds = 0: the library draws its random inputs with the device Philox
# generator keyed by BraOpts.seed; pass drawn inputs as in idfact above to keep `Random.seed!` reproducibility) ----

const NORAND = BraRand(0, 0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL)

function getinfo()
  info = Ref{BraInfo}()
  ccall((:bra_get_info, libbra), Cint, (Ptr{Cvoid}, Ref{BraInfo}), CTX[], info)
  info[]
end

function fetch!(which::Integer, dst::Array, ld::Integer)
  isempty(dst) && return dst
  rc = ccall((:bra_fetch, libbra), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64), CTX[], which, dst, ld)
  rc == 0 || throw_bra(rc)
  dst
end

function pqrfact(trans::Symbol, A::Matrix{Float64}, opts::LRAOptions=LRAOptions(); args...)   # src/pqr.jl:290-307
  chktrans(trans)
  opts = copy(opts; args...)
  chkopts!(opts, A)
  m, n = size(A)
  rc = ccall((:bra_pqrfact_f64, libbra), Cint,
             (Ptr{Cvoid}, Cchar, Int64, Int64, Ptr{Float64}, Int64, Ref{BraOpts}, Ref{BraRand}),
             CTX[], trans == :n ? 'n' : 'c', m, n, A, stride(A, 2), BraOpts(opts), NORAND)
  rc == 0 || throw_bra(rc)
  i = getinfo()
  p = fetch!(1, Vector{Int}(undef, i.n), i.n)
  Q = fetch!(3, Matrix{Float64}(undef, i.m, i.k), i.m)
  R = fetch!(4, Matrix{Float64}(undef, i.k, i.n), max(i.k, 1))
  LowRankApprox.PartialQR(Q, R, p)
end

function psvdfact(A::Matrix{Float64}, opts::LRAOptions=LRAOptions(); args...)                  # src/psvd.jl:238-272
  opts = copy(opts; args...)
  chkopts!(opts, A)
  m, n = size(A)
  rc = ccall((:bra_psvdfact_f64, libbra), Cint,
             (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Ref{BraOpts}, Ref{BraRand}),
             CTX[], m, n, A, stride(A, 2), BraOpts(opts), NORAND)
  rc == 0 || throw_bra(rc)
  k = getinfo().ksvd
  U  = fetch!(5, Matrix{Float64}(undef, m, k), m)
  S  = fetch!(6, Vector{Float64}(undef, k), k)
  Vt = fetch!(7, Matrix{Float64}(undef, k, n), max(k, 1))
  LowRankApprox.PartialSVD(U, S, Vt)
end

function prange(trans::Symbol, A::Matrix{Float64}, opts::LRAOptions=LRAOptions(); args...)     # src/prange.jl:14-62
  trans in (:n, :c, :b) || throw(ArgumentError("trans"))
  opts = copy(opts; args...)
  chkopts!(opts, A)
  m, n = size(A)
  rc = ccall((:bra_prange_f64, libbra), Cint,
             (Ptr{Cvoid}, Cchar, Int64, Int64, Ptr{Float64}, Int64, Ref{BraOpts}, Ref{BraRand}, Ref{BraRand}),
             CTX[], Char(string(trans)[1]), m, n, A, stride(A, 2), BraOpts(opts), NORAND, NORAND)
  rc == 0 || throw_bra(rc)
  i = getinfo()
  fetch!(3, Matrix{Float64}(undef, i.m, i.k), i.m)
end

# snormdiff(A, L*R) with A, L, R already on the device (CuArray pointers): src/snorm.jl:14-53
function snormdiff_device(m, n, dA::Ptr{Float64}, lda, k, dL::Ptr{Float64}, ldl, dR::Ptr{Float64}, ldr, opts::LRAOptions)
  res = Ref{Cdouble}(0); nit = Ref{Int64}(0)
  rc = ccall((:bra_snorm_f64, libbra), Cint,
             (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64,
              Ref{BraOpts}, Int64, Ptr{Float64}, Ref{Cdouble}, Ref{Int64}),
             CTX[], m, n, dA, lda, k, dL, ldl, dR, ldr, BraOpts(opts), opts.snorm_niter, C_NULL, res, nit)
  rc == 0 || throw_bra(rc)
  res[]
end

end # module
