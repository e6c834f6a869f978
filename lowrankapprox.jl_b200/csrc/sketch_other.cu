// SRFT, sparse-Gaussian and random-subset sketches, (:left, :n) and (:left, :c) forms.
// Reference: src/sketch.jl:334-530 (SRFT: srft_init, srft_reshape!, srft_apply!, mul!(::SRFT)),
//            src/sketch.jl:566-657 (SparseRandGauss), src/sketch.jl:244-296 (RandomSubset).
//
// B200 design:
//  * SRFT: one CTA per (column of op(A), chunk of the m' = m/l "inner" index).  The reference reshapes a column x
//    into X[j,k] = d[i] x[i], i = j m' + k (l x m'), runs FFTW R2HC along j and then evaluates the sampled
//    frequencies f = c + l r with m' twiddled terms each.  Here the l x m'_c chunk is staged in shared memory
//    straight in the reshaped order (it IS the column, sign-flipped), read as l x (m'_c/2) COMPLEX numbers
//    (two real inner columns per complex one), transformed along j with in-place radix-4/2 DIF passes, and the
//    sampled frequencies are evaluated from the digit-reversed bins with the even/odd separation
//    X^[c,2q] = (Z_q[c] + conj Z_q[l-c])/2, X^[c,2q+1] = -i (Z_q[c] - conj Z_q[l-c])/2.  Chunks (needed when a
//    column does not fit in shared memory, and to get several CTAs per SM) produce partial sums that are added
//    in a fixed order.  Twiddles come from sincospi-built tables (the reference multiplies them up, drifting by
//    O(m' eps)); the passes are radix 8 (then 4 or 2), so a length-512 transform touches shared memory 3 times.
//    Shapes the radix passes do not cover (l not a power of two, odd m') go through an explicitly generated
//    SRFT matrix and the TMA + DMMA GEMM.
//  * sparse Gaussian: every entry of A is read exactly once; a column is staged in shared memory (coalesced)
//    and the order-many weighted sums are formed from there in the reference's summation order.
//  * random subset: a gather.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>

bool bra_make_map_3d_f64(CUtensorMap* map, const double* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1,
                         uint64_t s2, uint32_t b0, uint32_t b1);
int bra_splitk_reduce(bra_ctx* ctx, const double* part, int64_t split_stride, int splits, int64_t l, int64_t n,
                      double* out, int64_t ldo);

namespace {

// ------------------------------------------------------------------ random subset
// B[i, j] = op(A)[r_i, j]   (src/sketch.jl:248-257; trans 'c': :268-279)
__global__ void sketch_sub_kernel(char trans, const double* __restrict__ A, int64_t lda, int64_t nA, int64_t order,
                                  const int64_t* __restrict__ r1, double* __restrict__ B, int64_t ldb) {
  const int64_t total = order * nA;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e % order, j = e / order;
    const int64_t r = r1[i] - 1;
    B[i + j * ldb] = (trans == 'n') ? A[r + j * lda] : A[j + r * lda];
  }
}

// ------------------------------------------------------------------ sparse random Gaussian
// Row i of B = sum_{t < p_i} s[off_i + t] * op(A)[perm[off_i + t], :],  p_i = fld(m - i, order) + 1 (i 1-based),
// accumulated in t order like the reference (src/sketch.jl:571-589).
// One CTA per column j of op(A): the column is staged in shared memory when it fits (each entry of A is read
// from HBM exactly once, coalesced); thread i forms sketch row i.
__global__ void __launch_bounds__(256) sketch_sprn_kernel(char trans, const double* __restrict__ A, int64_t lda,
                                                          int64_t mA, int64_t nA, int64_t order,
                                                          const int64_t* __restrict__ perm1,
                                                          const double* __restrict__ s, double* __restrict__ B,
                                                          int64_t ldb, int stage_col) {
  extern __shared__ double col[];
  const int64_t q = mA / order, rem = mA % order;       // p_i = q + (i0 < rem), off_i = i0*q + min(i0, rem), i0 = i-1
  for (int64_t j = blockIdx.x; j < nA; j += gridDim.x) {
    const double* a = (trans == 'n') ? A + j * lda : A + j;
    const int64_t astride = (trans == 'n') ? 1 : lda;
    if (stage_col) {
      __syncthreads();
      for (int64_t r = threadIdx.x; r < mA; r += blockDim.x) col[r] = a[r * astride];
      __syncthreads();
    }
    for (int64_t i0 = threadIdx.x; i0 < order; i0 += blockDim.x) {
      const int64_t p = q + (i0 < rem ? 1 : 0);
      const int64_t off = i0 * q + (i0 < rem ? i0 : rem);
      double acc = 0.0;
      for (int64_t t = 0; t < p; ++t) {
        const int64_t row = perm1[off + t] - 1;
        const double x = stage_col ? col[row] : a[row * astride];
        acc += s[off + t] * x;                   // same order and (unfused) rounding as the reference loop
      }
      B[i0 + j * ldb] = acc;
    }
  }
}

// ------------------------------------------------------------------ SRFT
struct SrftParams {
  const double* A;        // op(A) column-major mA x nA (already transposed for trans 'c')
  int64_t lda;
  int64_t m, n;           // contracted length, number of columns
  int order;              // sketch rows
  int l, logl;            // FFT length (power of two)
  int mp;                 // m / l
  int mc;                 // inner indices per chunk (even), mp % mc == 0
  int tsplit, nhi;        // twiddle tables: exp(-2 pi i t/m) = Hi[t / tsplit] * Lo[t % tsplit], nhi = ceil(m / tsplit)
  const double* d;        // +-1, length m
  const uint32_t* dbits;  // bit i set <=> d[i] < 0 (m bits, padded to a word)
  int use_tma;            // stage the l x mc chunk with tiled TMA loads (3-D map over (kk, j, column))
  int boxrows;            // rows per TMA box (<= 256)
  const int64_t* idx1;    // 1-based sampled frequencies, length order
  double* out;            // order x n (ld = ldo) [+ chunk * split_stride]
  int64_t ldo, split_stride;
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}

// position of frequency bin c after the in-place DIF passes (radix 8 while >= 3 stages remain, then 4 or 2)
__device__ __forceinline__ int srft_binpos(int c, int l, int logl) {
  int pos = 0, N = l, rem = logl;
  while (rem >= 3) {
    pos += (c & 7) * (N >> 3);
    c >>= 3;
    N >>= 3;
    rem -= 3;
  }
  if (rem == 2) pos += (c & 3) * (N >> 2);
  else if (rem == 1) pos += (c & 1) * (N >> 1);
  return pos;
}

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmuli_neg(double2 a) { return make_double2(a.y, -a.x); }     // -i * a

// 4-point forward DFT: y_k = sum_t u_t (-i)^{kt}
__device__ __forceinline__ void dft4(double2 u0, double2 u1, double2 u2, double2 u3, double2& y0, double2& y1,
                                     double2& y2, double2& y3) {
  const double2 s02 = cadd(u0, u2), d02 = csub(u0, u2), s13 = cadd(u1, u3), d13 = cmuli_neg(csub(u1, u3));
  y0 = cadd(s02, s13);
  y2 = csub(s02, s13);
  y1 = cadd(d02, d13);
  y3 = csub(d02, d13);
}

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)),
               "l"(src), "r"(bytes), "r"(s_u32(bar))
               : "memory");
}
__device__ __forceinline__ double flip(double x, uint32_t neg) {      // neg in {0, 1}
  return __hiloint2double(__double2hiint(x) ^ (int)(neg << 31), __double2loint(x));
}

__global__ void srft_signbits_kernel(const double* __restrict__ d, int64_t m, uint32_t* __restrict__ bits) {
  const int64_t words = (m + 31) / 32;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (int64_t)gridDim.x * blockDim.x) {
    uint32_t v = 0;
    for (int b = 0; b < 32; ++b) {
      const int64_t i = w * 32 + b;
      if (i < m && d[i] < 0.0) v |= 1u << b;
    }
    bits[w] = v;
  }
}

__global__ void __launch_bounds__(256, 3) srft_kernel(const __grid_constant__ CUtensorMap mapA, SrftParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int l = P.l, mc = P.mc, Qc = mc >> 1;
  double* Xs = reinterpret_cast<double*>(smem_raw);                       // l * mc doubles = l * Qc complex
  double2* Z = reinterpret_cast<double2*>(smem_raw);
  double2* W = reinterpret_cast<double2*>(Xs + (size_t)l * mc);          // W[t] = exp(-2 pi i t / l), t < l
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  const int chunk = blockIdx.y;
  const int kk0 = chunk * mc;
  const int nsamp = (P.order + 1) >> 1;

  for (int t = tid; t < l; t += nthr) {
    double sn, cs;
    sincospi(-2.0 * (double)t / (double)l, &sn, &cs);
    W[t] = make_double2(cs, sn);
  }
  // twiddles of the sampled-frequency stage, exp(-2 pi i t / m), from two small tables (one complex multiply
  // per twiddle instead of a sincospi): Lo[t] = exp(-2 pi i t/m), t < tsplit; Hi[u] = exp(-2 pi i u tsplit/m)
  double2* Lo = W + l;
  double2* Hi = Lo + P.tsplit;
  for (int t = tid; t < P.tsplit + P.nhi; t += nthr) {
    const int64_t tt = (t < P.tsplit) ? t : (int64_t)(t - P.tsplit) * P.tsplit;
    double sn, cs;
    sincospi(-2.0 * (double)(tt % P.m) / (double)P.m, &sn, &cs);
    W[l + t] = make_double2(cs, sn);
  }
  const int tshift = 31 - __clz(P.tsplit);          // tsplit is a power of two
  auto twiddle = [&](int64_t t) -> double2 {        // t in [0, m)
    return cmul(Hi[(int)(t >> tshift)], Lo[(int)(t & (P.tsplit - 1))]);
  };
  // per-sample constants (the same for every column): frequency, bin positions of c and l - c
  int* sf = reinterpret_cast<int*>(Hi + P.nhi);      // [nsamp] f
  int* spc = sf + nsamp;                             // [nsamp] position of bin c
  int* spn = spc + nsamp;                            // [nsamp] position of bin (l - c) mod l
  for (int sp = tid; sp < nsamp; sp += nthr) {
    const int64_t f = P.idx1[2 * sp] - 1;
    const int c = (int)(f % l);
    sf[sp] = (int)f;
    spc[sp] = srft_binpos(c, l, P.logl);
    spn[sp] = srft_binpos((l - c) & (l - 1), l, P.logl);
  }
  // sign bits of this chunk's rows and the mbarrier of the bulk loads
  uint32_t* sbits = reinterpret_cast<uint32_t*>(spn + nsamp);           // [(m + 31) / 32]
  const int nwords = (int)((P.m + 31) / 32);
  uint64_t* bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sbits + nwords) + 7) & ~uintptr_t(7));
  for (int w = tid; w < nwords; w += nthr) sbits[w] = P.dbits[w];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t phase = 0;
  const bool mpow2 = (P.m & (P.m - 1)) == 0;
  // lanes of a warp are split into groups of GS lanes; a group evaluates one sampled frequency
  int GS = 1;
  while (GS < 32 && GS < Qc) GS <<= 1;
  const int gpw = 32 / GS, sub = lane / GS, ql = lane % GS;
  // division-free work distribution over (j or butterfly index, q): item w = tid + k*nthr <-> (w / Qc, w % Qc)
  const int q0 = tid % Qc, b0 = tid / Qc, dq = nthr % Qc, db = nthr / Qc;
  const int t0 = tid % mc, j0 = tid / mc, dt = nthr % mc, dj = nthr / mc;

  for (int64_t col = blockIdx.x; col < P.n; col += gridDim.x) {
    const double* x = P.A + col * P.lda;
    __syncthreads();
    // ---- srft_reshape!: X[j, kk] = d[i] x[i], i = j*mp + kk, restricted to this chunk of kk.  TMA path: the
    // l x mc tile arrives by tiled TMA loads (<= 256 rows per box), the sign flip is folded into the first pass ----
    if (P.use_tma) {
      if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)),
                     "r"((uint32_t)(l * mc * 8))
                     : "memory");
        for (int jr = 0; jr < l; jr += P.boxrows)
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
                  "r"(s_u32(Xs + (size_t)jr * mc)),
              "l"(&mapA), "r"(kk0), "r"(jr), "r"((int)col), "r"(s_u32(bar))
              : "memory");
      }
      uint32_t ok;
      do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(s_u32(bar)), "r"(phase)
            : "memory");
      } while (!ok);
      phase ^= 1;
    } else {
      int j = j0, t = t0;
      for (int e = tid; e < l * mc; e += nthr) {
        const int64_t i = (int64_t)j * P.mp + kk0 + t;
        Xs[e] = P.d[i] * x[i];
        t += dt;
        j += dj;
        if (t >= mc) {
          t -= mc;
          ++j;
        }
      }
      __syncthreads();
    }
    bool first = P.use_tma != 0;        // the first pass still has to apply d
    auto ldz = [&](const double2* z, int row, int q) -> double2 {
      double2 v = *z;
      if (first) {
        const int64_t i = (int64_t)row * P.mp + kk0 + 2 * q;          // even: both signs sit in one word
        const uint32_t wbits = sbits[i >> 5] >> (i & 31);
        v.x = flip(v.x, wbits & 1u);
        v.y = flip(v.y, (wbits >> 1) & 1u);
      }
      return v;
    };
    // ---- length-l DFT along j of the Qc complex columns, in place, decimation in frequency ----
    int N = l, rem = P.logl;
    while (rem >= 3) {
      // radix 8: y_s = (sum_t x_t W_8^{st}) W_N^{sj}, stored at b*N + s*(N/8) + j
      const int h = N >> 3, tw = l / N, hmask = h - 1, hshift = 31 - __clz(h);
      const int items = (l >> 3) * Qc;
      int q = q0, bf = b0;
      for (int w = tid; w < items; w += nthr) {
        const int b = bf >> hshift, j = bf & hmask;
        double2* z = Z + ((size_t)(b * N + j)) * Qc + q;
        const size_t hs = (size_t)h * Qc;
        double2 v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = ldz(z + t * hs, b * N + j + t * h, q);
        // first radix-2 stage: a_t = x_t + x_{t+4}, b_t = (x_t - x_{t+4}) W_8^t
        const double r = 0.70710678118654752440;
        const double2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
        const double2 e0 = csub(v[0], v[4]), e1 = csub(v[1], v[5]), e2 = csub(v[2], v[6]), e3 = csub(v[3], v[7]);
        const double2 c0 = e0;
        const double2 c1 = make_double2(r * (e1.x + e1.y), r * (e1.y - e1.x));       // * (1 - i)/sqrt2
        const double2 c2 = cmuli_neg(e2);                                            // * (-i)
        const double2 c3 = make_double2(r * (e3.y - e3.x), -r * (e3.x + e3.y));      // * (-1 - i)/sqrt2
        double2 y[8];
        dft4(a0, a1, a2, a3, y[0], y[2], y[4], y[6]);
        dft4(c0, c1, c2, c3, y[1], y[3], y[5], y[7]);
        z[0] = y[0];
        if (h == 1) {
          // last stage of the decimation (N == 8): every twiddle is W^0 = 1 -- no table reads, no multiplies
#pragma unroll
          for (int sI = 1; sI < 8; ++sI) z[sI * hs] = y[sI];
        } else {
          const int jt = j * tw;
#pragma unroll
          for (int sI = 1; sI < 8; ++sI) z[sI * hs] = cmul(y[sI], W[(sI * jt) & (l - 1)]);
        }
        q += dq;
        bf += db;
        if (q >= Qc) {
          q -= Qc;
          ++bf;
        }
      }
      __syncthreads();
      first = false;
      N >>= 3;
      rem -= 3;
    }
    if (rem == 2) {
      const int h = N >> 2, tw = l / N, hmask = h - 1, hshift = 31 - __clz(h);
      const int items = (l >> 2) * Qc;
      int q = q0, bf = b0;
      for (int w = tid; w < items; w += nthr) {
        const int b = bf >> hshift, j = bf & hmask;
        double2* z = Z + ((size_t)(b * N + j)) * Qc + q;
        const size_t hs = (size_t)h * Qc;
        double2 y0, y1, y2, y3;
        dft4(ldz(z, b * N + j, q), ldz(z + hs, b * N + j + h, q), ldz(z + 2 * hs, b * N + j + 2 * h, q),
             ldz(z + 3 * hs, b * N + j + 3 * h, q), y0, y1, y2, y3);
        z[0] = y0;
        if (h == 1) {                     // N == 4: all twiddles are 1
          z[hs] = y1;
          z[2 * hs] = y2;
          z[3 * hs] = y3;
        } else {
          const int jt = j * tw;
          z[hs] = cmul(y1, W[jt & (l - 1)]);
          z[2 * hs] = cmul(y2, W[(2 * jt) & (l - 1)]);
          z[3 * hs] = cmul(y3, W[(3 * jt) & (l - 1)]);
        }
        q += dq;
        bf += db;
        if (q >= Qc) {
          q -= Qc;
          ++bf;
        }
      }
      __syncthreads();
    } else if (rem == 1) {
      // N == 2: last radix-2 stage, twiddle W_2^0 = 1
      const int items = (l >> 1) * Qc;
      int q = q0, b = b0;
      for (int w = tid; w < items; w += nthr) {
        double2* z = Z + ((size_t)(b * 2)) * Qc + q;
        const double2 x0 = ldz(z, b * 2, q), x1 = ldz(z + Qc, b * 2 + 1, q);
        z[0] = cadd(x0, x1);
        z[Qc] = csub(x0, x1);
        q += dq;
        b += db;
        if (q >= Qc) {
          q -= Qc;
          ++b;
        }
      }
      __syncthreads();
    }
    // ---- sampled frequencies (srft_apply!, src/sketch.jl:396-450): rows (i, i+1) <- (Re z, Im z) ----
    double* o = P.out + (int64_t)chunk * P.split_stride + col * P.ldo;
    for (int sp0 = warp * gpw; sp0 < nsamp; sp0 += nwarp * gpw) {
      const int sp = sp0 + sub;
      const bool on = sp < nsamp;
      const int i = 2 * sp;
      double zr = 0.0, zi = 0.0;
      if (on) {
        const int64_t f = sf[sp];
        const int pc = spc[sp], pn = spn[sp];
        const double2 w1 = twiddle(f);                           // w = exp(-2 pi i f / m)
        // w^kk, kk = kk0 + 2q: one table lookup for this lane's first q, then -- when a lane handles several q (short FFTs:
        // Qc > GS) -- a recurrence with w^(2 GS) instead of two table reads + a 64-bit modulo per term (measured at
        // 16384 rows: order 40 0.845 -> 0.744 ms; the same trick on the FFT-pass twiddles made the FP64-bound long passes
        // slower and was dropped)
        const int64_t t0 = (int64_t)(kk0 + 2 * ql) * f, ts = (int64_t)(2 * GS) * f;
        double2 w = twiddle(mpow2 ? (t0 & (P.m - 1)) : (t0 % P.m));
        const double2 wst = (Qc > GS) ? twiddle(mpow2 ? (ts & (P.m - 1)) : (ts % P.m)) : make_double2(1.0, 0.0);
        for (int q = ql; q < Qc; q += GS) {
          const double2 E = Z[(size_t)pc * Qc + q];
          const double2 On = Z[(size_t)pn * Qc + q];             // O = conj(On)
          // even inner column: (E + O)/2 ; odd: -i (E - O)/2
          const double2 xe = make_double2(0.5 * (E.x + On.x), 0.5 * (E.y - On.y));
          const double2 dm = make_double2(0.5 * (E.x - On.x), 0.5 * (E.y + On.y));
          const double2 xo = make_double2(dm.y, -dm.x);
          const double2 term = cmul(w, make_double2(xe.x + fma(w1.x, xo.x, -w1.y * xo.y),
                                                    xe.y + fma(w1.x, xo.y, w1.y * xo.x)));
          zr += term.x;
          zi += term.y;
          w = cmul(w, wst);
        }
      }
      for (int sh = GS >> 1; sh > 0; sh >>= 1) {
        zr += __shfl_xor_sync(0xffffffffu, zr, sh);
        zi += __shfl_xor_sync(0xffffffffu, zi, sh);
      }
      if (on && ql == 0) {
        o[i] = zr;
        if (i + 1 < P.order) o[i + 1] = zi;       // the last row alone gets Re only (src/sketch.jl:425, i == k)
      }
    }
  }
}

// explicit SRFT matrix, K-major: Omt[t + i*ldt], rows (i, i+1) = (d_t cos, -d_t sin)(2 pi t f / m)
__global__ void srft_matrix_kernel(int64_t m, int order, const double* __restrict__ d, const int64_t* __restrict__ idx1,
                                   double* __restrict__ Omt, int64_t ldt) {
  const int nsamp = (order + 1) >> 1;
  const int64_t total = (int64_t)nsamp * m;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int sp = (int)(e / m);
    const int64_t t = e - (int64_t)sp * m;
    const int i = 2 * sp;
    const int64_t f = idx1[i] - 1;
    // (t * f) mod m without overflow for m < 2^31
    const int64_t tf = (t * f) % m;
    double sn, cs;
    sincospi(2.0 * (double)tf / (double)m, &sn, &cs);
    Omt[t + (int64_t)i * ldt] = d[t] * cs;
    if (i + 1 < order) Omt[t + (int64_t)(i + 1) * ldt] = -d[t] * sn;
  }
}

// +-1 signs and uniform indices for the fast mode (device Philox is used for Gaussians; these are O(m) metadata)
struct SplitMix {
  uint64_t s;
  explicit SplitMix(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  uint64_t below(uint64_t n) {       // unbiased enough for sketching: 64-bit multiply-shift
    return (uint64_t)(((unsigned __int128)next() * n) >> 64);
  }
};

__device__ __forceinline__ uint64_t splitmix_at(uint64_t s0, uint64_t i) {      // (i+1)-th output of SplitMix64 seeded s0
  uint64_t z = s0 + (i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// kind 0: +-1 signs.  kind 1: `count` indices in 1..range (random subset).  kind 3: `count` SRFT index entries for an FFT of
// length `range`.
// The reference samples WITH replacement (rand(1:n, k), src/sketch.jl:252, 361).  For the random subset every duplicate
// is a redundant sketch row.  For the real SRFT it is worse: srft_apply! (src/sketch.jl:396-447) turns the entries at the
// ODD (1-based) positions of idx into TWO sketch rows each -- the real and the imaginary part of that frequency -- and
// skips the entry behind it; frequencies f and n - f are complex conjugates, so a sample that holds both (or the same one
// twice) gives four rows of rank two, and frequency 0 or n/2 gives a zero row.  Once k^2 approaches 4n the sketch loses
// rank and the adaptive loop stops short of the requested accuracy (order 264 from n = 1326: rank 252 < 256 ->
// "converged" at a tenth of the true rank's accuracy).  The library's own draws are therefore DISTINCT whenever enough
// distinct values exist (and cover them all before repeating otherwise): kind 1 takes index i to its image under a keyed bijection of [0, range); kind 3 takes the
// j-th USED entry to frequency 1 + (image of j under a keyed bijection of the (range - 1) / 2 conjugate classes,
// DC and Nyquist excluded).  The bijection is a 4-round Feistel network on the smallest even number of bits that covers
// the domain, cycle-walked back into it (< 4 steps expected).  Caller-supplied indices (parity mode) are used as given.
__device__ __forceinline__ uint64_t keyed_bijection(uint64_t x, uint64_t domain, int h, uint64_t s0) {
  const uint64_t mask = (uint64_t(1) << h) - 1;
  do {
    uint64_t L = x >> h, R = x & mask;
#pragma unroll
    for (int rd = 0; rd < 4; ++rd) {
      const uint64_t f = splitmix_at(s0 + 0x632BE59BD9B4E019ull * (uint64_t)(rd + 1), R) & mask;
      const uint64_t t = L ^ f;
      L = R;
      R = t;
    }
    x = (L << h) | R;
  } while (x >= domain);
  return x;
}

__global__ void meta_fill_kernel(int kind, void* __restrict__ dst, int64_t count, int64_t range, uint64_t s0) {
  const int64_t classes = (range - 1) / 2;                       // kind 3: frequencies 1 .. classes
  const int64_t domain = (kind == 3) ? classes : range;
  // kind 3 with more used entries than conjugate classes (a sketch larger than the transform itself): the first
  // `classes` entries cover every class once, then DC, then Nyquist, then the classes again
  const bool distinct = (kind == 1 && count <= range) || (kind == 3 && classes >= 1);
  int h = 1;
  while (h < 31 && (int64_t(1) << (2 * h)) < domain) ++h;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    if (kind == 0) {
      reinterpret_cast<double*>(dst)[i] = (splitmix_at(s0, (uint64_t)i) >> 63) ? 1.0 : -1.0;
    } else if (!distinct || (kind == 3 && (i & 1))) {
      // with replacement (and the entries of an SRFT index vector that srft_apply! never reads)
      reinterpret_cast<int64_t*>(dst)[i] = (int64_t)__umul64hi(splitmix_at(s0, (uint64_t)i), (uint64_t)range) + 1;
    } else if (kind == 1) {
      reinterpret_cast<int64_t*>(dst)[i] = (int64_t)keyed_bijection((uint64_t)i, (uint64_t)domain, h, s0) + 1;
    } else {
      // 1-based index idx = frequency + 1
      int64_t j = i >> 1, f;
      if (j < classes) {
        f = (int64_t)keyed_bijection((uint64_t)j, (uint64_t)domain, h, s0) + 1;
      } else {
        j -= classes;
        const int64_t extra = 1 + ((range & 1) ? 0 : 1);           // DC, and Nyquist for an even length
        const int64_t per = classes + extra;
        const int64_t jj = j % per;
        if (jj == 0) f = 0;
        else if (jj == 1 && extra == 2) f = range / 2;
        else f = (int64_t)keyed_bijection((uint64_t)(jj - extra), (uint64_t)domain, h, s0) + 1;
      }
      reinterpret_cast<int64_t*>(dst)[i] = f + 1;
    }
  }
}

}  // namespace

int bra_sketch_sub(bra_ctx* ctx, char trans, const double* A, int64_t lda, int64_t mA, int64_t nA, int64_t order,
                   const int64_t* r1_dev, double* B, int64_t ldb) {
  if (order <= 0 || nA <= 0) return BRA_OK;
  ProfScope ps(ctx, BRA_PROF_SKETCH_OTHER);
  const int64_t total = order * nA;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->num_sms * 16);
  sketch_sub_kernel<<<blocks, 256, 0, ctx->stream>>>(trans, A, lda, nA, order, r1_dev, B, ldb);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

int bra_sketch_sprn(bra_ctx* ctx, char trans, const double* A, int64_t lda, int64_t mA, int64_t nA, int64_t order,
                    const int64_t* perm1_dev, const double* s_dev, double* B, int64_t ldb) {
  if (order <= 0 || nA <= 0) return BRA_OK;
  ProfScope ps(ctx, BRA_PROF_SKETCH_OTHER);
  const size_t colbytes = (size_t)mA * 8;
  const int stage = colbytes <= (size_t)ctx->smem_optin - 1024 ? 1 : 0;
  const size_t smem = stage ? colbytes : 0;
  if (smem > 48 * 1024)
    BRA_CUDA(cudaFuncSetAttribute(sketch_sprn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks = (int)std::min<int64_t>(nA, (int64_t)ctx->num_sms * 8);
  sketch_sprn_kernel<<<blocks, 256, smem, ctx->stream>>>(trans, A, lda, mA, nA, order, perm1_dev, s_dev, B, ldb, stage);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

// op(A) must already be column-major mA x nA ('c' callers transpose first).
int bra_sketch_srft(bra_ctx* ctx, const double* A, int64_t lda, int64_t mA, int64_t nA, int64_t order,
                    const double* d_dev, const int64_t* idx1_dev, double* B, int64_t ldb) {
  if (order <= 0 || nA <= 0) return BRA_OK;
  // l = largest divisor of m not exceeding order (src/sketch.jl:354-357)
  int64_t l = order;
  while (l > 1 && mA % l > 0) --l;
  if (mA == 0) l = order;
  const int64_t mp = (l > 0) ? mA / l : 0;
  const bool pow2 = l >= 2 && (l & (l - 1)) == 0;
  if (pow2 && mp >= 2 && (mp & 1) == 0 && l <= 4096 && mA < (int64_t(1) << 31)) {
    int logl = 0;
    while ((int64_t(1) << logl) < l) ++logl;
    // chunk of the inner index: even divisor of mp with l*mc*8 <= ~64 KB (several CTAs per SM), at least 2
    const size_t target = (getenv("BRA_SRFT_KB") ? atoi(getenv("BRA_SRFT_KB")) : 64) * 1024;
    const int nthreads = getenv("BRA_SRFT_THREADS") ? atoi(getenv("BRA_SRFT_THREADS")) : 256;
    int64_t mc = mp;
    while (mc > 2 && (size_t)l * mc * 8 > target) {
      // next smaller even divisor
      int64_t c = mc - 2;
      while (c >= 2 && (mp % c != 0 || (c & 1))) c -= 1;
      if (c < 2) break;
      mc = c;
    }
    int64_t tsplit = 1;
    while (tsplit * tsplit < mA) tsplit <<= 1;
    const int64_t nhi = (mA + tsplit - 1) / tsplit;
    const size_t smem = (size_t)l * mc * 8 + (size_t)(l + tsplit + nhi) * 16 + (size_t)3 * ((order + 1) / 2) * 4 +
                        (size_t)((mA + 31) / 32) * 4 + 32;
    if (smem <= (size_t)ctx->smem_optin - 1024) {
      const int chunks = (int)(mp / mc);
      SrftParams P;
      P.A = A;
      P.lda = lda;
      P.m = mA;
      P.n = nA;
      P.order = (int)order;
      P.l = (int)l;
      P.logl = logl;
      P.mp = (int)mp;
      P.mc = (int)mc;
      P.tsplit = (int)tsplit;
      P.nhi = (int)nhi;
      P.d = d_dev;
      P.idx1 = idx1_dev;
      BRA_CUDA(ctx->scratch3.reserve((size_t)((mA + 31) / 32) * 4 + 64));
      P.dbits = ctx->scratch3.as<uint32_t>();
      srft_signbits_kernel<<<std::max<int>(1, (int)std::min<int64_t>(((mA + 31) / 32 + 255) / 256, 1024)), 256, 0, ctx->stream>>>(
          d_dev, mA, ctx->scratch3.as<uint32_t>());
      ctx->launches++;
      // TMA needs 16-byte aligned base and strides (even lda, even m'), box dims <= 256, 32-bit coordinates
      CUtensorMap mapA;
      std::memset(&mapA, 0, sizeof(mapA));
      P.boxrows = (int)std::min<int64_t>(l, 256);
      P.use_tma = 0;
      if ((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (lda & 1) == 0 && (mp & 1) == 0 && mc <= 256 && (mc * 8) % 16 == 0 &&
          nA < (int64_t(1) << 31) && ((size_t)P.boxrows * mc * 8) % 128 == 0)
        P.use_tma = bra_make_map_3d_f64(&mapA, A, (uint64_t)mp, (uint64_t)l, (uint64_t)nA, (uint64_t)mp * 8,
                                        (uint64_t)lda * 8, (uint32_t)mc, (uint32_t)P.boxrows)
                        ? 1
                        : 0;
      if (chunks > 1) {
        BRA_CUDA(ctx->partial.reserve((size_t)chunks * order * nA * 8));
        P.out = ctx->partial.as<double>();
        P.ldo = order;
        P.split_stride = order * nA;
      } else {
        P.out = B;
        P.ldo = ldb;
        P.split_stride = 0;
      }
      BRA_CUDA(cudaFuncSetAttribute(srft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)ctx->smem_optin) / (smem + 1024)));
      int gx = (int)std::min<int64_t>(nA, std::max<int64_t>(1, (int64_t)ctx->num_sms * per_sm / chunks));
      {
        ProfScope ps(ctx, BRA_PROF_SKETCH_OTHER);
        srft_kernel<<<dim3(gx, chunks), nthreads, smem, ctx->stream>>>(mapA, P);
      }
      ctx->launches++;
      BRA_CUDA(cudaGetLastError());
      if (chunks > 1) return bra_splitk_reduce(ctx, P.out, P.split_stride, chunks, order, nA, B, ldb);
      return BRA_OK;
    }
  }
  // general shapes: explicit SRFT matrix (K-major) + the TMA/DMMA GEMM
  const int64_t ldt = (mA + 1) & ~int64_t(1);
  BRA_CUDA(ctx->omega_t.reserve((size_t)order * ldt * 8));
  {
    ProfScope ps(ctx, BRA_PROF_SKETCH_OTHER);
    const int64_t total = ((order + 1) / 2) * mA;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->num_sms * 16);
    if (total > 0) {
      srft_matrix_kernel<<<blocks, 256, 0, ctx->stream>>>(mA, (int)order, d_dev, idx1_dev, ctx->omega_t.as<double>(), ldt);
      ctx->launches++;
      BRA_CUDA(cudaGetLastError());
    }
  }
  return bra_gemm_tn(ctx, ctx->omega_t.as<double>(), ldt, order, mA, A, lda, nA, B, ldb);
}

// Fast-mode random inputs: O(m) metadata from SplitMix64 (the Gaussians stay on the Philox path).
// kind: 0 = signs d (+-1, double, count), 1 = uniform indices in [1, range] (int64, count), 2 = permutation of 1..count.
// Signs and indices are counter-based -- element i is the (i+1)-th output of the SplitMix64 stream of (seed, stream, kind) --
// and generated ON THE DEVICE (no host loop, no copy, no synchronisation); the permutation is a sequential Fisher-Yates
// shuffle and stays on the host.
int bra_fill_meta(bra_ctx* ctx, int kind, void* dst_dev, int64_t count, int64_t range, uint64_t seed, uint64_t stream_id) {
  if (count <= 0) return BRA_OK;
  const uint64_t s0 = seed * 0x9E3779B97F4A7C15ull + stream_id * 0xD1B54A32D192ED03ull + (uint64_t)kind + 1;
  if (kind == 0 || kind == 1 || kind == 3) {
    const int64_t blocks = (count + 255) / 256;
    meta_fill_kernel<<<(unsigned)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, ctx->stream>>>(kind, dst_dev, count, range, s0);
    ctx->launches++;
    BRA_CUDA(cudaGetLastError());
    return BRA_OK;
  }
  SplitMix g(s0);
  ctx->h_meta.resize((size_t)count * 8);
  {
    int64_t* h = reinterpret_cast<int64_t*>(ctx->h_meta.data());
    for (int64_t i = 0; i < count; ++i) h[i] = i + 1;
    for (int64_t i = count - 1; i > 0; --i) std::swap(h[i], h[(int64_t)g.below((uint64_t)i + 1)]);
  }
  BRA_CUDA(cudaMemcpyAsync(dst_dev, ctx->h_meta.data(), (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));       // h_meta is pageable and reused
  return BRA_OK;
}
