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
//    in a fixed order.  Twiddles come from sincospi (the reference multiplies them up, drifting by O(m' eps)).
//    Shapes the radix passes do not cover (l not a power of two, odd m') go through an explicitly generated
//    SRFT matrix and the TMA + DMMA GEMM.
//  * sparse Gaussian: every entry of A is read exactly once; a column is staged in shared memory (coalesced)
//    and the order-many weighted sums are formed from there in the reference's summation order.
//  * random subset: a gather.
#include "common.cuh"
#include <algorithm>

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
  const double* d;        // +-1, length m
  const int64_t* idx1;    // 1-based sampled frequencies, length order
  double* out;            // order x n (ld = ldo) [+ chunk * split_stride]
  int64_t ldo, split_stride;
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}

// position of frequency bin c after the in-place DIF passes (radix 4 while >= 2 stages remain, then radix 2)
__device__ __forceinline__ int srft_binpos(int c, int l, int logl) {
  int pos = 0, N = l, rem = logl;
  while (rem >= 2) {
    pos += (c & 3) * (N >> 2);
    c >>= 2;
    N >>= 2;
    rem -= 2;
  }
  if (rem == 1) pos += (c & 1) * (N >> 1);
  return pos;
}

__global__ void __launch_bounds__(256) srft_kernel(SrftParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
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

  for (int64_t col = blockIdx.x; col < P.n; col += gridDim.x) {
    const double* x = P.A + col * P.lda;
    __syncthreads();
    // ---- srft_reshape!: X[j, kk] = d[i] x[i], i = j*mp + kk, restricted to this chunk of kk ----
    for (int e = tid; e < l * mc; e += nthr) {
      const int j = e / mc, t = e - j * mc;
      const int64_t i = (int64_t)j * P.mp + kk0 + t;
      Xs[e] = P.d[i] * x[i];
    }
    __syncthreads();
    // ---- length-l DFT along j of the Qc complex columns, in place, decimation in frequency ----
    int N = l, rem = P.logl;
    while (rem >= 2) {
      const int h = N >> 2, tw = l / N;
      const int items = (l >> 2) * Qc;
      for (int w = tid; w < items; w += nthr) {
        const int q = w % Qc, bf = w / Qc;
        const int b = bf / h, j = bf - b * h;
        double2* z = Z + ((size_t)(b * N + j)) * Qc + q;
        const size_t hs = (size_t)h * Qc;
        const double2 x0 = z[0], x1 = z[hs], x2 = z[2 * hs], x3 = z[3 * hs];
        const double2 s02 = make_double2(x0.x + x2.x, x0.y + x2.y), d02 = make_double2(x0.x - x2.x, x0.y - x2.y);
        const double2 s13 = make_double2(x1.x + x3.x, x1.y + x3.y), d13 = make_double2(x1.x - x3.x, x1.y - x3.y);
        // -i * d13 = (d13.y, -d13.x)
        const double2 y0 = make_double2(s02.x + s13.x, s02.y + s13.y);
        const double2 y2 = make_double2(s02.x - s13.x, s02.y - s13.y);
        const double2 y1 = make_double2(d02.x + d13.y, d02.y - d13.x);
        const double2 y3 = make_double2(d02.x - d13.y, d02.y + d13.x);
        z[0] = y0;
        z[hs] = cmul(y1, W[(j * tw) & (l - 1)]);
        z[2 * hs] = cmul(y2, W[(2 * j * tw) & (l - 1)]);
        z[3 * hs] = cmul(y3, W[(3 * j * tw) & (l - 1)]);
      }
      __syncthreads();
      N >>= 2;
      rem -= 2;
    }
    if (rem == 1) {
      // N == 2: last radix-2 stage, twiddle W_2^0 = 1
      const int items = (l >> 1) * Qc;
      for (int w = tid; w < items; w += nthr) {
        const int q = w % Qc, b = w / Qc;
        double2* z = Z + ((size_t)(b * 2)) * Qc + q;
        const double2 x0 = z[0], x1 = z[Qc];
        z[0] = make_double2(x0.x + x1.x, x0.y + x1.y);
        z[Qc] = make_double2(x0.x - x1.x, x0.y - x1.y);
      }
      __syncthreads();
    }
    // ---- sampled frequencies (srft_apply!, src/sketch.jl:396-450): rows (i, i+1) <- (Re z, Im z) ----
    double* o = P.out + (int64_t)chunk * P.split_stride + col * P.ldo;
    for (int sp = warp; sp < nsamp; sp += nwarp) {
      const int i = 2 * sp;
      const int64_t f = P.idx1[i] - 1;
      const int c = (int)(f % l);
      const int pc = srft_binpos(c, l, P.logl), pn = srft_binpos((l - c) & (l - 1), l, P.logl);
      double zr = 0.0, zi = 0.0;
      if (lane < Qc) {
        // w^kk = exp(-2 pi i kk f / m); per-lane start at kk = kk0 + 2 lane, stride 64 in kk
        double sn, cs;
        int64_t t0 = ((int64_t)(kk0 + 2 * lane) * f) % P.m;
        sincospi(-2.0 * (double)t0 / (double)P.m, &sn, &cs);
        double2 w = make_double2(cs, sn);
        int64_t t1 = f % P.m;
        sincospi(-2.0 * (double)t1 / (double)P.m, &sn, &cs);
        const double2 w1 = make_double2(cs, sn);                 // w^1
        int64_t t64 = (64 * f) % P.m;
        sincospi(-2.0 * (double)t64 / (double)P.m, &sn, &cs);
        const double2 w64 = make_double2(cs, sn);                // w^64
        for (int q = lane; q < Qc; q += 32) {
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
          w = cmul(w, w64);
        }
      }
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        zr += __shfl_xor_sync(0xffffffffu, zr, s);
        zi += __shfl_xor_sync(0xffffffffu, zi, s);
      }
      if (lane == 0) {
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
    const size_t target = 64 * 1024;
    int64_t mc = mp;
    while (mc > 2 && (size_t)l * mc * 8 > target) {
      // next smaller even divisor
      int64_t c = mc - 2;
      while (c >= 2 && (mp % c != 0 || (c & 1))) c -= 1;
      if (c < 2) break;
      mc = c;
    }
    const size_t smem = (size_t)l * mc * 8 + (size_t)l * 16;
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
      P.d = d_dev;
      P.idx1 = idx1_dev;
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
        srft_kernel<<<dim3(gx, chunks), 256, smem, ctx->stream>>>(P);
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

// Fast-mode random inputs (host SplitMix64 -> device; O(m) metadata, Gaussians stay on the device Philox path).
// kind: 0 = signs d (+-1, double, count), 1 = uniform indices in [1, range] (int64, count), 2 = permutation of 1..count
int bra_fill_meta(bra_ctx* ctx, int kind, void* dst_dev, int64_t count, int64_t range, uint64_t seed, uint64_t stream_id) {
  if (count <= 0) return BRA_OK;
  SplitMix g(seed * 0x9E3779B97F4A7C15ull + stream_id * 0xD1B54A32D192ED03ull + (uint64_t)kind + 1);
  ctx->h_meta.resize((size_t)count * 8);
  if (kind == 0) {
    double* h = reinterpret_cast<double*>(ctx->h_meta.data());
    for (int64_t i = 0; i < count; ++i) h[i] = (g.next() >> 63) ? 1.0 : -1.0;
  } else if (kind == 1) {
    int64_t* h = reinterpret_cast<int64_t*>(ctx->h_meta.data());
    for (int64_t i = 0; i < count; ++i) h[i] = (int64_t)g.below((uint64_t)range) + 1;
  } else {
    int64_t* h = reinterpret_cast<int64_t*>(ctx->h_meta.data());
    for (int64_t i = 0; i < count; ++i) h[i] = i + 1;
    for (int64_t i = count - 1; i > 0; --i) std::swap(h[i], h[(int64_t)g.below((uint64_t)i + 1)]);
  }
  BRA_CUDA(cudaMemcpyAsync(dst_dev, ctx->h_meta.data(), (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));       // h_meta is pageable and reused
  return BRA_OK;
}
