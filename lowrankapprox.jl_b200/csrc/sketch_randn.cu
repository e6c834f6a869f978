// Gaussian sketch B = Omega * op(A)   (reference: src/sketch.jl:87-151, mul!(C, ::RandomGaussian, B)
// -> BLAS dgemm at src/sketch.jl:93; crandn at src/util.jl:4).
//
// B200 design: FP64 has no tcgen05/UMMA kind, so "tensor cores" for this
// contraction means the DMMA path (mma.sync.m8n8k4.f64).  The kernel is
// warp-specialised: one producer thread streams K-major tiles of A and of
// Omega^T into a multi-stage shared-memory ring with TMA (cp.async.bulk.tensor
// + mbarrier complete_tx), eight consumer warps run the DMMA loop out of
// conflict-free 64-byte-row tiles.  The sketch is skinny (l = 40..1032 rows), so
// the machine is filled by split-K over the contraction; partial sums are
// reduced by a second kernel in a FIXED order, so results are reproducible
// run to run and independent of scheduling.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
          "r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// ------------------------------------------------------------------ GEMM (TMA + DMMA)
// D[j][i] = sum_k A[k][j] * Omt[k][i]   (MMA rows = columns j of A, MMA cols = sketch rows i)
constexpr int GW = 8;              // consumer warps
constexpr int WB = 4;              // 8-column blocks of A per warp  -> warp covers 32 columns of A
constexpr int TJ = GW * WB * 8;    // 256 columns of A per CTA
constexpr int KC = 8;              // k per chunk (one 64-byte row)
constexpr int CHUNKS = 2;          // chunks per stage
constexpr int KS = KC * CHUNKS;    // k per stage

template <int WA>
struct GemmCfg {
  static constexpr int TI = 8 * WA;                                 // sketch rows per CTA
  static constexpr int STAGE_BYTES = CHUNKS * (TJ + TI) * KC * 8;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int WA>
__global__ void __launch_bounds__((GW + 1) * 32, 1)
gemm_sketch_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapO,
                   int64_t m, int64_t n, int64_t l, int64_t kper, double* __restrict__ out, int64_t ldo,
                   int64_t split_stride, int panel_h) {
  using Cfg = GemmCfg<WA>;
  constexpr int TI = Cfg::TI;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte aligned base for the TMA destinations
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t j0 = (int64_t)blockIdx.x * TJ;
  const int64_t i0 = (int64_t)blockIdx.y * TI;
  const int64_t kbeg = (int64_t)blockIdx.z * kper;
  const int64_t kend = min(m, kbeg + kper);
  const int nst = (int)((kend - kbeg + KS - 1) / KS);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], GW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == GW) {
    // ===== TMA producer (one elected lane) =====
    if (lane == 0) {
      for (int it = 0; it < nst; ++it) {
        const int s = it % STAGES;
        if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) - 1) & 1);
        unsigned char* st = base + s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
        const int64_t k0 = kbeg + (int64_t)it * KS;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          const int64_t ka = k0 + c * KC;
          if (panel_h) {        // A repacked into row panels (tall matrices): 3-D map (row in panel, column, panel)
            const int64_t pn = ka / panel_h;
            tma_load_3d(st + c * (TJ + TI) * KC * 8, &mapA, (int)(ka - pn * panel_h), (int)j0, (int)pn, &full[s]);
          } else {
            tma_load_2d(st + c * (TJ + TI) * KC * 8, &mapA, (int)ka, (int)j0, &full[s]);
          }
          tma_load_2d(st + c * (TJ + TI) * KC * 8 + TJ * KC * 8, &mapO, (int)(k0 + c * KC), (int)i0, &full[s]);
        }
      }
    }
    return;
  }

  // ===== DMMA consumers =====
  double acc[WB][WA][2];
#pragma unroll
  for (int b = 0; b < WB; ++b)
#pragma unroll
    for (int a = 0; a < WA; ++a) acc[b][a][0] = acc[b][a][1] = 0.0;

  const int row = lane >> 2, q = lane & 3;
  for (int it = 0; it < nst; ++it) {
    const int s = it % STAGES;
    mbar_wait(&full[s], (it / STAGES) & 1);
    const unsigned char* st = base + s * Cfg::STAGE_BYTES;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      const double2* As = reinterpret_cast<const double2*>(st + c * (TJ + TI) * KC * 8);
      const double2* Os = reinterpret_cast<const double2*>(st + c * (TJ + TI) * KC * 8 + TJ * KC * 8);
      double2 af[WB], of[WA];
#pragma unroll
      for (int b = 0; b < WB; ++b) af[b] = As[(warp * (WB * 8) + b * 8 + row) * (KC / 2) + q];
#pragma unroll
      for (int a = 0; a < WA; ++a) of[a] = Os[(a * 8 + row) * (KC / 2) + q];
#pragma unroll
      for (int b = 0; b < WB; ++b)
#pragma unroll
        for (int a = 0; a < WA; ++a) {
          dmma884(acc[b][a][0], acc[b][a][1], af[b].x, of[a].x);
          dmma884(acc[b][a][0], acc[b][a][1], af[b].y, of[a].y);
        }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // ===== epilogue: D[j][i] -> out[i + j*ldo] (+ split offset) =====
  double* o = out + (int64_t)blockIdx.z * split_stride;
  const bool vec = ((ldo & 1) == 0) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
#pragma unroll
  for (int b = 0; b < WB; ++b) {
    const int64_t j = j0 + warp * (WB * 8) + b * 8 + row;
    if (j >= n) continue;
#pragma unroll
    for (int a = 0; a < WA; ++a) {
      const int64_t i = i0 + a * 8 + 2 * q;
      if (i + 1 < l && vec) {
        *reinterpret_cast<double2*>(o + i + j * ldo) = make_double2(acc[b][a][0], acc[b][a][1]);
      } else {
        if (i < l) o[i + j * ldo] = acc[b][a][0];
        if (i + 1 < l) o[i + 1 + j * ldo] = acc[b][a][1];
      }
    }
  }
}

// deterministic split-K reduction: out[e] = sum_{s=0..S-1} part[s][e], fixed order
__global__ void splitk_reduce_kernel(const double* __restrict__ part, int64_t split_stride, int splits,
                                     int64_t l, int64_t n, double* __restrict__ out, int64_t ldo) {
  const int64_t total = l * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int s = 0; s < splits; ++s) acc += part[(int64_t)s * split_stride + e];
    const int64_t i = e % l, j = e / l;
    out[i + j * ldo] = acc;
  }
}

// ------------------------------------------------------------------ generic fallback (any alignment / trans)
// C[i][j] = sum_k Om[i*osi + k*osk] * A[k*sk + j*sj];  T x T tile (T = 64: 4x4 per thread, T = 32: 2x2 per thread,
// four times the CTAs -- the k x k x k products of the tails are latency-critical and too small to fill the machine
// with 64 x 64 tiles), DFMA.
template <int T>
__global__ void __launch_bounds__(256) gemm_generic_kernel(const double* __restrict__ Om, int64_t osi, int64_t osk,
                                                           const double* __restrict__ A, int64_t sk, int64_t sj,
                                                           int64_t l, int64_t n, int64_t K, double* __restrict__ C,
                                                           int64_t ldc) {
  constexpr int MT = T / 16;          // micro-tile edge
  __shared__ double Os[16][T + 1];
  __shared__ double As[16][T + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * T, j0 = (int64_t)blockIdx.x * T;
  double acc[MT][MT] = {};
  for (int64_t k0 = 0; k0 < K; k0 += 16) {
    for (int e = threadIdx.x; e < 16 * T; e += 256) {
      // let the unit-stride index vary fastest across threads
      int kk = (osk == 1) ? e % 16 : e / T, ii = (osk == 1) ? e / 16 : e % T;
      int64_t k = k0 + kk;
      Os[kk][ii] = (k < K && i0 + ii < l) ? Om[(i0 + ii) * osi + k * osk] : 0.0;
      int kk2 = (sk == 1) ? e % 16 : e / T, jj2 = (sk == 1) ? e / 16 : e % T;
      int64_t k2 = k0 + kk2;
      As[kk2][jj2] = (k2 < K && j0 + jj2 < n) ? A[k2 * sk + (j0 + jj2) * sj] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double o[MT], a[MT];
#pragma unroll
      for (int u = 0; u < MT; ++u) {
        o[u] = Os[kk][tx + 16 * u];
        a[u] = As[kk][ty + 16 * u];
      }
#pragma unroll
      for (int u = 0; u < MT; ++u)
#pragma unroll
        for (int v = 0; v < MT; ++v) acc[u][v] = fma(o[u], a[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < MT; ++u)
#pragma unroll
    for (int v = 0; v < MT; ++v) {
      int64_t i = i0 + tx + 16 * u, j = j0 + ty + 16 * v;
      if (i < l && j < n) C[i + j * ldc] = acc[u][v];
    }
}

// ------------------------------------------------------------------ Omega handling
// Omt[k + i*ldt] = Om[i + k*ldo]   (K-major copy of a caller-supplied Omega)
__global__ void transpose_kernel(const double* __restrict__ src, int64_t lds, int64_t rows, int64_t cols,
                                 double* __restrict__ dst, int64_t ldd) {
  __shared__ double t[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    int64_t r = r0 + threadIdx.x, c = c0 + y;
    t[y][threadIdx.x] = (r < rows && c < cols) ? src[r + c * lds] : 0.0;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    int64_t c = c0 + threadIdx.x, r = r0 + y;
    if (r < rows && c < cols) dst[c + r * ldd] = t[threadIdx.x][y];
  }
}

// Philox4x32-10 counter-based generator + Box-Muller: element e of stream (seed, stream_id)
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0;
  c[1] = n1;
  c[2] = n2;
  c[3] = n3;
}
__global__ void fill_randn_kernel(double* __restrict__ dst, int64_t count, uint64_t seed, uint64_t stream_id) {
  const int64_t pairs = (count + 1) / 2;
  for (int64_t pidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pidx < pairs;
       pidx += (int64_t)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)pidx, (uint32_t)((uint64_t)pidx >> 32), (uint32_t)stream_id,
                     (uint32_t)(stream_id >> 32)};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    // two 53-bit uniforms in (0,1]
    const double u1 = ((double)(((uint64_t)c[0] << 21) ^ (c[1] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)(((uint64_t)c[2] << 21) ^ (c[3] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    dst[2 * pidx] = rad * cs;
    if (2 * pidx + 1 < count) dst[2 * pidx + 1] = rad * sn;
  }
}

// Same stream, addressed by GLOBAL position: local element (i, k) of the K-major Omega^T block of a row shard is
// element e = i*ldg + row0 + k of the (seed, stream_id) sequence, ldg = roundup(m_global, 2).  With row0 = 0 and
// ldg = ldt this reproduces fill_randn_kernel exactly.
__device__ __forceinline__ void philox_normal_pair(uint64_t pidx, uint64_t seed, uint64_t stream_id, double& z0, double& z1) {
  uint32_t c[4] = {(uint32_t)pidx, (uint32_t)(pidx >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const double u1 = ((double)(((uint64_t)c[0] << 21) ^ (c[1] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);
  const double u2 = ((double)(((uint64_t)c[2] << 21) ^ (c[3] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
  const double rad = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  z0 = rad * cs;
  z1 = rad * sn;
}
__global__ void fill_randn_rows_kernel(double* __restrict__ dst, int64_t ldt, int64_t order, int64_t row0, int64_t ldg,
                                       uint64_t seed, uint64_t stream_id) {
  const int64_t hp = ldt / 2, pairs = order * hp;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < pairs; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / hp, k = 2 * (t - i * hp);
    const uint64_t e0 = (uint64_t)(i * ldg + row0 + k);
    double a, b, z0, z1;
    philox_normal_pair(e0 >> 1, seed, stream_id, z0, z1);
    if ((e0 & 1) == 0) {
      a = z0;
      b = z1;
    } else {
      a = z1;
      philox_normal_pair((e0 + 1) >> 1, seed, stream_id, z0, z1);
      b = z0;
    }
    dst[k + i * ldt] = a;
    dst[k + 1 + i * ldt] = b;
  }
}

// ------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D FP64 tensor: dim0 = contraction index (contiguous), dim1 = rows/cols, box (KC, box1)
bool make_map(CUtensorMap* map, const double* base, int64_t d0, int64_t d1, int64_t ld, int box1) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace

// 3-D FP64 tensor map (dim0 contiguous) for the SRFT row staging: (d0, d1, d2) with byte strides (s1, s2), box (b0, b1, 1)
bool bra_make_map_3d_f64(CUtensorMap* map, const double* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1,
                         uint64_t s2, uint32_t b0, uint32_t b1) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1, s2};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

namespace {

template <int WA>
int launch_gemm(bra_ctx* ctx, const CUtensorMap& mapA, const double* Omt, int64_t ldt, int64_t l, int64_t m,
                int64_t n, double* B, int64_t ldb, int panel_h = 0) {
  using Cfg = GemmCfg<WA>;
  CUtensorMap mapO;
  if (!make_map(&mapO, Omt, m, l, ldt, Cfg::TI)) {
    ctx->set_error("cuTensorMapEncodeTiled(Omega^T) failed");
    return BRA_ERR_CUDA;
  }
  const int jt = (int)((n + TJ - 1) / TJ);
  const int itl = (int)((l + Cfg::TI - 1) / Cfg::TI);
  // split-K: every CTA does the same work, so a partial last wave is pure loss (352 tiles on 148 SMs run as 3
  // waves at 79 %).  Pick the split count that minimises   waves(sp) * t_tile / sp  +  sp * t_reduce   with
  // t_tile = one full-K tile at ~90 % of an SM's DMMA rate and t_reduce = writing and re-reading one set of partial
  // sums; each split keeps >= 8 stages and the partial buffer stays under 512 MB.
  int64_t maxs = (m + 8 * KS - 1) / (8 * KS);
  if (maxs > 32) maxs = 32;
  const int64_t memcap = (int64_t(512) << 20) / (l * n * 8 > 0 ? l * n * 8 : 1);
  if (maxs > memcap) maxs = memcap;
  if (maxs < 1) maxs = 1;
  int splits = 1;
  {
    const int tiles = jt * itl, sms = ctx->num_sms;
    const double t_tile = 2.0 * Cfg::TI * TJ * (double)m / 225e9;
    const double t_red = 16.0 * (double)l * (double)n / 4e12 + 3e-6;
    double best = 1e30;
    for (int sp = 1; sp <= (int)maxs; ++sp) {
      const int64_t total = (int64_t)tiles * sp;
      const int64_t waves = (total + sms - 1) / sms;
      const double t = (double)waves * t_tile / sp + (sp > 1 ? sp * t_red : 0.0);
      if (t < best * 0.995) {
        best = t;
        splits = sp;
      }
    }
  }
  int64_t kper = (m + splits - 1) / splits;
  kper = (kper + KS - 1) / KS * KS;
  splits = (int)((m + kper - 1) / kper);
  double* out = B;
  int64_t ldo = ldb, sstride = 0;
  if (splits > 1) {
    BRA_CUDA(ctx->ws_partial().reserve((size_t)splits * l * n * 8));
    out = ctx->ws_partial().as<double>();
    ldo = l;
    sstride = l * n;
  }
  BRA_CUDA(cudaFuncSetAttribute(gemm_sketch_kernel<WA>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  dim3 grid(jt, itl, splits);
  {
    ProfScope ps(ctx, ctx->gemm_tag);
    gemm_sketch_kernel<WA><<<grid, (GW + 1) * 32, Cfg::SMEM, ctx->stream>>>(mapA, mapO, m, n, l, kper, out, ldo, sstride,
                                                                            panel_h);
  }
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  if (splits > 1) {
    ProfScope ps(ctx, BRA_PROF_SPLITK);
    int64_t total = l * n;
    int blocks = (int)((total + 255) / 256 < (int64_t)ctx->num_sms * 16 ? (total + 255) / 256 : (int64_t)ctx->num_sms * 16);
    splitk_reduce_kernel<<<blocks, 256, 0, ctx->stream>>>(out, sstride, splits, l, n, B, ldb);
    ctx->launches++;
    BRA_CUDA(cudaGetLastError());
  }
  return BRA_OK;
}

}  // namespace

int bra_splitk_reduce(bra_ctx* ctx, const double* part, int64_t split_stride, int splits, int64_t l, int64_t n,
                      double* out, int64_t ldo) {
  ProfScope ps(ctx, BRA_PROF_SPLITK);
  const int64_t total = l * n;
  if (total <= 0) return BRA_OK;
  int blocks = (int)((total + 255) / 256 < (int64_t)ctx->num_sms * 16 ? (total + 255) / 256 : (int64_t)ctx->num_sms * 16);
  splitk_reduce_kernel<<<blocks, 256, 0, ctx->stream>>>(part, split_stride, splits, l, n, out, ldo);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

int bra_transpose_omega(bra_ctx* ctx, const double* Om, int64_t ldo, int64_t l, int64_t m, double* Omt) {
  // Om is l x m (col-major); Omt[k + i*ldt], ldt = m rounded up to even
  const int64_t ldt = (m + 1) & ~int64_t(1);
  dim3 grid((unsigned)((l + 31) / 32), (unsigned)((m + 31) / 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(Om, ldo, l, m, Omt, ldt);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

int bra_fill_randn(bra_ctx* ctx, double* dst, int64_t count, uint64_t seed, uint64_t stream_id) {
  if (count <= 0) return BRA_OK;
  fill_randn_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(dst, count, seed, stream_id);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

int bra_fill_randn_rows(bra_ctx* ctx, double* dst, int64_t ldt, int64_t order, int64_t row0, int64_t ldg, uint64_t seed,
                        uint64_t stream_id) {
  if (order <= 0 || ldt <= 0) return BRA_OK;
  fill_randn_rows_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(dst, ldt, order, row0, ldg, seed, stream_id);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

// Can the TMA path take this A?  (16-byte aligned base, even lda, 32-bit coordinates)
bool bra_gemm_tma_ok(const double* A, int64_t lda, int64_t m, int64_t n) {
  return ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((lda & 1) == 0) && m >= 1 && m < (int64_t(1) << 31) &&
         n < (int64_t(1) << 31) && get_encode() != nullptr;
}

// B (l x n) = Omega (l x m) * A (m x n); Omt is the K-major copy of Omega, ldt = roundup(m, 2)
int bra_gemm_sketch(bra_ctx* ctx, const double* Omt, int64_t l, int64_t m, const double* A, int64_t lda, int64_t n,
                    double* B, int64_t ldb) {
  return bra_gemm_tn(ctx, Omt, (m + 1) & ~int64_t(1), l, m, A, lda, n, B, ldb);
}

// A tall A (column stride >= 1 MB) puts every column of a 256-column tile in its own 2 MB page; the TMA stream then
// thrashes the TLB (measured: 33 TF at 64 K rows, 24 TF at 256 K rows, 18.7 TF at 1 M rows).  Such an A is repacked
// ONCE per factorization into row panels of PANEL_H rows, P[panel][column][row] (column stride 128 KB), and the
// kernel walks it through a 3-D tensor map.  Costs one read + one write of A (a few % of one round's GEMM).
constexpr int PANEL_H = 16384;

__global__ void repack_panels_kernel(const double* __restrict__ A, int64_t lda, int64_t m, int64_t n,
                                     double* __restrict__ P) {
  // one CTA per (column, panel) segment: PANEL_H contiguous doubles in, PANEL_H contiguous doubles out
  const int64_t npan = (m + PANEL_H - 1) / PANEL_H;
  for (int64_t seg = blockIdx.x; seg < n * npan; seg += gridDim.x) {
    const int64_t j = seg % n, p = seg / n;
    const double2* src = reinterpret_cast<const double2*>(A + p * PANEL_H + j * lda);
    double2* dst = reinterpret_cast<double2*>(P + (p * n + j) * (int64_t)PANEL_H);
    const int64_t rows = min((int64_t)PANEL_H, m - p * PANEL_H);
    for (int64_t r = threadIdx.x; r < PANEL_H / 2; r += blockDim.x) {
      double2 v = make_double2(0.0, 0.0);
      if (2 * r + 1 < rows) v = src[r];
      else if (2 * r < rows) v.x = A[p * PANEL_H + j * lda + 2 * r];
      dst[r] = v;
    }
  }
}

bool bra_gemm_wants_panels(const double* A, int64_t lda, int64_t m, int64_t n) {
  return bra_gemm_tma_ok(A, lda, m, n) && lda * 8 >= (int64_t(1) << 20) && m >= 4 * (int64_t)PANEL_H;
}

int bra_repack_panels(bra_ctx* ctx, const double* A, int64_t lda, int64_t m, int64_t n, double* P) {
  const int64_t npan = (m + PANEL_H - 1) / PANEL_H;
  const int64_t segs = n * npan;
  repack_panels_kernel<<<(unsigned)(segs < 148 * 16 ? segs : 148 * 16), 256, 0, ctx->stream>>>(A, lda, m, n, P);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

int64_t bra_panel_bytes(int64_t m, int64_t n) { return ((m + PANEL_H - 1) / PANEL_H) * n * (int64_t)PANEL_H * 8; }

static int pick_wa(int64_t l) {
  // the row-tile height that wastes the fewest padded sketch rows (l = 8a: 40, 72, 136, 264, 520, ...)
  int best = 5;
  int64_t bestpad = -1;
  for (int wa : {5, 4, 3}) {
    int64_t ti = 8 * wa, pad = (l + ti - 1) / ti * ti;
    if (bestpad < 0 || pad < bestpad) {
      bestpad = pad;
      best = wa;
    }
  }
  return best;
}

// B (l x n) = Omega * A with A given as row panels (bra_repack_panels)
int bra_gemm_sketch_panels(bra_ctx* ctx, const double* Omt, int64_t l, int64_t m, const double* P, int64_t n, double* B,
                           int64_t ldb) {
  if (l <= 0 || n <= 0) return BRA_OK;
  const int64_t ldt = (m + 1) & ~int64_t(1);
  const int64_t npan = (m + PANEL_H - 1) / PANEL_H;
  CUtensorMap mapA;
  if (!bra_make_map_3d_f64(&mapA, P, PANEL_H, (uint64_t)n, (uint64_t)npan, (uint64_t)PANEL_H * 8,
                           (uint64_t)PANEL_H * 8 * (uint64_t)n, KC, TJ)) {
    ctx->set_error("cuTensorMapEncodeTiled(A panels) failed");
    return BRA_ERR_CUDA;
  }
  const int best = pick_wa(l);
  if (best == 5) return launch_gemm<5>(ctx, mapA, Omt, ldt, l, m, n, B, ldb, PANEL_H);
  if (best == 4) return launch_gemm<4>(ctx, mapA, Omt, ldt, l, m, n, B, ldb, PANEL_H);
  return launch_gemm<3>(ctx, mapA, Omt, ldt, l, m, n, B, ldb, PANEL_H);
}

// General "TN" product on the same TMA + DMMA kernel: C (l x n) = X^T Y with X (m x l, ldx) and Y (m x n, ldy)
// both column-major, i.e. both contiguous along the contraction index.
int bra_gemm_tn(bra_ctx* ctx, const double* Omt, int64_t ldt, int64_t l, int64_t m, const double* A, int64_t lda,
                int64_t n, double* B, int64_t ldb) {
  if (l <= 0 || n <= 0) return BRA_OK;
  if (!bra_gemm_tma_ok(A, lda, m, n) || !bra_gemm_tma_ok(Omt, ldt, m, l))
    return bra_gemm_generic(ctx, Omt, ldt, 1, A, 1, lda, l, n, m, B, ldb);
  CUtensorMap mapA;
  if (!make_map(&mapA, A, m, n, lda, TJ)) {
    ctx->set_error("cuTensorMapEncodeTiled(A) failed");
    return BRA_ERR_CUDA;
  }
  const int best = pick_wa(l);
  if (best == 5) return launch_gemm<5>(ctx, mapA, Omt, ldt, l, m, n, B, ldb);
  if (best == 4) return launch_gemm<4>(ctx, mapA, Omt, ldt, l, m, n, B, ldb);
  return launch_gemm<3>(ctx, mapA, Omt, ldt, l, m, n, B, ldb);
}

// C (l x n) = Om * Aop with arbitrary strides: Om(i,k) = Om[i*osi + k*osk], Aop(k,j) = A[k*sk + j*sj].
// Used for unaligned operands and for the (:left,:c) form (contraction along A's rows' stride).
int bra_gemm_generic(bra_ctx* ctx, const double* Om, int64_t osi, int64_t osk, const double* A, int64_t sk,
                     int64_t sj, int64_t l, int64_t n, int64_t K, double* C, int64_t ldc) {
  if (l <= 0 || n <= 0) return BRA_OK;
  ProfScope ps(ctx, ctx->gemm_tag);
  if (((n + 63) / 64) * ((l + 63) / 64) < 2 * (int64_t)ctx->num_sms) {
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((l + 31) / 32));
    gemm_generic_kernel<32><<<grid, 256, 0, ctx->stream>>>(Om, osi, osk, A, sk, sj, l, n, K, C, ldc);
  } else {
    dim3 grid((unsigned)((n + 63) / 64), (unsigned)((l + 63) / 64));
    gemm_generic_kernel<64><<<grid, 256, 0, ctx->stream>>>(Om, osi, osk, A, sk, sj, l, n, K, C, ldc);
  }
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}
