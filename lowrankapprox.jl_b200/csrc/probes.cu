// Hardware probes used as roofline denominators by bench.py (MEASURED_PEAKS.json has no FP64 figure):
// register-resident DMMA / DFMA issue loops, and the latency of the LL all-gather exchange the QRCP
// kernel performs once per pivot step.
#include "common.cuh"

namespace {

__device__ __forceinline__ void p_dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void p_dmma1684(double (&d)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void p_dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void p_dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
      "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
        "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// mode 0: m8n8k4, 1: DFMA, 2: m16n8k4, 3: m16n8k8, 4: m16n8k16.  8 independent accumulator chains per warp.
template <int MODE>
__global__ void __launch_bounds__(256) peak_kernel(int iters, double seed, double* sink) {
  const double a0 = seed + threadIdx.x * 1e-9, b0 = 1.0 - seed;
  double acc[8][4];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[c][e] = 0.0;
  double a8[8], b4[4];
#pragma unroll
  for (int e = 0; e < 8; ++e) a8[e] = a0 + e * 1e-3;
#pragma unroll
  for (int e = 0; e < 4; ++e) b4[e] = b0 + e * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (MODE == 0) {
        p_dmma884(acc[c][0], acc[c][1], a8[0], b4[0]);
        p_dmma884(acc[c][2], acc[c][3], a8[1], b4[1]);
      } else if (MODE == 1) {
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[c][e] = fma(a8[e], b4[e], acc[c][e]);
      } else if (MODE == 2) {
        double a2[2] = {a8[0], a8[1]};
        p_dmma1684(acc[c], a2, b4[0]);
      } else if (MODE == 3) {
        double a4[4] = {a8[0], a8[1], a8[2], a8[3]};
        double b2[2] = {b4[0], b4[1]};
        p_dmma1688(acc[c], a4, b2);
      } else {
        p_dmma16816(acc[c], a8, b4);
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int e = 0; e < 4; ++e) s += acc[c][e];
  if (s == 123.456) sink[0] = s;
}

template <int MODE>
int run_peak(bra_ctx* ctx, double flops_per_warp_iter, double* tf) {
  const int iters = 4096;
  const int blocks = ctx->num_sms * 4;
  cudaEvent_t e0, e1;
  BRA_CUDA(cudaEventCreate(&e0));
  BRA_CUDA(cudaEventCreate(&e1));
  BRA_CUDA(ctx->scratch3.reserve(64));
  peak_kernel<MODE><<<blocks, 256, 0, ctx->stream>>>(64, 0.25, ctx->scratch3.as<double>());
  BRA_CUDA(cudaEventRecord(e0, ctx->stream));
  peak_kernel<MODE><<<blocks, 256, 0, ctx->stream>>>(iters, 0.25, ctx->scratch3.as<double>());
  BRA_CUDA(cudaEventRecord(e1, ctx->stream));
  BRA_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  BRA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double flops = flops_per_warp_iter * 8.0 /*warps*/ * blocks * iters;
  *tf = flops / (ms * 1e-3) / 1e12;
  ctx->launches += 2;
  return BRA_OK;
}

// ---- exchange latency: the same LL all-gather as qrcp.cu, nothing else ----
struct __align__(16) PLL16 {
  uint32_t lo, s0, hi, s1;
};
__global__ void __launch_bounds__(512, 1) exchange_kernel(PLL16* rec, int iters, uint32_t epoch, int* fail) {
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
  __shared__ int s_bad;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  for (int s = 0; s < iters; ++s) {
    const uint32_t stamp = epoch + s;
    PLL16* mine = rec + ((size_t)(s & 1) * G + cta) * 4;
    if (tid == 0)
      asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(mine), "r"((uint32_t)s), "r"(stamp),
                   "r"((uint32_t)cta), "r"(stamp)
                   : "memory");
    if (tid < G) {
      const PLL16* r = rec + ((size_t)(s & 1) * G + tid) * 4;
      uint32_t lo, s0, hi, s1, spins = 0;
      do {
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(lo), "=r"(s0), "=r"(hi), "=r"(s1)
                     : "l"(r)
                     : "memory");
      } while ((s0 != stamp || s1 != stamp) && ++spins < (1u << 22));
      if (s0 != stamp || s1 != stamp) s_bad = 1;
    }
    __syncthreads();
    if (s_bad) break;
  }
  if (tid == 0 && s_bad) *fail = 1;
}

__device__ __forceinline__ void pst(PLL16* p, uint32_t a, uint32_t b, uint32_t stamp) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(stamp), "r"(b), "r"(stamp)
               : "memory");
}
__device__ __forceinline__ bool pld(const PLL16* p, uint32_t stamp) {
  uint32_t lo, s0, hi, s1, spins = 0;
  do {
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(lo), "=r"(s0), "=r"(hi), "=r"(s1)
                 : "l"(p)
                 : "memory");
  } while ((s0 != stamp || s1 != stamp) && ++spins < (1u << 22));
  return s0 == stamp && s1 == stamp;
}

// mode 0: push `hw` contiguous words into every CTA's inbox, each CTA polls its own inbox.
// mode 1: two hops through `leaders` leader CTAs (slot -> group result -> everyone).
__global__ void __launch_bounds__(512, 1) exchange2_kernel(PLL16* buf, int iters, uint32_t epoch, int hw, int mode,
                                                           int leaders, int* fail) {
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
  __shared__ int s_bad;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  PLL16* inbox = buf;                                  // [2][G][G][hw]
  PLL16* slot = buf;                                   // mode 1: [2][G][hw]
  PLL16* gslot = buf + (size_t)2 * G * hw;             // mode 1: [2][leaders][hw]
  for (int s = 0; s < iters; ++s) {
    const uint32_t stamp = epoch + s;
    const int par = s & 1;
    if (mode == 0) {
      for (int e = tid; e < G * hw; e += 512) {
        const int dst = e / hw, w = e - dst * hw;
        pst(inbox + (((size_t)par * G + dst) * G + cta) * hw + w, (uint32_t)s, (uint32_t)w, stamp);
      }
      for (int e = tid; e < G * hw; e += 512)
        if (!pld(inbox + ((size_t)par * G + cta) * G * hw + e, stamp)) s_bad = 1;
    } else {
      if (tid < hw) pst(slot + ((size_t)par * G + cta) * hw + tid, (uint32_t)s, (uint32_t)tid, stamp);
      if (cta < leaders) {
        // group of leader c: CTAs c, c + leaders, c + 2*leaders, ...
        const int members = (G - cta + leaders - 1) / leaders;
        for (int e = tid; e < members * hw; e += 512) {
          const int mbr = cta + (e / hw) * leaders, w = e % hw;
          if (!pld(slot + ((size_t)par * G + mbr) * hw + w, stamp)) s_bad = 1;
        }
        __syncthreads();
        if (tid < hw) pst(gslot + ((size_t)par * leaders + cta) * hw + tid, (uint32_t)s, (uint32_t)tid, stamp);
      }
      for (int e = tid; e < leaders * hw; e += 512)
        if (!pld(gslot + (size_t)par * leaders * hw + e, stamp)) s_bad = 1;
    }
    __syncthreads();
    if (s_bad) break;
  }
  if (tid == 0 && s_bad) *fail = 1;
}

}  // namespace

extern "C" {

int bra_probe_exchange2(bra_ctx* ctx, int ctas, int hw, int mode, int leaders, int iters, double* usec) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(ctas >= 1 && ctas <= ctx->num_sms, 2, "ctas");
  BRA_CHECK_ARG(hw >= 1 && hw <= 8, 3, "hw");
  BRA_CHECK_ARG(leaders >= 1 && leaders <= ctas, 5, "leaders");
  BRA_CUDA(cudaSetDevice(ctx->device));
  const size_t words = (size_t)2 * ctas * ctas * hw + 64;
  BRA_CUDA(ctx->scratch3.reserve(words * 16 + 16));
  BRA_CUDA(cudaMemsetAsync(ctx->scratch3.p, 0, words * 16 + 16, ctx->stream));
  PLL16* buf = ctx->scratch3.as<PLL16>();
  int* fail = reinterpret_cast<int*>(buf + words);
  uint32_t epoch = 1;
  cudaEvent_t e0, e1;
  BRA_CUDA(cudaEventCreate(&e0));
  BRA_CUDA(cudaEventCreate(&e1));
  void* args[] = {(void*)&buf, (void*)&iters, (void*)&epoch, (void*)&hw, (void*)&mode, (void*)&leaders, (void*)&fail};
  BRA_CUDA(cudaEventRecord(e0, ctx->stream));
  BRA_CUDA(cudaLaunchCooperativeKernel((void*)exchange2_kernel, dim3(ctas), dim3(512), args, 0, ctx->stream));
  BRA_CUDA(cudaEventRecord(e1, ctx->stream));
  BRA_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  BRA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  ctx->launches++;
  int hfail = 0;
  BRA_CUDA(cudaMemcpy(&hfail, fail, 4, cudaMemcpyDeviceToHost));
  if (hfail) {
    ctx->set_error("exchange2 probe timed out");
    return BRA_ERR_INTERNAL;
  }
  *usec = (double)ms * 1e3 / iters;
  return BRA_OK;
}

int bra_probe_fp64_peak(bra_ctx* ctx, double* out) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(out != nullptr, 2, "out");
  BRA_CUDA(cudaSetDevice(ctx->device));
  // flops per warp per loop iteration (8 chains):
  //  m8n8k4: 2 mma * 8*8*4*2 = 1024 per chain; DFMA: 4 fma * 32 lanes * 2 = 256; m16n8k4: 1024; k8: 2048; k16: 4096
  int rc;
  if ((rc = run_peak<0>(ctx, 8 * 1024.0, &out[0]))) return rc;
  if ((rc = run_peak<1>(ctx, 8 * 256.0, &out[1]))) return rc;
  if ((rc = run_peak<2>(ctx, 8 * 1024.0, &out[2]))) return rc;
  if ((rc = run_peak<3>(ctx, 8 * 2048.0, &out[3]))) return rc;
  if ((rc = run_peak<4>(ctx, 8 * 4096.0, &out[4]))) return rc;
  return BRA_OK;
}

int bra_probe_exchange_latency(bra_ctx* ctx, int ctas, int iters, double* usec) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(ctas >= 1 && ctas <= ctx->num_sms, 2, "ctas");
  BRA_CHECK_ARG(iters >= 1, 3, "iters");
  BRA_CHECK_ARG(usec != nullptr, 4, "usec");
  BRA_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)2 * ctas * 4 * 16 + 16;
  BRA_CUDA(ctx->scratch3.reserve(bytes));
  BRA_CUDA(cudaMemsetAsync(ctx->scratch3.p, 0, bytes, ctx->stream));
  PLL16* rec = ctx->scratch3.as<PLL16>();
  int* fail = reinterpret_cast<int*>(rec + (size_t)2 * ctas * 4);
  uint32_t epoch = 1;
  cudaEvent_t e0, e1;
  BRA_CUDA(cudaEventCreate(&e0));
  BRA_CUDA(cudaEventCreate(&e1));
  void* args[] = {(void*)&rec, (void*)&iters, (void*)&epoch, (void*)&fail};
  BRA_CUDA(cudaEventRecord(e0, ctx->stream));
  BRA_CUDA(cudaLaunchCooperativeKernel((void*)exchange_kernel, dim3(ctas), dim3(512), args, 0, ctx->stream));
  BRA_CUDA(cudaEventRecord(e1, ctx->stream));
  BRA_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  BRA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  ctx->launches++;
  int hfail = 0;
  BRA_CUDA(cudaMemcpy(&hfail, fail, 4, cudaMemcpyDeviceToHost));
  if (hfail) {
    ctx->set_error("exchange probe timed out");
    return BRA_ERR_INTERNAL;
  }
  *usec = (double)ms * 1e3 / iters;
  return BRA_OK;
}

}  // extern "C"
