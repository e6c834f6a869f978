// Diagnostic micro-benchmark (not part of libbrapprox): tensor memory (TMEM) as a scratchpad for FP64 data --
// tcgen05.st / tcgen05.ld 32x32b bandwidth and latency per SM, alone and together with shared-memory traffic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_probe tools/tmem_probe.cu && tools/tmem_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tm_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mode 0: TMEM x4 loads (dependent: wait after each)   1: TMEM x4 loads, 8 in flight   2: x32 loads   3: x4 stores
// mode 4: shared-memory LDS.128 stream only            5: LDS.128 stream + TMEM x32 loads (interleaved)
__global__ void __launch_bounds__(512, 1) k_tmem(int mode, int iters, long long* cyc, double* sink) {
  __shared__ uint32_t s_base;
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = s_base;
  // warp w may touch lanes 32 (w % 4) ..; each warp gets its own 128-column window of its quadrant
  const uint32_t taddr0 = tb + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 128u;
  uint32_t z[4] = {(uint32_t)lane, 1u, 2u, 3u};
  for (int c = 0; c < 128; c += 4) tm_st4(taddr0 + c, z);
  tm_wait_st();
  double2* sm2 = reinterpret_cast<double2*>(smem);
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm2[i] = make_double2(1.0, 2.0);
  __syncthreads();
  double acc = 0.0;
  uint32_t acc_u = 0;
  long long t0 = clock64();
  if (mode == 0) {
    for (int it = 0; it < iters; ++it) {
      uint32_t r[4];
      tm_ld4(taddr0 + ((it * 4) & 127), r);
      tm_wait_ld();
      acc_u += r[0] + r[3];
    }
  } else if (mode == 1) {
    for (int it = 0; it < iters; it += 8) {
      uint32_t r[8][4];
#pragma unroll
      for (int u = 0; u < 8; ++u) tm_ld4(taddr0 + (((it + u) * 4) & 127), r[u]);
      tm_wait_ld();
#pragma unroll
      for (int u = 0; u < 8; ++u) acc_u += r[u][0] + r[u][3];
    }
  } else if (mode == 2) {
    for (int it = 0; it < iters; it += 8) {
      uint32_t r[32];
      tm_ld32(taddr0 + (((it >> 3) * 32) & 127), r);
      tm_wait_ld();
#pragma unroll
      for (int u = 0; u < 32; u += 4) acc_u += r[u];
    }
  } else if (mode == 3) {
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) tm_st4(taddr0 + (((it + u) * 4) & 127), z);
      tm_wait_st();
    }
  } else if (mode == 4 || mode == 5) {
    for (int it = 0; it < iters; it += 8) {
      double2 x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) x[u] = sm2[((it + u) * 32 + lane + warp * 7) & 8191];
      if (mode == 5) {
        uint32_t r[32];
        tm_ld32(taddr0 + (((it >> 3) * 32) & 127), r);
        tm_wait_ld();
#pragma unroll
        for (int u = 0; u < 32; u += 4) acc_u += r[u];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += x[u].x;
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + acc_u;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int main() {
  long long* cyc;
  double* sink;
  cudaMallocManaged(&cyc, 4096);
  cudaMalloc(&sink, 148 * 512 * 8);
  cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
  const char* names[] = {"tmem ld x4 dependent", "tmem ld x4, 8 in flight", "tmem ld x32", "tmem st x4, 8 in flight", "smem LDS.128 stream",
                         "smem LDS.128 + tmem ld x32"};
  for (int warps : {1, 4, 8, 16}) {
    for (int mode = 0; mode < 6; ++mode) {
      const int iters = 4096;
      k_tmem<<<1, warps * 32, 8192 * 16>>>(mode, iters, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("%s: %s\n", names[mode], cudaGetErrorString(e));
        return 1;
      }
      // bytes per iteration per warp: 512 (one x4 load of 32 lanes x 16 B, or one LDS.128); mode 5 moves 512 + 512
      const double bytes = (double)iters * 512.0 * warps * (mode == 5 ? 2 : 1);
      printf("warps %2d  %-28s %8lld cycles  %6.1f B/clk/SM  (%.0f cycles per op per warp)\n", warps, names[mode], cyc[0],
             bytes / (double)cyc[0], (double)cyc[0] / iters);
    }
  }
  return 0;
}
