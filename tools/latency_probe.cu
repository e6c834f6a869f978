// Diagnostic micro-benchmarks (not part of libbrapprox): FP64 dependent-op latencies, warp collectives,
// CTA barrier, and L2 store->load round trips, as seen by the persistent QRCP kernel's critical path.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/latency_probe tools/latency_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>

__global__ void k_alu(double* out, long long* cyc, double x0) {
  const int N = 256;
  double x = x0 + threadIdx.x * 1e-9, y = 1.0000001;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (t1 - t0);
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x + y;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = (t1 - t0);
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = sqrt(x + 2.0);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = (t1 - t0);
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = 1.0 / (x + 2.0);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = (t1 - t0);
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = rsqrt(x + 2.0);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = (t1 - t0);
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1 + (i & 15));
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = (t1 - t0);
  int q = (int)x0 + threadIdx.x;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) q = __reduce_max_sync(0xffffffffu, q + i) ^ threadIdx.x;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = (t1 - t0);
  t0 = clock64();
  for (int i = 0; i < N; ++i) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = (t1 - t0);
  __shared__ double sm[1024];
  sm[threadIdx.x] = x;
  __syncthreads();
  int idx = threadIdx.x;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) idx = (int)sm[idx & 511] & 511;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[8] = (t1 - t0);
  // float sqrt/div for comparison
  float f = (float)x0 + 1.5f;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) f = sqrtf(f + 2.0f);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[9] = (t1 - t0);
  out[threadIdx.x] = x + q + idx + f;
}

struct __align__(16) LL16 { uint32_t lo, s0, hi, s1; };
__device__ __forceinline__ void ll_store(LL16* p, uint32_t lo, uint32_t hi, uint32_t st) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(lo), "r"(st), "r"(hi), "r"(st) : "memory");
}
__device__ __forceinline__ void ll_wait(const LL16* p, uint32_t st, uint32_t& lo) {
  uint32_t s0, hi, s1;
  do {
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo), "=r"(s0), "=r"(hi), "=r"(s1) : "l"(p) : "memory");
  } while (s0 != st || s1 != st);
}

// ping-pong between CTA 0 and CTA `peer`: RTT in cycles
__global__ void k_pingpong(LL16* buf, int iters, int peer, long long* cyc) {
  if (threadIdx.x != 0) return;
  uint32_t lo;
  if (blockIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
      ll_store(buf + 0, i, i, i);
      ll_wait(buf + 64, i, lo);
    }
    cyc[0] = (clock64() - t0) / iters;
  } else if (blockIdx.x == peer) {
    for (int i = 1; i <= iters; ++i) {
      ll_wait(buf + 0, i, lo);
      ll_store(buf + 64, i, i, i);
    }
  }
}

struct __align__(32) LL32 { uint32_t w[8]; };
// all-to-all of one 32-byte word per (src, dst), G CTAs, 512 threads each (threads < G take part)
__global__ void k_alltoall(LL32* inbox, int iters, long long* cyc, int extra_sync) {
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
  __shared__ int sink;
  long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    const int par = it & 1;
    const uint32_t st = it;
    if (tid < G) {
      LL32* d = inbox + ((size_t)par * G + tid) * G + cta;
      asm volatile("st.relaxed.gpu.global.v8.b32 [%0], {%1,%2,%1,%2,%1,%2,%1,%2};" ::"l"(d), "r"((uint32_t)cta), "r"(st) : "memory");
      const LL32* s = inbox + ((size_t)par * G + cta) * G + tid;
      uint32_t a, s0, b, s1, c, s2, e, s3;
      do {
        asm volatile("ld.relaxed.gpu.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a), "=r"(s0), "=r"(b), "=r"(s1), "=r"(c), "=r"(s2), "=r"(e), "=r"(s3) : "l"(s) : "memory");
      } while (s0 != st || s1 != st || s2 != st || s3 != st);
      if (a == 0xffffffffu) sink = 1;
    }
    __syncthreads();
    for (int i = 0; i < extra_sync; ++i) __syncthreads();
  }
  if (tid == 0) cyc[cta] = (clock64() - t0) / iters;
}

// pull variant: every CTA stores ONE 32-byte word into a shared array, everybody polls the whole array
__global__ void k_pull(LL32* arr, int iters, long long* cyc) {
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
  __shared__ int sink;
  long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    const int par = it & 1;
    const uint32_t st = it;
    if (tid == 0) {
      LL32* d = arr + (size_t)par * G + cta;
      asm volatile("st.relaxed.gpu.global.v8.b32 [%0], {%1,%2,%1,%2,%1,%2,%1,%2};" ::"l"(d), "r"((uint32_t)cta), "r"(st) : "memory");
    }
    if (tid < G) {
      const LL32* s = arr + (size_t)par * G + tid;
      uint32_t a, s0, b, s1, c, s2, e, s3;
      do {
        asm volatile("ld.relaxed.gpu.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a), "=r"(s0), "=r"(b), "=r"(s1), "=r"(c), "=r"(s2), "=r"(e), "=r"(s3) : "l"(s) : "memory");
      } while (s0 != st || s1 != st || s2 != st || s3 != st);
      if (a == 0xffffffffu) sink = 1;
    }
    __syncthreads();
  }
  if (tid == 0) cyc[cta] = (clock64() - t0) / iters;
}

// atomic variant: red.max on a 64-bit key + arrival counter in the same 16-byte word; poll that single word
__global__ void k_atomic(unsigned long long* slots, int iters, long long* cyc) {
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
  __shared__ unsigned long long s_key;
  long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    unsigned long long* slot = slots + (size_t)(it & 3) * 16;       // 4 rotating slots, 128 B apart
    if (tid == 0) {
      const unsigned long long key = ((unsigned long long)it << 32) | (unsigned)((cta * 2654435761u) >> 8);
      asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(slot), "l"(key) : "memory");
      asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(slot + 1), "l"(1ull) : "memory");
      unsigned long long k, c;
      const unsigned long long want = (unsigned long long)G * ((it + 3) / 4);
      do {
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(k), "=l"(c) : "l"(slot) : "memory");
      } while (c < want);
      s_key = k;
    }
    __syncthreads();
  }
  if (tid == 0) cyc[cta] = (clock64() - t0) / iters + (s_key == 7);
}

int main() {
  cudaSetDevice(0);
  double* out;
  long long* cyc;
  cudaMalloc(&out, 8192);
  cudaMallocManaged(&cyc, 8192);
  k_alu<<<1, 512>>>(out, cyc, 1.25);
  cudaDeviceSynchronize();
  const char* names[] = {"dfma", "dadd", "dsqrt(+dadd)", "ddiv(+dadd)", "drsqrt(+dadd)", "shfl64+dadd", "redux.max+xor", "syncthreads512", "lds chase", "fsqrt(+fadd)"};
  for (int i = 0; i < 10; ++i) printf("%-16s %6.1f cycles/op\n", names[i], cyc[i] / 256.0);
  LL16* buf;
  cudaMalloc(&buf, 1 << 20);
  for (int peer : {1, 2, 37, 74, 147}) {
    cudaMemset(buf, 0, 1 << 20);
    void* args[] = {(void*)&buf, nullptr, (void*)&peer, (void*)&cyc};
    int iters = 2000;
    args[1] = &iters;
    cudaLaunchCooperativeKernel((void*)k_pingpong, dim3(148), dim3(32), args, 0, 0);
    cudaDeviceSynchronize();
    printf("pingpong cta0<->cta%-3d  RTT %lld cycles (2 store->load hops)\n", peer, cyc[0]);
  }
  LL32* inbox;
  cudaMalloc(&inbox, (size_t)2 * 148 * 148 * 32);
  for (int G : {2, 16, 74, 148}) {
    for (int extra : {0}) {
      cudaMemset(inbox, 0, (size_t)2 * 148 * 148 * 32);
      int iters = 2000;
      void* args[] = {(void*)&inbox, (void*)&iters, (void*)&cyc, (void*)&extra};
      cudaError_t e = cudaLaunchCooperativeKernel((void*)k_alltoall, dim3(G), dim3(512), args, 0, 0);
      cudaDeviceSynchronize();
      long long mx = 0;
      for (int i = 0; i < G; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
      printf("alltoall G=%-3d  %lld cycles/iter (%s)\n", G, mx, cudaGetErrorString(e));
    }
  }
  for (int G : {16, 74, 148}) {
    cudaMemset(inbox, 0, (size_t)2 * 148 * 148 * 32);
    int iters = 2000;
    void* args[] = {(void*)&inbox, (void*)&iters, (void*)&cyc};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k_pull, dim3(G), dim3(512), args, 0, 0);
    cudaDeviceSynchronize();
    long long mx = 0;
    for (int i = 0; i < G; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
    printf("pull     G=%-3d  %lld cycles/iter (%s)\n", G, mx, cudaGetErrorString(e));
  }
  for (int G : {16, 74, 148}) {
    cudaMemset(inbox, 0, 4096);
    int iters = 2000;
    unsigned long long* slots = (unsigned long long*)inbox;
    void* args[] = {(void*)&slots, (void*)&iters, (void*)&cyc};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k_atomic, dim3(G), dim3(512), args, 0, 0);
    cudaDeviceSynchronize();
    long long mx = 0;
    for (int i = 0; i < G; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
    printf("atomic   G=%-3d  %lld cycles/iter (%s)\n", G, mx, cudaGetErrorString(e));
  }
  return 0;
}
