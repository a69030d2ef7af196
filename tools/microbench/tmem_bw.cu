// Microbenchmark: TMEM read bandwidth per SM (tcgen05.ld 32x32b.x32) vs number of reading warps.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../this_and_that_vdm_b200/csrc/ptx.cuh"
using namespace ttvdm;

__global__ void tmem_read(int iters, int waits_every, long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v[32];
    tmem_ld_32x32(base + ((i * 32) & 511), v);
    if ((i % waits_every) == waits_every - 1) tmem_ld_wait();
    acc += __uint_as_float(v[i & 31]);
  }
  tmem_ld_wait();
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

int main() {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    for (int we : {1, 4}) {
      tmem_read<<<148, warps * 32>>>(iters, we, cyc, sink);
      cudaDeviceSynchronize();
      tmem_read<<<148, warps * 32>>>(iters, we, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
      double bytes = (double)warps * iters * 32 * 32 * 4;
      printf("warps=%2d wait_every=%d cycles=%lld  -> %.1f B/clk/SM (%s)\n", warps, we, h[0], bytes / h[0], cudaGetErrorString(e));
    }
  }
  return 0;
}
