// Microbenchmark: round-trip latency of mbarrier ping-pong between two warps of a CTA, (a) plain arrive,
// (b) with the reply sent through tcgen05.commit (no MMA in flight), (c) commit after one small tcgen05.mma.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../this_and_that_vdm_b200/csrc/ptx.cuh"
using namespace ttvdm;

__global__ void pingpong(int iters, int mode, long long* cycles) {
  __shared__ uint64_t a2b, b2a;
  __shared__ uint32_t slot;
  __shared__ __align__(1024) uint8_t tile[2][16384];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&a2b, 1); mbar_init(&b2a, 1); mbar_fence_init(); }
  if (warp == 2) tmem_alloc<64>(&slot);
  for (int i = threadIdx.x; i < 2 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tile)[i] = 0;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t idesc = make_idesc_bf16(128, 64);
  long long t0 = clock64();
  if (warp == 0) {
    for (int i = 0; i < iters; ++i) {
      if (lane == 0) mbar_arrive(&a2b);
      mbar_wait(&b2a, i & 1);
    }
  } else if (warp == 1) {
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&a2b, i & 1);
      if (lane == 0) {
        if (mode == 0) mbar_arrive(&b2a);
        else {
          if (mode == 2) {
            tc_fence_after();
            tc_mma_ss(slot, make_sdesc_sw128(smem_u32(tile[0]), 16, 1024), make_sdesc_sw128(smem_u32(tile[1]), 16, 1024), idesc, 0);
          }
          tc_commit(&b2a);
        }
      }
      __syncwarp();
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<64>(slot);
}

int main() {
  long long* cyc; cudaMalloc(&cyc, 8);
  for (int mode = 0; mode < 3; ++mode) {
    pingpong<<<1, 96>>>(20000, mode, cyc); cudaDeviceSynchronize();
    pingpong<<<1, 96>>>(20000, mode, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("mode %d (%s): %.1f cycles per round trip (%s)\n", mode, mode == 0 ? "arrive/arrive" : mode == 1 ? "arrive/commit" : "arrive/mma+commit", h / 20000.0, cudaGetErrorString(e));
  }
  return 0;
}
