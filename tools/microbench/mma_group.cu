// Microbenchmark: cost of a GROUP of n back-to-back tcgen05.mma (SS, M128 N=64/128/256 K16, one accumulator chain) as a
// function of n, issued (a) under `if (threadIdx.x == 32)` and (b) under `if (elect_one())` (elect.sync).
// With (a) ptxas cannot prove the descriptor operands warp-uniform and wraps every MMA in ELECT + R2UR + branch
// instructions: the measured cost is ~65-85 cycles per MMA whatever N is (issue-bound) plus ~200 cycles per group;
// with (b) descriptors live in uniform registers and the cost is the tensor pipe's (N/2 cycles per MMA at M = 128).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../this_and_that_vdm_b200/csrc/ptx.cuh"
using namespace ttvdm;
constexpr int kTile = 16384;

template <bool kElect>
__global__ void __launch_bounds__(128, 1) grp(int n, int sw, int commit, int N, int groups, long long* out, int extra = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done, bar[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&done, 1); mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  if (warp == 2) tmem_alloc<512>(&slot);
  for (int i = threadIdx.x; i < 4 * kTile / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (kElect ? (warp == 1 && elect_one()) : (threadIdx.x == 32)) {
    const uint32_t id = make_idesc_bf16(128, N, 0, 0);
    const uint64_t a0 = make_sdesc_sw128(smem_u32(smem), 16, 1024), b0 = make_sdesc_sw128(smem_u32(smem + 2 * kTile), 16, 1024);
    long long t0 = clock64();
    for (int gI = 0; gI < groups; ++gI) {
      const int b = sw ? (gI & 1) : 0;
      const uint64_t a = a0 + b * (kTile >> 4), k = b0 + b * (kTile >> 4);
      for (int i = 0; i < n; ++i) tc_mma_ss(tm + b * 256, a + 2 * (i & 3), k + 2 * (i & 3), id, i != 0);
      if (commit) tc_commit(&bar[b]);
      if (extra & 1) tc_fence_after();
      if (extra & 2) out[4 + (gI & 1)] = clock64();
      if (extra & 4) out[6] += mbar_test_wait(&bar[b ^ 1], 1);
    }
    long long t1 = clock64();
    tc_commit(&done);
    while (!mbar_try_wait(&done, 0)) {}
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tm);
}

template <bool kElect>
static void sweep(long long* d, const char* how) {
  cudaFuncSetAttribute(grp<kElect>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kTile + 1024);
  const int groups = 256;
  for (int N : {64, 128, 256}) {
    printf("%-22s N=%3d (one chain per group, commit per group):", how, N);
    for (int n : {1, 2, 4, 8, 16, 32}) {
      grp<kElect><<<1, 128, 4 * kTile + 1024>>>(n, 0, 1, N, groups, d, 0); cudaDeviceSynchronize();
      grp<kElect><<<1, 128, 4 * kTile + 1024>>>(n, 0, 1, N, groups, d, 0);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("  n=%d: %.0f/grp (%.0f/mma)%s", n, h[1] / double(groups), h[1] / double(groups) / n, e ? cudaGetErrorString(e) : "");
    }
    printf("\n");
  }
}

int main() {
  long long* d; cudaMalloc(&d, 128); cudaMemset(d, 0, 128);
  sweep<false>(d, "if (threadIdx.x == 32)");
  sweep<true>(d, "if (elect_one())");
  return 0;
}
