// Microbenchmark: where do the ~200-300 "extra" cycles per short tcgen05.mma chain (mma_group.cu) come from?
// One CTA, `groups` back-to-back groups of n MMAs (SS, M128 K16), steady-state cycles per group for
//   mode 0: first MMA of a group overwrites (scale-d = 0), tcgen05.commit after every group   (= mma_group.cu)
//   mode 1: first MMA overwrites, NO commit between groups
//   mode 2: every MMA accumulates, commit after every group
//   mode 3: every MMA accumulates, no commit between groups
//   mode 4: like 0, but two issuer warps run their own groups concurrently on disjoint accumulators / operands
//           (reported per group of ONE issuer: equal to mode 0 => the chains of the two issuers overlap perfectly,
//            2x mode 0 => the pipe serialises them)
//   mode 5: like 0 with the A operand of every MMA of a group at a different smem address (no re-read of the same 4 KB)
#include <cstdio>
#include <cuda_runtime.h>
#include "../../this_and_that_vdm_b200/csrc/ptx.cuh"
using namespace ttvdm;
constexpr int kTile = 16384;

template <int kN>
__global__ void __launch_bounds__(160, 1) chain(int mode, int N, int groups, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done[2], bar[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&done[0], 1); mbar_init(&done[1], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<512>(&slot);
  for (int i = threadIdx.x; i < 8 * kTile / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const bool two = mode == 4;
  const int who = warp == 1 ? 0 : (warp == 3 ? 1 : -1);
  if (who >= 0 && (who == 0 || two) && elect_one()) {
    const uint32_t id = make_idesc_bf16(128, N, 0, 0);
    const uint64_t a0 = make_sdesc_sw128(smem_u32(smem + who * 4 * kTile), 16, 1024);
    const uint64_t b0 = make_sdesc_sw128(smem_u32(smem + who * 4 * kTile + 2 * kTile), 16, 1024);
    const bool over = mode == 0 || mode == 1 || mode == 4 || mode == 5;
    const bool commit = mode == 0 || mode == 2 || mode == 4 || mode == 5;
    const int astep = mode == 5 ? (2048 >> 4) : 0;  // next 8-row group of the tile: another 2 KB of smem
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
#pragma unroll
      for (int i = 0; i < kN; ++i)
        tc_mma_ss(tm + who * 256, a0 + 2 * (i & 3) + astep * (i >> 2), b0 + 2 * (i & 3), id, over ? i != 0 : true);
      if (commit) tc_commit(&bar[who * 2 + (g & 1)]);
    }
    tc_commit(&done[who]);
    while (!mbar_try_wait(&done[who], 0)) {}
    long long t2 = clock64();
    out[who] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tm);
}

template <int kN>
static void one(long long* d, int mode, int N) {
  const int groups = 256;
  cudaFuncSetAttribute(chain<kN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * kTile + 1024);
  for (int rep = 0; rep < 2; ++rep) {
    chain<kN><<<1, 160, 8 * kTile + 1024>>>(mode, N, groups, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf(" CUDA error %s", cudaGetErrorString(e)); return; }
  }
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const long long t = mode == 4 ? (h[0] > h[1] ? h[0] : h[1]) : h[0];
  printf("  n=%d: %.0f", kN, t / double(groups));
}

int main() {
  long long* d; cudaMalloc(&d, 128); cudaMemset(d, 0, 128);
  const char* names[6] = {"0 overwrite+commit", "1 overwrite, no commit", "2 accumulate+commit", "3 accumulate, no commit",
                          "4 two issuers (overwrite+commit)", "5 overwrite+commit, distinct A"};
  for (int N : {64, 80, 96, 128, 160, 192, 240, 256}) {
    for (int mode = 0; mode < 6; mode += (N == 64 || N == 128 || N == 256) ? 1 : 6) {
      printf("N=%3d  %-34s cycles/group:", N, names[mode]);
      one<1>(d, mode, N); one<2>(d, mode, N); one<4>(d, mode, N); one<8>(d, mode, N); one<16>(d, mode, N);
      printf("\n");
    }
  }
  return 0;
}
