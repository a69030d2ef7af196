// Microbenchmark: issue cost and execution rate of the tcgen05.mma shapes the attention kernel uses, from ONE CTA:
//   mode 0  SS  M128 N128 K16  (S = Q K^T step; A, B K-major in smem)
//   mode 1  SS  M128 N64  K16  (O += P V step as of round 1: P in smem, V MN-major)
//   mode 2  TS  M128 N64  K16  (P read from TMEM, V MN-major in smem)
//   mode 3  attention mix SS:  4 x mode0 + 8 x mode1 per "tile"
//   mode 4  attention mix TS:  4 x mode0 + 8 x mode2 per "tile"
// each optionally with 8 extra warps storing 16-byte vectors to shared memory (the P-tile writes of the softmax warps).
// Also checks the TS operand layout: P written with tcgen05.st as packed bf16 pairs (col c = keys 2c, 2c+1).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../this_and_that_vdm_b200/csrc/ptx.cuh"
using namespace ttvdm;

constexpr int kTile = 16384;

__global__ void __launch_bounds__(384, 1) rate(int mode, int n_tiles, int hammer, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;               // 2 tiles (Q / P)
  uint8_t* sB = smem + 2 * kTile;   // 2 tiles (K / V)
  uint8_t* sH = smem + 4 * kTile;   // 4 tiles hammered by the store warps
  __shared__ uint64_t done, opbar[2], probe_bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&done, 1); mbar_init(&opbar[0], 1); mbar_init(&opbar[1], 1); mbar_init(&probe_bar, 1); mbar_fence_init(); stop = 0; }
  if (warp == 2) tmem_alloc<512>(&slot);
  for (int i = threadIdx.x; i < 8 * kTile / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t id_qk = make_idesc_bf16(128, 128, 0, 0), id_pv = make_idesc_bf16(128, 64, 0, 1);
  if (warp == 1) {
    if (lane == 0) {
      uint32_t sink = 0;
      long long t0 = clock64();
      for (int it = 0; it < n_tiles; ++it) {
        const int b = it & 1;
        if (mode == 0 || mode == 3 || mode == 4) {
          const uint64_t a = make_sdesc_sw128(smem_u32(sA + b * kTile), 16, 1024);
          const uint64_t k = make_sdesc_sw128(smem_u32(sB + b * kTile), 16, 1024);
          const int reps = (mode == 0) ? 12 : 4;
          for (int i = 0; i < reps; ++i) tc_mma_ss(tm + b * 128, a + 2 * (i & 3), k + 2 * (i & 3), id_qk, i != 0);
        }
        if (mode == 1 || mode == 3) {
          const int reps = (mode == 1) ? 24 : 8;
          for (int i = 0; i < reps; ++i) {
            const int k = i & 7;
            const uint64_t p = make_sdesc_sw128(smem_u32(sA) + (k >> 2) * kTile + (k & 3) * 32, 16, 1024);
            const uint64_t v = make_sdesc_sw128(smem_u32(sB + b * kTile) + k * 2048, 16, 1024);
            tc_mma_ss(tm + 256 + b * 64, p, v, id_pv, 1);
          }
        }
        if (mode == 5) {  // SS 128x64x16, K-major B (V^T tile)
          const uint32_t id = make_idesc_bf16(128, 64, 0, 0);
          for (int i = 0; i < 24; ++i) {
            const int k = i & 7;
            const uint64_t p = make_sdesc_sw128(smem_u32(sA) + (k >> 2) * kTile + (k & 3) * 32, 16, 1024);
            const uint64_t v = make_sdesc_sw128(smem_u32(sB) + (k >> 2) * kTile + (k & 3) * 32, 16, 1024);
            tc_mma_ss(tm + 256 + b * 64, p, v, id, 1);
          }
        }
        if (mode == 6) {  // SS 128x128x16, MN-major B
          const uint32_t id = make_idesc_bf16(128, 128, 0, 1);
          for (int i = 0; i < 12; ++i) {
            const int k = i & 7;
            const uint64_t p = make_sdesc_sw128(smem_u32(sA) + (k >> 2) * kTile + (k & 3) * 32, 16, 1024);
            const uint64_t v = make_sdesc_sw128(smem_u32(sB) + k * 2048, kTile, 1024);
            tc_mma_ss(tm + b * 128, p, v, id, 1);
          }
        }
        if (mode == 7) {  // SS 128x256x16, K-major B
          const uint32_t id = make_idesc_bf16(128, 256, 0, 0);
          for (int i = 0; i < 6; ++i) {
            const int k = i & 3;
            const uint64_t p = make_sdesc_sw128(smem_u32(sA) + k * 32, 16, 1024);
            const uint64_t v = make_sdesc_sw128(smem_u32(sB) + k * 32, 16, 1024);
            tc_mma_ss(tm + b * 256, p, v, id, 1);
          }
        }
        if (mode == 8) {  // TS 128x64x16, K-major B
          const uint32_t id = make_idesc_bf16(128, 64, 0, 0);
          for (int i = 0; i < 24; ++i) {
            const int k = i & 7;
            const uint64_t v = make_sdesc_sw128(smem_u32(sB) + (k >> 2) * kTile + (k & 3) * 32, 16, 1024);
            tc_mma_ts(tm + 256 + b * 64, tm + 384 + b * 64 + k * 8, v, id, 1);
          }
        }
        if (mode >= 9 && mode <= 14) {  // N=64, independent accumulator chains
          const int nacc = (mode == 9 || mode == 12) ? 2 : (mode == 10 || mode == 13) ? 4 : 1;
          const bool mn = mode <= 11, ts = (mode == 11 || mode == 14);
          const uint32_t id = make_idesc_bf16(128, 64, 0, mn ? 1 : 0);
          for (int i = 0; i < 24; ++i) {
            const int k = i & 7;
            const uint64_t p = make_sdesc_sw128(smem_u32(sA) + (k >> 2) * kTile + (k & 3) * 32, 16, 1024);
            const uint64_t v = mn ? make_sdesc_sw128(smem_u32(sB + b * kTile) + k * 2048, 16, 1024)
                                  : make_sdesc_sw128(smem_u32(sB) + (k >> 2) * kTile + (k & 3) * 32, 16, 1024);
            if (ts) tc_mma_ts(tm + 256 + (i & 1) * 64, tm + 384 + k * 8, v, id, 1);
            else tc_mma_ss(tm + 256 + (i % nacc) * 64, p, v, id, 1);
          }
        }
        if (mode == 15) {  // N=128 K-major, 2 accumulators
          const uint64_t a = make_sdesc_sw128(smem_u32(sA + b * kTile), 16, 1024);
          const uint64_t k = make_sdesc_sw128(smem_u32(sB + b * kTile), 16, 1024);
          for (int i = 0; i < 12; ++i) tc_mma_ss(tm + (i & 1) * 128, a + 2 * (i & 3), k + 2 * (i & 3), id_qk, 1);
        }
        if (mode >= 16 && mode <= 21) {  // scheduler-op cost: 4 MMAs + commit (+ wait / probes)
          const uint64_t a = make_sdesc_sw128(smem_u32(sA + b * kTile), 16, 1024);
          const uint64_t k = make_sdesc_sw128(smem_u32(sB + b * kTile), 16, 1024);
          for (int i = 0; i < 4; ++i) tc_mma_ss(tm + b * 128, a + 2 * i, k + 2 * i, id_qk, i != 0);
          tc_commit(&opbar[b]);
          if (mode == 16) { while (!mbar_test_wait(&opbar[b], (it >> 1) & 1)) {} }        // full drain each op
          if (mode == 17) { while (!mbar_try_wait(&opbar[b], (it >> 1) & 1)) {} }
          if (mode == 18) { sink += mbar_test_wait(&probe_bar, 1); }                      // one probe of a completed barrier
          if (mode == 19) { sink += mbar_test_wait6(&probe_bar, 1, &probe_bar, 1, &probe_bar, 1, &probe_bar, 1, &probe_bar, 1, &probe_bar, 1); }
          if (mode == 20) { for (int q = 0; q < 6; ++q) sink += mbar_test_wait(&probe_bar, 1); }
          // mode 21: nothing (issue + commit only)
        }
        if (mode >= 22 && mode <= 26) {
          const uint64_t a = make_sdesc_sw128(smem_u32(sA + b * kTile), 16, 1024);
          const uint64_t k = make_sdesc_sw128(smem_u32(sB + b * kTile), 16, 1024);
          for (int i = 0; i < 4; ++i) tc_mma_ss(tm + b * 128, a + 2 * i, k + 2 * i, id_qk, i != 0);
          if (mode == 24) tc_commit(&opbar[b]);
          if (mode == 26) tc_commit(&opbar[0]);
          for (int i = 0; i < 4; ++i) tc_mma_ss(tm + b * 128, a + 2 * i, k + 2 * i, id_qk, mode == 25 ? 1 : i != 0);
          if (mode == 23 || mode == 24) tc_commit(&opbar[b]);
          if (mode == 26) tc_commit(&opbar[1]);
        }
        if (mode == 2 || mode == 4) {
          const int reps = (mode == 2) ? 24 : 8;
          for (int i = 0; i < reps; ++i) {
            const int k = i & 7;
            const uint64_t v = make_sdesc_sw128(smem_u32(sB + b * kTile) + k * 2048, 16, 1024);
            tc_mma_ts(tm + 256 + b * 64, tm + 384 + b * 64 + k * 8, v, id_pv, 1);
          }
        }
      }
      long long t1 = clock64();
      tc_commit(&done);
      while (!mbar_try_wait(&done, 0)) {}
      long long t2 = clock64();
      out[0] = t1 - t0; out[3] = sink;
      out[1] = t2 - t0;
      stop = 1;
    }
  } else if (warp >= 4 && (hammer & 6)) {
    uint32_t acc = 0;
    float f = float(threadIdx.x) * 1e-3f;
    const uint32_t ta = tm + (uint32_t((warp & 3) * 32) << 16) + 128 * ((warp >> 3) & 1);
    while (!stop) {
      if (hammer & 2) {
        uint32_t v[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32(ta + c * 32, v[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) acc ^= v[c][i];
      }
      if (hammer & 4) {
#pragma unroll
        for (int i = 0; i < 128; ++i) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f)); }
      }
      if (hammer & 1) {
        const int r = (warp & 3) * 32 + lane;
        uint8_t* row = sH + (warp >> 3) * 2 * kTile + r * 128;
#pragma unroll
        for (int c = 0; c < 16; ++c)
          *reinterpret_cast<uint4*>(row + (c >> 3 & 1) * kTile + ((c ^ (r & 7)) & 7) * 16) = make_uint4(acc, acc + 1, acc + 2, c);
      }
    }
    out[2] = acc + __float_as_uint(f);
  } else if (warp >= 4 && hammer) {
    // 16-byte swizzled row stores like the softmax warps' P writes
    const int r = (warp & 3) * 32 + lane;
    uint8_t* row = sH + (warp >> 3) * 2 * kTile + r * 128;
    uint32_t x = threadIdx.x;
    while (!stop) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        *reinterpret_cast<uint4*>(row + (c >> 2 & 1) * kTile + ((c ^ (r & 7)) & 7) * 16) = make_uint4(x, x + 1, x + 2, x + 3);
        x = x * 1664525u + 1013904223u;
      }
      if (hammer > 1) {  // paced: ~hammer FMAs between stores
        float f = __uint_as_float(x | 0x3f800000u);
        for (int q = 0; q < hammer; ++q) f = fmaf(f, 1.0001f, 0.5f);
        x ^= __float_as_uint(f);
      }
    }
    out[2] = x;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tm);
}

// ---- TS layout check: O[128 x 64] = P[128 x 128] V[128 x 64], P from TMEM
__global__ void __launch_bounds__(128, 1) ts_check(float* o_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __nv_bfloat16* sV = reinterpret_cast<__nv_bfloat16*>(smem);  // [128 keys][64 d], 128B-swizzled rows
  __shared__ uint64_t done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&done, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<256>(&slot);
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int key = i / 64, d = i % 64;
    const float val = float(((key * 3 + d * 5) % 7) - 3);
    const int chunk = (d / 8) ^ (key & 7);
    sV[key * 64 + chunk * 8 + (d % 8)] = __float2bfloat16(val);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const int r = threadIdx.x;
  const uint32_t lane_addr = uint32_t(warp * 32) << 16;
  // P[r][k] = ((r + 2k) % 5) - 2, packed pairs (k even -> low half)
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t pk[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k0 = (c * 32 + i) * 2;
      pk[i] = pack_bf16(float(((r + 2 * k0) % 5) - 2), float(((r + 2 * (k0 + 1)) % 5) - 2));
    }
    tmem_st_32x32(tm + 128 + c * 32 + lane_addr, pk);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t id_pv = make_idesc_bf16(128, 64, 0, 1);
    for (int k = 0; k < 8; ++k) {
      const uint64_t v = make_sdesc_sw128(smem_u32(smem) + k * 2048, 16, 1024);
      tc_mma_ts(tm, tm + 128 + k * 8, v, id_pv, k != 0);
    }
    tc_commit(&done);
  }
  mbar_wait(&done, 0);
  tc_fence_after();
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(tm + c * 32 + lane_addr, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) o_out[r * 64 + c * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tm);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * kTile + 1024);
  const char* names[27] = {"SS 128x128x16 (x12)", "SS 128x64x16 MN-B (x24)", "TS 128x64x16 MN-B (x24)", "mix SS 4 QK + 8 PV", "mix TS 4 QK + 8 PV",
                          "SS 128x64x16 K-B (x24)", "SS 128x128x16 MN-B (x12)", "SS 128x256x16 K-B (x6)", "TS 128x64x16 K-B (x24)",
                          "SS N64 MN-B 2acc (x24)", "SS N64 MN-B 4acc (x24)", "TS N64 MN-B 2acc (x24)", "SS N64 K-B 2acc (x24)", "SS N64 K-B 4acc (x24)", "TS N64 K-B 2acc (x24)", "SS N128 K-B 2acc (x12)",
 "op: 4QK+commit+test_wait drain", "op: 4QK+commit+try_wait drain", "op: 4QK+commit+1 probe", "op: 4QK+commit+probe6 batched", "op: 4QK+commit+6 probes serial", "op: 4QK+commit",
 "4+4 QK no commit", "4+4 QK, commit", "4 QK commit 4 QK commit", "4+4 QK all accumulate", "4 QK commit(b0) 4 QK commit(b1)"};
  const int hammers[] = {0};
  for (int hammer : hammers) {
    for (int mode = 16; mode < 27; ++mode) {
      const int n_tiles = 200;
      rate<<<1, 384, 8 * kTile + 1024>>>(mode, n_tiles, hammer, d); cudaDeviceSynchronize();
      rate<<<1, 384, 8 * kTile + 1024>>>(mode, n_tiles, hammer, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("hammer %d(st%d tmld%d mufu%d)  %-26s issue %.0f cyc/tile  complete %.0f cyc/tile (%s)\n", hammer, hammer & 1, hammer >> 1 & 1, hammer >> 2 & 1, names[mode], h[0] / double(n_tiles), h[1] / double(n_tiles), cudaGetErrorString(e));
    }
  }
  // TS layout check
  float* o; cudaMalloc(&o, 128 * 64 * 4);
  cudaFuncSetAttribute(ts_check, cudaFuncAttributeMaxDynamicSharedMemorySize, kTile + 1024);
  ts_check<<<1, 128, kTile + 1024>>>(o);
  cudaError_t e = cudaDeviceSynchronize();
  static float h[128 * 64];
  cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r = 0; r < 128; ++r)
    for (int dd = 0; dd < 64; ++dd) {
      float ref = 0;
      for (int k = 0; k < 128; ++k) ref += float(((r + 2 * k) % 5) - 2) * float(((k * 3 + dd * 5) % 7) - 3);
      if (ref != h[r * 64 + dd]) { if (bad < 5) printf("  mismatch r=%d d=%d got %g want %g\n", r, dd, h[r * 64 + dd], ref); ++bad; }
    }
  printf("TS layout check: %d mismatches (%s)\n", bad, cudaGetErrorString(e));
  return 0;
}
