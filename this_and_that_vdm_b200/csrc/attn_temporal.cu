// K7 — temporal self-attention: F (<= 16) frames x F frames per (b, s, head), head_dim 64. ~0.05 % of the FLOPs
// and HBM-bound (8 * C bytes per token and layer). Rows stay in the (b, f, s) token order — frames are walked with
// stride S*ld instead of permuting the activation.
// One warp per item (b, s, head): Q, K, V (16 x 64 bf16 each, frames >= F zero-filled) land in the warp's
// shared-memory tiles with cp.async, S = Q K^T and O = P V run on warp-level mma.sync (m16n8k16, ldmatrix operands),
// the 16 x 16 softmax lives in the accumulator registers of the quad that owns each row, and the output leaves
// through shared memory in 16-byte row pieces. The scalar version of round 1 (one query frame per thread, K/V
// broadcast from shared memory) was bound by shared-memory instruction issue at 1.5-2.2 TB/s: every thread re-read
// all of K and V. tcgen05 is pointless here (M = 16); the kernel only has to keep HBM busy.
#include "common.h"
#include "ptx.cuh"

namespace ttvdm {

constexpr int kTaMaxF = 16;
// head dim kD = 64 / 128: rows of kD * 2 B of data + 16 B pad (conflict-free ldmatrix); 4 / 2 items (warps) per CTA so the
// static shared memory stays under 48 KB

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int kD>
__global__ void __launch_bounds__((kD == 64 ? 4 : 2) * 32)
attn_temporal_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                     const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ out, int ldq, int ldk, int ldv,
                     int ldo, int B, int F, int S, int heads, float scale) {
  constexpr int kTaWarps = kD == 64 ? 4 : 2;
  constexpr int kTaRowBytes = kD * 2 + 16;
  constexpr int kCh = kD / 8;  // 16-byte chunks per row
  __shared__ __align__(16) uint8_t tiles[kTaWarps][3][kTaMaxF * kTaRowBytes];  // per warp: Q (later O), K, V
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long items = (long long)B * S * heads;
  const long long item = (long long)blockIdx.x * kTaWarps + warp;
  if (item >= items) return;  // warps are independent (no CTA-wide barrier below)
  const int head = (int)(item % heads);
  const long long bs = item / heads;
  const int s = (int)(bs % S);
  const int b = (int)(bs / S);
  const long long row0 = ((long long)b * F) * S + s;  // frame 0 row; frame f is row0 + f*S
  const uint32_t sq = smem_u32(tiles[warp][0]), sk = smem_u32(tiles[warp][1]), sv = smem_u32(tiles[warp][2]);

  // ---- Q, K, V -> shared memory: 4 frames x 8 chunks of 16 B per instruction, frames >= F zero-filled
  {
    const int fr = lane >> 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int f = i * 4 + fr;
      const int ok = f < F ? 16 : 0;
      const long long row = row0 + (long long)(f < F ? f : 0) * S;
#pragma unroll
      for (int c8 = 0; c8 < kCh; c8 += 8) {
        const int ch = c8 + (lane & 7);
        const uint32_t off = f * kTaRowBytes + ch * 16;
        cp_async16(sq + off, q + row * ldq + head * kD + ch * 8, ok);
        cp_async16(sk + off, k + row * ldk + head * kD + ch * 8, ok);
        cp_async16(sv + off, v + row * ldv + head * kD + ch * 8, ok);
      }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncwarp();
  }

  // ---- S = Q K^T (16 x 16, fp32): 4 k-steps of 16 head-dim columns, 2 n-tiles of 8 key frames
  float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
  for (int kk = 0; kk < kD / 16; ++kk) {
    uint32_t a[4], bm[4];
    // A: matrices (rows 0-7 | 8-15) x (cols kk*16 .. +7 | +8 .. +15)
    ldsm_x4(sq + (lane & 15) * kTaRowBytes + kk * 32 + (lane >> 4) * 16, a);
    // B: key frames (lane/16)*8 + lane%8, head-dim halves ((lane/8)%2)*8: {b0, b1} of n-tile 0, {b0, b1} of n-tile 1
    ldsm_x4(sk + ((lane >> 4) * 8 + (lane & 7)) * kTaRowBytes + kk * 32 + ((lane >> 3) & 1) * 16, bm);
    mma_16816(sc[0], a, bm[0], bm[1]);
    mma_16816(sc[1], a, bm[2], bm[3]);
  }
  // ---- softmax over key frames: thread holds rows r0 = lane/4 (elements 0, 1) and r0 + 8 (elements 2, 3),
  //      key frames nt*8 + (lane%4)*2 + {0, 1}
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = nt * 8 + (lane & 3) * 2 + (e & 1);
      sc[nt][e] = col < F ? sc[nt][e] * scale : -INFINITY;
      mx[e >> 1] = fmaxf(mx[e >> 1], sc[nt][e]);
    }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
    mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
  }
  float sum[2] = {0.f, 0.f};
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      sc[nt][e] = __expf(sc[nt][e] - mx[e >> 1]);  // exp(-inf) = 0 for padded key frames
      sum[e >> 1] += sc[nt][e];
    }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
    sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
  }
  const float inv[2] = {1.f / sum[0], 1.f / sum[1]};
  // P (normalised, bf16) as the A operand of O = P V: the accumulator layout of two n-tiles IS the A fragment of k = 16
  uint32_t pa[4];
  pa[0] = pack_bf16(sc[0][0] * inv[0], sc[0][1] * inv[0]);
  pa[1] = pack_bf16(sc[0][2] * inv[1], sc[0][3] * inv[1]);
  pa[2] = pack_bf16(sc[1][0] * inv[0], sc[1][1] * inv[0]);
  pa[3] = pack_bf16(sc[1][2] * inv[1], sc[1][3] * inv[1]);

  // ---- O = P V (16 x 64): 8 n-tiles of 8 head-dim columns, k = 16 key frames; V^T fragments via ldmatrix.trans
  float o[kD / 8][4];
#pragma unroll
  for (int dt = 0; dt < kD / 8; ++dt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[dt][e] = 0.f;
#pragma unroll
  for (int d2 = 0; d2 < kD / 16; ++d2) {
    uint32_t bm[4];
    // matrices: (key frames 0-7 | 8-15) x head-dim tile 2*d2, then the same for tile 2*d2 + 1
    ldsm_x4_trans(sv + (lane & 15) * kTaRowBytes + (d2 * 2 + (lane >> 4)) * 16, bm);
    mma_16816(o[d2 * 2], pa, bm[0], bm[1]);
    mma_16816(o[d2 * 2 + 1], pa, bm[2], bm[3]);
  }
  // ---- O -> the warp's Q tile (Q is dead) -> global in 16-byte row pieces
  __syncwarp();
  {
    uint8_t* so = tiles[warp][0];
    const int r0 = lane >> 2, c0 = (lane & 3) * 2;
#pragma unroll
    for (int dt = 0; dt < kD / 8; ++dt) {
      *reinterpret_cast<uint32_t*>(so + r0 * kTaRowBytes + (dt * 8 + c0) * 2) = pack_bf16(o[dt][0], o[dt][1]);
      *reinterpret_cast<uint32_t*>(so + (r0 + 8) * kTaRowBytes + (dt * 8 + c0) * 2) = pack_bf16(o[dt][2], o[dt][3]);
    }
    __syncwarp();
    const int fr = lane >> 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int f = i * 4 + fr;
      if (f < F) {
#pragma unroll
        for (int c8 = 0; c8 < kCh; c8 += 8) {
          const int ch = c8 + (lane & 7);
          const uint4 u = *reinterpret_cast<const uint4*>(so + f * kTaRowBytes + ch * 16);
          *reinterpret_cast<uint4*>(out + (row0 + (long long)f * S) * ldo + head * kD + ch * 8) = u;
        }
      }
    }
  }
}

}  // namespace ttvdm

using namespace ttvdm;

extern "C" int ttvdm_attn_temporal(const ttvdm_tattn_params* p, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!p || !p->q || !p->k || !p->v || !p->out) return fail(TTVDM_ERR_SHAPE, "attn_temporal: null");
  if (p->F < 1 || p->F > kTaMaxF) return fail(TTVDM_ERR_SHAPE, "attn_temporal: F=%d (max %d)", p->F, kTaMaxF);
  if ((p->ldq | p->ldk | p->ldv | p->ldo) % 8 != 0) return fail(TTVDM_ERR_SHAPE, "attn_temporal: ld %% 8 != 0");
  if ((reinterpret_cast<uintptr_t>(p->q) | reinterpret_cast<uintptr_t>(p->k) | reinterpret_cast<uintptr_t>(p->v) |
       reinterpret_cast<uintptr_t>(p->out)) & 15)
    return fail(TTVDM_ERR_SHAPE, "attn_temporal: q/k/v/out must be 16-byte aligned");
  const long long items = (long long)p->B * p->S * p->heads;
  if (items <= 0) return fail(TTVDM_ERR_SHAPE, "attn_temporal: empty");
  const int hd = p->head_dim == 0 ? 64 : p->head_dim;
  if (hd != 64 && hd != 128) return fail(TTVDM_ERR_SHAPE, "attn_temporal: head_dim %d (64 or 128)", hd);
  const int warps = hd == 64 ? 4 : 2;
  const long long grid = (items + warps - 1) / warps;
  if (grid > 0x7fffffffLL) return fail(TTVDM_ERR_SHAPE, "attn_temporal: too many items");
  auto* kq = static_cast<const __nv_bfloat16*>(p->q);
  auto* kk = static_cast<const __nv_bfloat16*>(p->k);
  auto* kv = static_cast<const __nv_bfloat16*>(p->v);
  auto* ko = static_cast<__nv_bfloat16*>(p->out);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (hd == 64)
    attn_temporal_kernel<64><<<(int)grid, warps * 32, 0, st>>>(kq, kk, kv, ko, p->ldq, p->ldk, p->ldv, p->ldo, p->B, p->F,
                                                               p->S, p->heads, p->scale);
  else
    attn_temporal_kernel<128><<<(int)grid, warps * 32, 0, st>>>(kq, kk, kv, ko, p->ldq, p->ldk, p->ldv, p->ldo, p->B, p->F,
                                                                p->S, p->heads, p->scale);
  TTVDM_CHECK_LAUNCH("attn_temporal_kernel");
  return 0;
}
