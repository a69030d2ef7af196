// K7 — temporal self-attention: F (<= 32) frames x F frames per (b, s, head), head_dim 64 (or 128). ~0.05 % of the FLOPs
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

constexpr int kTaMaxF = 32;
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

// kD = head dim (64 / 128), kMT = 16-frame tiles (1: F <= 16, the reference's 14 frames; 2: F <= 32, SVD-XT's 25)
template <int kD, int kMT>
constexpr int ta_warps() { return 4 / (kMT * (kD / 64)) > 0 ? 4 / (kMT * (kD / 64)) : 1; }

template <int kD, int kMT>
__global__ void __launch_bounds__(ta_warps<kD, kMT>() * 32)
attn_temporal_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                     const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ out, int ldq, int ldk, int ldv,
                     int ldo, int B, int F, int S, int heads, float scale) {
  constexpr int kTaWarps = ta_warps<kD, kMT>();
  constexpr int kTaRowBytes = kD * 2 + 16;
  constexpr int kCh = kD / 8;        // 16-byte chunks per row
  constexpr int kRows = 16 * kMT;    // padded frames
  __shared__ __align__(16) uint8_t tiles[kTaWarps][3][kRows * kTaRowBytes];  // per warp: Q (later O), K, V
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
    for (int i = 0; i < 4 * kMT; ++i) {
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

  // ---- S = Q K^T (kRows x kRows, fp32): per 16-row query tile mt, k-steps of 16 head-dim columns, n-tiles of 8 key frames
  float sc[kMT][2 * kMT][4];
#pragma unroll
  for (int mt = 0; mt < kMT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2 * kMT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) sc[mt][nt][e] = 0.f;
#pragma unroll
  for (int kk = 0; kk < kD / 16; ++kk) {
    uint32_t bm[kMT][4];
    // B: key frames j*16 + (lane/16)*8 + lane%8, head-dim halves ((lane/8)%2)*8: {b0, b1} of n-tile 2j, {b0, b1} of 2j + 1
#pragma unroll
    for (int j = 0; j < kMT; ++j)
      ldsm_x4(sk + (j * 16 + (lane >> 4) * 8 + (lane & 7)) * kTaRowBytes + kk * 32 + ((lane >> 3) & 1) * 16, bm[j]);
#pragma unroll
    for (int mt = 0; mt < kMT; ++mt) {
      uint32_t a[4];
      // A: matrices (rows 0-7 | 8-15 of the tile) x (cols kk*16 .. +7 | +8 .. +15)
      ldsm_x4(sq + (mt * 16 + (lane & 15)) * kTaRowBytes + kk * 32 + (lane >> 4) * 16, a);
#pragma unroll
      for (int j = 0; j < kMT; ++j) {
        mma_16816(sc[mt][2 * j], a, bm[j][0], bm[j][1]);
        mma_16816(sc[mt][2 * j + 1], a, bm[j][2], bm[j][3]);
      }
    }
  }
  // ---- softmax over key frames: per tile a thread holds rows r0 = lane/4 (elements 0, 1) and r0 + 8 (elements 2, 3),
  //      key frames nt*8 + (lane%4)*2 + {0, 1}; P (normalised, bf16) becomes the A operand of O = P V: the accumulator
  //      layout of two n-tiles IS the A fragment of one k = 16 step
  uint32_t pa[kMT][kMT][4];
#pragma unroll
  for (int mt = 0; mt < kMT; ++mt) {
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 2 * kMT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + (lane & 3) * 2 + (e & 1);
        sc[mt][nt][e] = col < F ? sc[mt][nt][e] * scale : -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], sc[mt][nt][e]);
      }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 2 * kMT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        sc[mt][nt][e] = __expf(sc[mt][nt][e] - mx[e >> 1]);  // exp(-inf) = 0 for padded key frames
        sum[e >> 1] += sc[mt][nt][e];
      }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
      sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
    }
    const float inv[2] = {1.f / sum[0], 1.f / sum[1]};
#pragma unroll
    for (int ks = 0; ks < kMT; ++ks) {
      pa[mt][ks][0] = pack_bf16(sc[mt][2 * ks][0] * inv[0], sc[mt][2 * ks][1] * inv[0]);
      pa[mt][ks][1] = pack_bf16(sc[mt][2 * ks][2] * inv[1], sc[mt][2 * ks][3] * inv[1]);
      pa[mt][ks][2] = pack_bf16(sc[mt][2 * ks + 1][0] * inv[0], sc[mt][2 * ks + 1][1] * inv[0]);
      pa[mt][ks][3] = pack_bf16(sc[mt][2 * ks + 1][2] * inv[1], sc[mt][2 * ks + 1][3] * inv[1]);
    }
  }

  // ---- O = P V (kRows x kD): n-tiles of 8 head-dim columns, k-steps of 16 key frames; V^T fragments via ldmatrix.trans.
  //      One query tile at a time (its accumulators go to the warp's Q tile — Q is dead — before the next tile starts).
  __syncwarp();
  uint8_t* so = tiles[warp][0];
#pragma unroll
  for (int mt = 0; mt < kMT; ++mt) {
    float o[kD / 8][4];
#pragma unroll
    for (int dt = 0; dt < kD / 8; ++dt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[dt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < kMT; ++ks)
#pragma unroll
      for (int d2 = 0; d2 < kD / 16; ++d2) {
        uint32_t bm[4];
        // matrices: (key frames 0-7 | 8-15 of the k-step) x head-dim tile 2*d2, then the same for tile 2*d2 + 1
        ldsm_x4_trans(sv + (ks * 16 + (lane & 15)) * kTaRowBytes + (d2 * 2 + (lane >> 4)) * 16, bm);
        mma_16816(o[d2 * 2], pa[mt][ks], bm[0], bm[1]);
        mma_16816(o[d2 * 2 + 1], pa[mt][ks], bm[2], bm[3]);
      }
    const int r0 = mt * 16 + (lane >> 2), c0 = (lane & 3) * 2;
#pragma unroll
    for (int dt = 0; dt < kD / 8; ++dt) {
      *reinterpret_cast<uint32_t*>(so + r0 * kTaRowBytes + (dt * 8 + c0) * 2) = pack_bf16(o[dt][0], o[dt][1]);
      *reinterpret_cast<uint32_t*>(so + (r0 + 8) * kTaRowBytes + (dt * 8 + c0) * 2) = pack_bf16(o[dt][2], o[dt][3]);
    }
  }
  // ---- O -> global in 16-byte row pieces
  __syncwarp();
  {
    const int fr = lane >> 3;
#pragma unroll
    for (int i = 0; i < 4 * kMT; ++i) {
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
  const int mt = p->F <= 16 ? 1 : 2;
  const int warps = hd == 64 ? (mt == 1 ? ta_warps<64, 1>() : ta_warps<64, 2>()) : (mt == 1 ? ta_warps<128, 1>() : ta_warps<128, 2>());
  const long long grid = (items + warps - 1) / warps;
  if (grid > 0x7fffffffLL) return fail(TTVDM_ERR_SHAPE, "attn_temporal: too many items");
  auto* kq = static_cast<const __nv_bfloat16*>(p->q);
  auto* kk = static_cast<const __nv_bfloat16*>(p->k);
  auto* kv = static_cast<const __nv_bfloat16*>(p->v);
  auto* ko = static_cast<__nv_bfloat16*>(p->out);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
#define TA_LAUNCH(D, MT)                                                                                                 \
  attn_temporal_kernel<D, MT><<<(int)grid, warps * 32, 0, st>>>(kq, kk, kv, ko, p->ldq, p->ldk, p->ldv, p->ldo, p->B, p->F, \
                                                                p->S, p->heads, p->scale)
  if (hd == 64 && mt == 1) TA_LAUNCH(64, 1);
  else if (hd == 64) TA_LAUNCH(64, 2);
  else if (mt == 1) TA_LAUNCH(128, 1);
  else TA_LAUNCH(128, 2);
#undef TA_LAUNCH
  TTVDM_CHECK_LAUNCH("attn_temporal_kernel");
  return 0;
}
