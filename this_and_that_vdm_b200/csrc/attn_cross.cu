// K6 / K8 — cross-attention against the short CLIP context (L <= 128 keys, head_dim 64), spatial and temporal.
//
// HBM-bound: per query row and head 128 B in, 128 B out, against ~20 kFLOP. K/V of one (context, head) are tiny
// (L x 128 B each), so a CTA pins them in shared memory and its 8 warps stream 16-row query tiles through warp-level
// mma.sync (m16n8k16, ldmatrix operands): S = Q K^T (16 x L) stays in accumulator registers, the softmax runs in the
// quad that owns each row, P is re-used as the A fragments of O = P V, and O leaves through shared memory in 16-byte
// row pieces. Round 1 ran these calls through the tcgen05 flash kernel (one 128-key KV tile per 128-row Q tile): 40 %
// of the exponentials and MMAs were padding, the temporal variant visited every context tile for every row, and the
// kernel sat at 1.0-1.6 TB/s; tcgen05 buys nothing when the whole K/V is one small tile.
//
// Which rows read which context (include/ttvdm.h, ttvdm_xattn_params):
//   spatial  (BasicTransformerBlock.attn2): row (b, f, s) reads context b_global.
//   temporal (TemporalBasicTransformerBlock.attn2): the reference's quirk — the temporal batch row (b, s) reads
//            context (b_global * S + s) mod n_ctx (svd/diffusion_arch/transformer_temporal.py:310-319). Here a CTA
//            serves ONE context and gathers exactly the rows that read it (s = s0(b) + n_ctx * i), so nothing is masked.
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace ttvdm {

constexpr int kXaWarps = 8;
// rows of kD * 2 B of data + 16 B pad (conflict-free ldmatrix); kD = 64, or 128 for the reference UNet's class-default heads
constexpr int kXaMaxL = 128;

__device__ __forceinline__ void xa_cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void xa_ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void xa_ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void xa_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct XaArgs {
  const __nv_bfloat16* q;
  const __nv_bfloat16* kc;
  const __nv_bfloat16* vc;
  __nv_bfloat16* out;
  int ldq, ldo, heads, L, F, S, n_ctx, b_local, batch_offset, temporal;
  int tiles_per_unit;  // 16-row tiles per (b, f) unit (upper bound over the units a CTA serves)
  int chunks;          // CTAs per (context, head)
  float scale_log2;    // scale * log2(e)
};

// kPairs = ceil(L / 16): 16-key steps (pairs of 8-key n-tiles) that hold a valid key
template <int kPairs, int kD = 64>
__global__ void __launch_bounds__(kXaWarps * 32)
attn_cross_kernel(const XaArgs g) {
  constexpr int kXaRowBytes = kD * 2 + 16;
  constexpr int kCh = kD / 8;  // 16-byte chunks per row
  extern __shared__ __align__(16) uint8_t xa_smem[];
  uint8_t* sK = xa_smem;                               // [kPairs * 16][144 B]
  uint8_t* sV = sK + kPairs * 16 * kXaRowBytes;
  uint8_t* sQ = sV + kPairs * 16 * kXaRowBytes;        // [kXaWarps][16][144 B] (Q, later O)
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int ctx = blockIdx.z;
  // units (b_local, f) this CTA serves, and the first / step of s inside a unit
  int unit0, n_units;
  if (g.temporal) {
    unit0 = 0;
    n_units = g.b_local * g.F;
  } else {
    const int bl = ctx - g.batch_offset;  // spatial: context = global batch index
    if (bl < 0 || bl >= g.b_local) return;
    unit0 = bl * g.F;
    n_units = g.F;
  }
  const int stride = g.temporal ? g.n_ctx : 1;

  // ---- K, V of (ctx, head) -> shared memory, rows >= L zero-filled (P is 0 there, but 0 * garbage must stay 0)
  {
    const __nv_bfloat16* kb = g.kc + ((size_t)ctx * g.L) * (g.heads * kD) + head * kD;
    const __nv_bfloat16* vb = g.vc + ((size_t)ctx * g.L) * (g.heads * kD) + head * kD;
    for (int i = threadIdx.x; i < kPairs * 16 * kCh; i += kXaWarps * 32) {
      const int row = i / kCh, ch = i % kCh;
      const int ok = row < g.L ? 16 : 0;
      const size_t off = (size_t)(row < g.L ? row : 0) * (g.heads * kD) + ch * 8;
      xa_cp_async16(smem_u32(sK) + row * kXaRowBytes + ch * 16, kb + off, ok);
      xa_cp_async16(smem_u32(sV) + row * kXaRowBytes + ch * 16, vb + off, ok);
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }

  const uint32_t sk = smem_u32(sK), sv = smem_u32(sV), sq = smem_u32(sQ + warp * 16 * kXaRowBytes);
  uint8_t* const sq_gen = sQ + warp * 16 * kXaRowBytes;
  const int items = n_units * g.tiles_per_unit;
  const int per_cta = (items + g.chunks - 1) / g.chunks;
  const int it_end = min(items, ((int)blockIdx.x + 1) * per_cta);
  for (int it = (int)blockIdx.x * per_cta + warp; it < it_end; it += kXaWarps) {
    const int unit = unit0 + it / g.tiles_per_unit;  // local (b, f)
    const int tt = it - (it / g.tiles_per_unit) * g.tiles_per_unit;
    int s0 = 0;
    if (g.temporal) {
      const long long bg = (long long)(g.batch_offset + unit / g.F) * g.S;  // b_global * S
      s0 = (int)(((ctx - bg) % g.n_ctx + g.n_ctx) % g.n_ctx);              // first s with (b_global * S + s) mod n_ctx == ctx
    }
    const long long unit_row0 = (long long)unit * g.S;
    // tile row i <-> s = s0 + stride * (16 * tt + i); rows with s >= S do not exist
    {
      const int fr = lane >> 3;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i * 4 + fr;
        const int s = s0 + stride * (16 * tt + r);
        const int ok = s < g.S ? 16 : 0;
#pragma unroll
        for (int c8 = 0; c8 < kCh; c8 += 8) {
          const int ch = c8 + (lane & 7);
          xa_cp_async16(sq + r * kXaRowBytes + ch * 16, g.q + (unit_row0 + (s < g.S ? s : 0)) * g.ldq + head * kD + ch * 8, ok);
        }
      }
      asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
      __syncwarp();
    }
    // ---- S = Q K^T (16 x 16*kPairs, fp32)
    float sc[2 * kPairs][4];
#pragma unroll
    for (int n = 0; n < 2 * kPairs; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) sc[n][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < kD / 16; ++kk) {
      uint32_t a[4];
      xa_ldsm_x4(sq + (lane & 15) * kXaRowBytes + kk * 32 + (lane >> 4) * 16, a);
#pragma unroll
      for (int np = 0; np < kPairs; ++np) {
        uint32_t bm[4];
        xa_ldsm_x4(sk + (np * 16 + (lane >> 4) * 8 + (lane & 7)) * kXaRowBytes + kk * 32 + ((lane >> 3) & 1) * 16, bm);
        xa_mma(sc[2 * np], a, bm[0], bm[1]);
        xa_mma(sc[2 * np + 1], a, bm[2], bm[3]);
      }
    }
    // ---- softmax over the L keys: thread holds rows lane/4 (elements 0, 1) and lane/4 + 8 (elements 2, 3)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < 2 * kPairs; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = n * 8 + (lane & 3) * 2 + (e & 1);
        sc[n][e] = col < g.L ? sc[n][e] * g.scale_log2 : -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], sc[n][e]);
      }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int n = 0; n < 2 * kPairs; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float p;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(sc[n][e] - mx[e >> 1]));  // exp2(-inf) = 0 for padded keys
        sc[n][e] = p;
        sum[e >> 1] += p;
      }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
      sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
    }
    const float inv[2] = {1.f / sum[0], 1.f / sum[1]};
    // ---- O = P V (16 x 64): P (bf16, unnormalised — values in (0, 1]) straight from the accumulator layout
    float o[kD / 8][4];
#pragma unroll
    for (int dt = 0; dt < kD / 8; ++dt)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[dt][e] = 0.f;
#pragma unroll
    for (int kp = 0; kp < kPairs; ++kp) {
      uint32_t pa[4];
      pa[0] = pack_bf16(sc[2 * kp][0], sc[2 * kp][1]);
      pa[1] = pack_bf16(sc[2 * kp][2], sc[2 * kp][3]);
      pa[2] = pack_bf16(sc[2 * kp + 1][0], sc[2 * kp + 1][1]);
      pa[3] = pack_bf16(sc[2 * kp + 1][2], sc[2 * kp + 1][3]);
#pragma unroll
      for (int d2 = 0; d2 < kD / 16; ++d2) {
        uint32_t bm[4];
        xa_ldsm_x4_trans(sv + (kp * 16 + (lane & 15)) * kXaRowBytes + (d2 * 2 + (lane >> 4)) * 16, bm);
        xa_mma(o[d2 * 2], pa, bm[0], bm[1]);
        xa_mma(o[d2 * 2 + 1], pa, bm[2], bm[3]);
      }
    }
    // ---- O / l -> the warp's Q tile (Q is dead) -> global in 16-byte row pieces
    __syncwarp();
    {
      const int r0 = lane >> 2, c0 = (lane & 3) * 2;
#pragma unroll
      for (int dt = 0; dt < kD / 8; ++dt) {
        *reinterpret_cast<uint32_t*>(sq_gen + r0 * kXaRowBytes + (dt * 8 + c0) * 2) =
            pack_bf16(o[dt][0] * inv[0], o[dt][1] * inv[0]);
        *reinterpret_cast<uint32_t*>(sq_gen + (r0 + 8) * kXaRowBytes + (dt * 8 + c0) * 2) =
            pack_bf16(o[dt][2] * inv[1], o[dt][3] * inv[1]);
      }
      __syncwarp();
      const int fr = lane >> 3;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i * 4 + fr;
        const int s = s0 + stride * (16 * tt + r);
        if (s < g.S) {
#pragma unroll
          for (int c8 = 0; c8 < kCh; c8 += 8) {
            const int ch = c8 + (lane & 7);
            const uint4 u = *reinterpret_cast<const uint4*>(sq_gen + r * kXaRowBytes + ch * 16);
            *reinterpret_cast<uint4*>(g.out + (unit_row0 + s) * g.ldo + head * kD + ch * 8) = u;
          }
        }
      }
      __syncwarp();  // the tile is re-filled by the next item's cp.async
    }
  }
}

}  // namespace ttvdm

using namespace ttvdm;

extern "C" int ttvdm_attn_cross(const ttvdm_xattn_params* p, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!p || !p->q || !p->kc || !p->vc || !p->out) return fail(TTVDM_ERR_SHAPE, "attn_cross: null");
  if (p->L <= 0 || p->L > kXaMaxL) return fail(TTVDM_ERR_SHAPE, "attn_cross: L=%d (1..128)", p->L);
  if (p->F <= 0 || p->S <= 0 || p->rows <= 0 || p->rows % (p->F * p->S) != 0)
    return fail(TTVDM_ERR_SHAPE, "attn_cross: rows=%d not a multiple of F*S=%d", p->rows, p->F * p->S);
  if ((p->ldq | p->ldo) % 8 != 0) return fail(TTVDM_ERR_SHAPE, "attn_cross: ld %% 8 != 0");
  if ((reinterpret_cast<uintptr_t>(p->q) | reinterpret_cast<uintptr_t>(p->kc) | reinterpret_cast<uintptr_t>(p->vc) |
       reinterpret_cast<uintptr_t>(p->out)) & 15)
    return fail(TTVDM_ERR_SHAPE, "attn_cross: q/kc/vc/out must be 16-byte aligned");
  const int b_local = p->rows / (p->F * p->S);
  if (p->n_ctx <= 0 || (!p->temporal && p->batch_offset + b_local > p->n_ctx))
    return fail(TTVDM_ERR_SHAPE, "attn_cross: batch %d+%d exceeds n_ctx=%d", p->batch_offset, b_local, p->n_ctx);
  if (p->heads <= 0 || p->heads > 65535 || p->n_ctx > 65535) return fail(TTVDM_ERR_SHAPE, "attn_cross: grid too large");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  {
    // tcgen05 / TMA path (the flash kernel's cross mode) for context strides <= 2; TTVDM_XATTN_LEGACY=1 keeps the
    // warp-level mma.sync kernel below for A/B runs, and it remains the path for n_ctx > 2 temporal calls
    static const int legacy = getenv("TTVDM_XATTN_LEGACY") ? atoi(getenv("TTVDM_XATTN_LEGACY")) : 0;
    if (!legacy && (p->head_dim == 0 || p->head_dim == 64)) {
      const int rc = launch_attn_cross_tc(p, stream);
      if (rc >= 0) return rc;
    }
  }
  XaArgs g;
  g.q = static_cast<const __nv_bfloat16*>(p->q);
  g.kc = static_cast<const __nv_bfloat16*>(p->kc);
  g.vc = static_cast<const __nv_bfloat16*>(p->vc);
  g.out = static_cast<__nv_bfloat16*>(p->out);
  g.ldq = p->ldq;
  g.ldo = p->ldo;
  g.heads = p->heads;
  g.L = p->L;
  g.F = p->F;
  g.S = p->S;
  g.n_ctx = p->n_ctx;
  g.b_local = b_local;
  g.batch_offset = p->batch_offset;
  g.temporal = p->temporal ? 1 : 0;
  g.scale_log2 = p->scale * 1.4426950408889634f;
  const int stride = g.temporal ? p->n_ctx : 1;
  const int rows_per_unit = (p->S + stride - 1) / stride;  // upper bound of the rows of a unit that read one context
  g.tiles_per_unit = (rows_per_unit + 15) / 16;
  const int n_units = g.temporal ? b_local * p->F : p->F;
  const int items = n_units * g.tiles_per_unit;  // per (context, head)
  // enough CTAs per (context, head) for ~4 CTAs per SM over the whole grid, at least 2 tiles per warp
  const int pairs_live = p->heads * (g.temporal ? p->n_ctx : b_local);
  int chunks = (4 * g_num_sms + pairs_live - 1) / pairs_live;
  const int max_chunks = (items + 2 * kXaWarps - 1) / (2 * kXaWarps);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  g.chunks = chunks;
  const int pairs = (p->L + 15) / 16;
  const int hd = p->head_dim == 0 ? 64 : p->head_dim;
  if (hd != 64 && hd != 128) return fail(TTVDM_ERR_SHAPE, "attn_cross: head_dim %d (64 or 128)", hd);
  const size_t smem = (size_t)(2 * pairs * 16 + kXaWarps * 16) * (hd * 2 + 16);
  dim3 grid(chunks, p->heads, p->n_ctx);
#define XA_LAUNCH(P)                                                                                              \
  do {                                                                                                            \
    static bool attr_set[2] = {false, false};                                                                     \
    if (!attr_set[hd == 128]) {                                                                                   \
      cudaError_t e = hd == 64 ? cudaFuncSetAttribute(attn_cross_kernel<P, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                                      (2 * P * 16 + kXaWarps * 16) * 144)                         \
                               : cudaFuncSetAttribute(attn_cross_kernel<P, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                                      (2 * P * 16 + kXaWarps * 16) * 272);                        \
      if (e != cudaSuccess) return fail(TTVDM_ERR_CUDA, "attn_cross: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); \
      attr_set[hd == 128] = true;                                                                                 \
    }                                                                                                             \
    if (hd == 64) attn_cross_kernel<P, 64><<<grid, kXaWarps * 32, smem, stream>>>(g);                             \
    else attn_cross_kernel<P, 128><<<grid, kXaWarps * 32, smem, stream>>>(g);                                     \
  } while (0)
  switch (pairs) {
    case 1: XA_LAUNCH(1); break;
    case 2: XA_LAUNCH(2); break;
    case 3: XA_LAUNCH(3); break;
    case 4: XA_LAUNCH(4); break;
    case 5: XA_LAUNCH(5); break;
    case 6: XA_LAUNCH(6); break;
    case 7: XA_LAUNCH(7); break;
    default: XA_LAUNCH(8); break;
  }
#undef XA_LAUNCH
  TTVDM_CHECK_LAUNCH("attn_cross_kernel");
  return 0;
}
