// K7 — temporal self-attention: F (<= 16) frames x F frames per (b, s, head), head_dim 64. ~0.05 % of the FLOPs
// and HBM-bound, so: shared-memory K/V staging, one query frame per thread, no tensor cores. Rows stay in the
// (b, f, s) token order — frames are walked with stride S*ld instead of permuting the activation.
// K/V are widened to fp32 while they are staged, so the inner loops are one 16-byte broadcast LDS per four FMAs
// (the bf16 staging of round 1 spent two ALU unpack instructions per FMA pair and ran at 1.5 TB/s, issue-bound).
#include "common.h"
#include "ptx.cuh"

namespace ttvdm {

#ifndef TTVDM_TA_ITEMS
#define TTVDM_TA_ITEMS 8
#endif
constexpr int kTaItemsPerCta = TTVDM_TA_ITEMS;  // 2 items per warp
constexpr int kTaMaxF = 16;

__global__ void __launch_bounds__(kTaItemsPerCta * 16)
attn_temporal_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                     const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ out, int ldq, int ldk, int ldv,
                     int ldo, int B, int F, int S, int heads, float scale) {
  extern __shared__ __align__(16) float4 ta_smem[];
  float4 (*sk)[kTaMaxF][16] = reinterpret_cast<float4 (*)[kTaMaxF][16]>(ta_smem);                   // [item][frame][d/4]
  float4 (*sv)[kTaMaxF][16] = sk + kTaItemsPerCta;
  const int slot = threadIdx.x >> 4;  // item slot within the CTA
  const int l16 = threadIdx.x & 15;
  const long long items = (long long)B * S * heads;
  const long long item = (long long)blockIdx.x * kTaItemsPerCta + slot;
  const bool item_ok = item < items;
  int head = 0, s = 0, b = 0;
  if (item_ok) {
    head = (int)(item % heads);
    const long long bs = item / heads;
    s = (int)(bs % S);
    b = (int)(bs / S);
  }
  const long long row0 = ((long long)b * F) * S + s;  // frame 0 row; frame f is row0 + f*S
  // the thread's own query row is requested together with K/V (one exposed memory latency instead of two)
  const long long qrow = row0 + (long long)l16 * S;
  uint4 qraw[8];
  if (item_ok && l16 < F) {
    const uint4* qp = reinterpret_cast<const uint4*>(q + qrow * ldq + head * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) qraw[i] = __ldg(qp + i);
  }
  if (item_ok) {
    // all 2*F row loads are issued before the first shared-memory store (in-order issue would otherwise expose one
    // full memory latency per frame)
    uint2 kk[kTaMaxF], vv[kTaMaxF];
#pragma unroll
    for (int f = 0; f < kTaMaxF; ++f) {
      if (f < F) {
        const long long row = row0 + (long long)f * S;
        kk[f] = __ldg(reinterpret_cast<const uint2*>(k + row * ldk + head * 64) + l16);
        vv[f] = __ldg(reinterpret_cast<const uint2*>(v + row * ldv + head * 64) + l16);
      }
    }
#pragma unroll
    for (int f = 0; f < kTaMaxF; ++f) {
      if (f < F) {
        const float2 k0 = unpack_bf16(kk[f].x), k1 = unpack_bf16(kk[f].y);
        const float2 v0 = unpack_bf16(vv[f].x), v1 = unpack_bf16(vv[f].y);
        sk[slot][f][l16] = make_float4(k0.x, k0.y, k1.x, k1.y);
        sv[slot][f][l16] = make_float4(v0.x, v0.y, v1.x, v1.y);
      }
    }
  }
  __syncthreads();
  if (!item_ok || l16 >= F) return;
  float qf[64];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t w4[4] = {qraw[i].x, qraw[i].y, qraw[i].z, qraw[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f2 = unpack_bf16(w4[j]);
      qf[i * 8 + j * 2] = f2.x * scale;
      qf[i * 8 + j * 2 + 1] = f2.y * scale;
    }
  }
  float sc[kTaMaxF];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kTaMaxF; ++j) {
    if (j < F) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const float4 kk = sk[slot][j][d];
        a0 = fmaf(qf[4 * d], kk.x, a0);
        a1 = fmaf(qf[4 * d + 1], kk.y, a1);
        a0 = fmaf(qf[4 * d + 2], kk.z, a0);
        a1 = fmaf(qf[4 * d + 3], kk.w, a1);
      }
      const float acc = a0 + a1;
      sc[j] = acc;
      mx = fmaxf(mx, acc);
    } else {
      sc[j] = -INFINITY;
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kTaMaxF; ++j) {
    sc[j] = (j < F) ? __expf(sc[j] - mx) : 0.f;
    sum += sc[j];
  }
  const float inv = 1.f / sum;
  float o[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < kTaMaxF; ++j) {
    if (j < F) {
      const float pj = sc[j] * inv;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const float4 vv = sv[slot][j][d];
        o[4 * d] = fmaf(pj, vv.x, o[4 * d]);
        o[4 * d + 1] = fmaf(pj, vv.y, o[4 * d + 1]);
        o[4 * d + 2] = fmaf(pj, vv.z, o[4 * d + 2]);
        o[4 * d + 3] = fmaf(pj, vv.w, o[4 * d + 3]);
      }
    }
  }
  uint4* op = reinterpret_cast<uint4*>(out + qrow * ldo + head * 64);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    op[i] = make_uint4(pack_bf16(o[i * 8], o[i * 8 + 1]), pack_bf16(o[i * 8 + 2], o[i * 8 + 3]),
                       pack_bf16(o[i * 8 + 4], o[i * 8 + 5]), pack_bf16(o[i * 8 + 6], o[i * 8 + 7]));
}

}  // namespace ttvdm

using namespace ttvdm;

extern "C" int ttvdm_attn_temporal(const ttvdm_tattn_params* p, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!p || !p->q || !p->k || !p->v || !p->out) return fail(TTVDM_ERR_SHAPE, "attn_temporal: null");
  if (p->F < 1 || p->F > kTaMaxF) return fail(TTVDM_ERR_SHAPE, "attn_temporal: F=%d (max %d)", p->F, kTaMaxF);
  if ((p->ldq | p->ldk | p->ldv | p->ldo) % 8 != 0) return fail(TTVDM_ERR_SHAPE, "attn_temporal: ld %% 8 != 0");
  const long long items = (long long)p->B * p->S * p->heads;
  if (items <= 0) return fail(TTVDM_ERR_SHAPE, "attn_temporal: empty");
  const int grid = (int)((items + kTaItemsPerCta - 1) / kTaItemsPerCta);
  constexpr int smem = 2 * kTaItemsPerCta * kTaMaxF * 16 * sizeof(float4);  // fp32 K and V of the CTA's items (8 KB per item)
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_temporal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(TTVDM_ERR_CUDA, "attn_temporal: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  attn_temporal_kernel<<<grid, kTaItemsPerCta * 16, smem, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(p->q), static_cast<const __nv_bfloat16*>(p->k),
      static_cast<const __nv_bfloat16*>(p->v), static_cast<__nv_bfloat16*>(p->out), p->ldq, p->ldk, p->ldv, p->ldo,
      p->B, p->F, p->S, p->heads, p->scale);
  TTVDM_CHECK_LAUNCH("attn_temporal_kernel");
  return 0;
}
