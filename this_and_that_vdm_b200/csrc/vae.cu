// HBM-bound kernels that only the VAE (AutoencoderKLTemporalDecoder) either side of the denoising loop needs:
//   * softmax_rows      — the single-head (d = 512) attention of the VAE mid blocks is run as two tcgen05 GEMMs
//                         (scores = Q K^T, out = P V) around this row softmax (fp32 scores in, bf16 probabilities out)
//   * im2col_s2_pad01   — Downsample2D(padding=0) of the encoder: F.pad(x, (0,1,0,1)) + Conv2d(3, stride 2)
//   * vae_time_conv_out — the decoder's last layer, Conv3d(3, 3, (3,1,1)) over frames, fused with the
//                         channels-last -> NCHW fp32 conversion of the decoded frames
#include "common.h"
#include "ptx.cuh"

namespace ttvdm {

constexpr int kSoftmaxThreads = 512;

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // red[] may still be read from the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  float r = red[0];
  for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

// One CTA per row. CACHED: the row is staged in shared memory once (one HBM read per element); otherwise it is
// re-read (rows longer than the shared-memory budget). VEC: 16-byte global loads / 8-byte stores (rows and strides
// that are multiples of 4 elements on 16-byte aligned bases; the launcher checks).
template <bool CACHED, bool VEC>
__global__ void __launch_bounds__(kSoftmaxThreads)
softmax_rows_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ out, long long ldo, int cols_all,
                    int cols_out, int causal) {
  extern __shared__ __align__(16) float row[];
  // causal > 0 (CLIP text tower): rows come in blocks of `causal` queries (one block per head); query i = row % causal
  // attends to keys 0..i; the masked probabilities are written as zeros
  const int cols = causal > 0 ? min(cols_all, (int)(blockIdx.x % causal) + 1) : cols_all;
  __shared__ float red[kSoftmaxThreads / 32];
  const float* xr = x + (long long)blockIdx.x * ldx;
  __nv_bfloat16* orow = out + (long long)blockIdx.x * ldo;
  float m = -INFINITY;
  if (VEC) {  // CACHED is implied; cols == cols_all is a multiple of 4
    for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(xr + c));
      *reinterpret_cast<float4*>(row + c) = v;
      m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
  } else {
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
      const float v = __ldg(xr + c);
      if (CACHED) row[c] = v;
      m = fmaxf(m, v);
    }
  }
  m = block_reduce(m, true, red);
  constexpr float kLog2e = 1.4426950408889634f;
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float e = exp2f(((CACHED ? row[c] : __ldg(xr + c)) - m) * kLog2e);
    if (CACHED) row[c] = e;
    s += e;
  }
  s = block_reduce(s, false, red);
  const float inv = 1.f / s;
  if (VEC) {  // cols_out is a multiple of 4 as well
    for (int c = threadIdx.x * 4; c < cols_out; c += blockDim.x * 4) {
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < cols) p = *reinterpret_cast<const float4*>(row + c);
      *reinterpret_cast<uint2*>(orow + c) = make_uint2(pack_bf16(p.x * inv, p.y * inv), pack_bf16(p.z * inv, p.w * inv));
    }
  } else {
    for (int c = threadIdx.x; c < cols_out; c += blockDim.x) {
      float p = 0.f;
      if (c < cols) p = (CACHED ? row[c] : exp2f((__ldg(xr + c) - m) * kLog2e)) * inv;
      orow[c] = __float2bfloat16(p);
    }
  }
}

__global__ void im2col_s2_pad01_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int n_img, int H, int W,
                                       int C8, int Ho, int Wo) {
  const long long total = (long long)n_img * Ho * Wo * 9 * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long r = i / C8;
    const int tap = (int)(r % 9);
    r /= 9;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const int h = 2 * ho + tap / 3, w = 2 * wo + tap % 3;  // zero padding on the bottom / right edge only
    uint4 v = make_uint4(0, 0, 0, 0);
    if (h < H && w < W) v = __ldg(x + (((long long)n * H + h) * W + w) * C8 + c);
    out[i] = v;
  }
}

struct TimeConvW {
  float w[27];  // [co][ci][t]
  float b[3];
};

__global__ void vae_time_conv_out_kernel(const float* __restrict__ x, int ldx, TimeConvW k, float* __restrict__ out,
                                         int B, int F, long long S) {
  const long long total = (long long)B * F * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long s = i % S;
    const long long bf = i / S;
    const int f = (int)(bf % F);
    float acc[3] = {k.b[0], k.b[1], k.b[2]};
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const int ff = f + t - 1;
      if (ff < 0 || ff >= F) continue;
      const float* xp = x + (i + (long long)(t - 1) * S) * ldx;
      const float v0 = __ldg(xp), v1 = __ldg(xp + 1), v2 = __ldg(xp + 2);
#pragma unroll
      for (int co = 0; co < 3; ++co)
        acc[co] += k.w[co * 9 + t] * v0 + k.w[co * 9 + 3 + t] * v1 + k.w[co * 9 + 6 + t] * v2;
    }
#pragma unroll
    for (int co = 0; co < 3; ++co) out[(bf * 3 + co) * S + s] = acc[co];
  }
}

static inline int grid_for_vae(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  const long long cap = (long long)g_num_sms * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace ttvdm

using namespace ttvdm;

extern "C" int ttvdm_softmax_rows(const float* x, int ldx, void* out, int ldo, int rows, int cols, int cols_out,
                                  int causal, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!x || !out || rows <= 0 || cols <= 0 || cols_out < cols || ldx < cols || ldo < cols_out)
    return fail(TTVDM_ERR_SHAPE, "softmax_rows: rows=%d cols=%d cols_out=%d ldx=%d ldo=%d", rows, cols, cols_out, ldx, ldo);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t smem = (size_t)cols * sizeof(float);
  constexpr size_t kMaxSmem = 200 * 1024;
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  if (smem <= kMaxSmem) {
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(softmax_rows_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)kMaxSmem);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(softmax_rows_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
      if (e != cudaSuccess) return fail(TTVDM_ERR_CUDA, "softmax_rows: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      attr = true;
    }
    const bool vec = causal == 0 && cols % 4 == 0 && cols_out % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0;
    if (vec)
      softmax_rows_kernel<true, true><<<rows, kSoftmaxThreads, smem, stream>>>(x, ldx, o, ldo, cols, cols_out, causal);
    else
      softmax_rows_kernel<true, false><<<rows, kSoftmaxThreads, smem, stream>>>(x, ldx, o, ldo, cols, cols_out, causal);
  } else {
    softmax_rows_kernel<false, false><<<rows, kSoftmaxThreads, 0, stream>>>(x, ldx, o, ldo, cols, cols_out, causal);
  }
  TTVDM_CHECK_LAUNCH("softmax_rows_kernel");
  return 0;
}

extern "C" int ttvdm_im2col_s2_pad01(const void* x, void* out, int n_img, int H, int W, int C, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!x || !out || C % 8 != 0 || H % 2 != 0 || W % 2 != 0 || n_img <= 0)
    return fail(TTVDM_ERR_SHAPE, "im2col_s2_pad01: need C%%8==0 and even H,W (got C=%d H=%d W=%d)", C, H, W);
  const long long total = (long long)n_img * (H / 2) * (W / 2) * 9 * (C / 8);
  im2col_s2_pad01_kernel<<<grid_for_vae(total, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(out), n_img, H, W, C / 8, H / 2, W / 2);
  TTVDM_CHECK_LAUNCH("im2col_s2_pad01_kernel");
  return 0;
}

extern "C" int ttvdm_vae_time_conv_out(const float* x, int ldx, const float* w_host, const float* bias_host, float* out,
                                       int B, int F, int S, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!x || !w_host || !bias_host || !out || B <= 0 || F <= 0 || S <= 0 || ldx < 3)
    return fail(TTVDM_ERR_SHAPE, "vae_time_conv_out: B=%d F=%d S=%d ldx=%d", B, F, S, ldx);
  TimeConvW k;
  for (int i = 0; i < 27; ++i) k.w[i] = w_host[i];
  for (int i = 0; i < 3; ++i) k.b[i] = bias_host[i];
  const long long total = (long long)B * F * S;
  vae_time_conv_out_kernel<<<grid_for_vae(total, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(x, ldx, k, out, B, F,
                                                                                                 (long long)S);
  TTVDM_CHECK_LAUNCH("vae_time_conv_out_kernel");
  return 0;
}
