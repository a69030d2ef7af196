// Weight repack entry points (run once per model load, SURVEY.md §8b "Weight repack entry points"): PyTorch-layout
// parameters (fp32 / fp16 / bf16, straight out of a diffusers state dict) -> the K-major bf16 operands, fp32 epilogue
// vectors and LayerNorm-fold side tables the GEMM family consumes. One launch per parameter; nothing here is on the
// per-step path.
#include "common.h"
#include "ptx.cuh"

#include <cuda_fp16.h>

namespace ttvdm {

__device__ __forceinline__ float load_src(const void* p, int dtype, size_t i) {
  if (dtype == TTVDM_DT_F32) return static_cast<const float*>(p)[i];
  if (dtype == TTVDM_DT_F16) return __half2float(static_cast<const __half*>(p)[i]);
  return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
}

// w [cout, cin, th, tw] -> out [cout, th*tw*cin_pad], tap-major then channel, zero padded channels
__global__ void pack_conv_weight_kernel(const void* __restrict__ w, int dtype, int cout, int cin, int taps, int cin_pad,
                                        __nv_bfloat16* __restrict__ out) {
  const size_t total = (size_t)cout * taps * cin_pad;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cin_pad);
    const size_t r = i / cin_pad;
    const int t = (int)(r % taps);
    const int o = (int)(r / taps);
    float v = 0.f;
    if (c < cin) v = load_src(w, dtype, ((size_t)o * cin + c) * taps + t);
    out[i] = __float2bfloat16(v);
  }
}

// One CTA per source row n: out_w[row(n), k] = bf16(w[n, k] * gamma[k]); colsum[row(n)] = sum_k float(out_w);
// out_bias[row(n)] = bias[n] + sum_k w[n, k] * beta[k]   (fp32; the un-rounded weight, like the reference's fp32 LayerNorm
// shift pushed through the linear).
__global__ void __launch_bounds__(256)
pack_linear_kernel(const void* __restrict__ w, const void* __restrict__ bias, const void* __restrict__ gamma,
                   const void* __restrict__ beta, int dtype, int N, int K, int geglu, int out_row0, int ldo,
                   __nv_bfloat16* __restrict__ out_w, float* __restrict__ out_bias, float* __restrict__ out_colsum) {
  __shared__ float red[2][8];
  const int n = blockIdx.x;
  const int half = N >> 1;
  const int row = out_row0 + (geglu ? (n < half ? 2 * n : 2 * (n - half) + 1) : n);
  float cs = 0.f, bs = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float v = load_src(w, dtype, (size_t)n * K + k);
    const float gk = gamma != nullptr ? load_src(gamma, dtype, k) : 1.f;
    const __nv_bfloat16 q = __float2bfloat16(v * gk);
    out_w[(size_t)row * ldo + k] = q;
    cs += __bfloat162float(q);
    if (beta != nullptr) bs = fmaf(v, load_src(beta, dtype, k), bs);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cs += __shfl_xor_sync(0xffffffffu, cs, o);
    bs += __shfl_xor_sync(0xffffffffu, bs, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = cs;
    red[1][threadIdx.x >> 5] = bs;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      a += red[0][i];
      b += red[1][i];
    }
    if (out_colsum != nullptr) out_colsum[row] = a;
    if (out_bias != nullptr) out_bias[row] = b + (bias != nullptr ? load_src(bias, dtype, n) : 0.f);
  }
}

__global__ void pack_vector_kernel(const void* __restrict__ src, int dtype, size_t n, float* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = load_src(src, dtype, i);
}

static bool dtype_ok(int d) { return d == TTVDM_DT_F32 || d == TTVDM_DT_F16 || d == TTVDM_DT_BF16; }

}  // namespace ttvdm

using namespace ttvdm;

extern "C" int ttvdm_pack_conv_weight(const void* w, int src_dtype, int cout, int cin, int taps, int cin_pad, void* out,
                                      void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!w || !out || cout <= 0 || cin <= 0 || taps <= 0 || !dtype_ok(src_dtype))
    return fail(TTVDM_ERR_SHAPE, "pack_conv_weight: bad arguments");
  if (cin_pad < cin) cin_pad = cin;
  const size_t total = (size_t)cout * taps * cin_pad;
  const int threads = 256;
  const int grid = (int)((total + threads - 1) / threads < 4096 ? (total + threads - 1) / threads : 4096);
  pack_conv_weight_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream_)>>>(w, src_dtype, cout, cin, taps, cin_pad,
                                                                                  static_cast<__nv_bfloat16*>(out));
  TTVDM_CHECK_LAUNCH("pack_conv_weight_kernel");
  return 0;
}

extern "C" int ttvdm_pack_linear(const ttvdm_pack_linear_params* p, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!p || !p->w || !p->out_w || p->N <= 0 || p->K <= 0 || !dtype_ok(p->src_dtype))
    return fail(TTVDM_ERR_SHAPE, "pack_linear: bad arguments");
  if (p->geglu && (p->N & 1)) return fail(TTVDM_ERR_SHAPE, "pack_linear: GEGLU interleave needs an even N");
  if ((p->beta || p->bias) && !p->out_bias) return fail(TTVDM_ERR_SHAPE, "pack_linear: out_bias missing");
  const int ldo = p->ldo > 0 ? p->ldo : p->K;
  pack_linear_kernel<<<p->N, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      p->w, p->bias, p->gamma, p->beta, p->src_dtype, p->N, p->K, p->geglu, p->out_row0, ldo,
      static_cast<__nv_bfloat16*>(p->out_w), p->out_bias, p->out_colsum);
  TTVDM_CHECK_LAUNCH("pack_linear_kernel");
  return 0;
}

extern "C" int ttvdm_pack_vector(const void* src, int src_dtype, size_t n, float* out, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!src || !out || n == 0 || !dtype_ok(src_dtype)) return fail(TTVDM_ERR_SHAPE, "pack_vector: bad arguments");
  const int threads = 256;
  const int grid = (int)((n + threads - 1) / threads < 1024 ? (n + threads - 1) / threads : 1024);
  pack_vector_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream_)>>>(src, src_dtype, n, out);
  TTVDM_CHECK_LAUNCH("pack_vector_kernel");
  return 0;
}

// ---- workspace queries ("no allocation inside ops; ttvdm_<op>_workspace_bytes() queries it")
extern "C" size_t ttvdm_groupnorm_workspace_bytes(int rows, int rows_per_inst) {
  if (rows <= 0 || rows_per_inst <= 0) return 0;
  return (size_t)(rows / rows_per_inst) * 64 * sizeof(double);
}
extern "C" size_t ttvdm_gemm_gn_stats_bytes(int M, int N, int gn_rows_per_inst) {
  if (M <= 0 || N <= 0 || gn_rows_per_inst <= 0) return 0;
  return (size_t)(M / gn_rows_per_inst) * (size_t)N * sizeof(double);
}
extern "C" size_t ttvdm_gemm_row_sums_bytes(int M, int N) {
  return (M > 0 && N > 0) ? (size_t)(N / 32) * (size_t)M * 2 * sizeof(float) : 0;
}
extern "C" size_t ttvdm_gemm_workspace_bytes(const ttvdm_gemm_params*) { return 0; }
extern "C" size_t ttvdm_attn_workspace_bytes(const ttvdm_attn_params*) { return 0; }
extern "C" size_t ttvdm_gesture_scratch_bytes(int n_points, int H, int W) {
  return (size_t)(n_points > 0 ? n_points : 1) * (size_t)(H + W) * sizeof(float);
}
