// Small HBM-bound kernels of the conditioning builder (encode_clip,
// svd/pipeline_stable_video_diffusion_controlnet.py:130-188): the CLIP towers' MLP activation and the joint
// LayerNorm((78, 1024)) over the concatenated text + image embedding. The towers' linears and attention GEMMs run on
// ttvdm_gemm, their LayerNorms on ttvdm_layernorm, their softmax on ttvdm_softmax_rows.
#include "common.h"
#include "ptx.cuh"

namespace ttvdm {

// kind 2: exact GELU (erf), kind 3: quick GELU x * sigmoid(1.702 x) — the two `hidden_act`s CLIP checkpoints use
__global__ void act_inplace_kernel(__nv_bfloat16* __restrict__ x, long long n, int kind) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = __bfloat162float(x[i]);
    const float r = kind == 2 ? 0.5f * v * (1.f + erff(v * 0.70710678118654752f)) : v / (1.f + __expf(-1.702f * v));
    x[i] = __float2bfloat16(r);
  }
}

// One CTA per row: out = (x - mean) / sqrt(var + eps) over ALL n elements of the row (no affine: the reference builds a
// fresh nn.LayerNorm((78, 1024)) on every call, weight 1 / bias 0). Two passes over a row that sits in L2 (n = 79872).
__global__ void __launch_bounds__(1024) layernorm_flat_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                              long long n, float eps) {
  __shared__ double red[2][32];
  const float* xr = x + (long long)blockIdx.x * n;
  float* orow = out + (long long)blockIdx.x * n;
  double s = 0.0, q = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = (double)__ldg(xr + i);
    s += v;
    q += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = s;
    red[1][warp] = q;
  }
  __syncthreads();
  double ts = 0.0, tq = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
    ts += red[0][i];
    tq += red[1][i];
  }
  const double mean = ts / (double)n;
  double var = tq / (double)n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float m = (float)mean, rstd = (float)(1.0 / sqrt(var + (double)eps));
  for (long long i = threadIdx.x; i < n; i += blockDim.x) orow[i] = (__ldg(xr + i) - m) * rstd;
}

}  // namespace ttvdm

using namespace ttvdm;

extern "C" int ttvdm_act_inplace(void* x, size_t n, int kind, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!x || n == 0 || (kind != 2 && kind != 3)) return fail(TTVDM_ERR_SHAPE, "act_inplace: n=%zu kind=%d", n, kind);
  long long g = ((long long)n + 255) / 256;
  const long long cap = (long long)g_num_sms * 16;
  if (g > cap) g = cap;
  act_inplace_kernel<<<(int)g, 256, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<__nv_bfloat16*>(x), (long long)n,
                                                                          kind);
  TTVDM_CHECK_LAUNCH("act_inplace_kernel");
  return 0;
}

extern "C" int ttvdm_layernorm_flat(const float* x, float* out, int rows, size_t n, float eps, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!x || !out || rows <= 0 || n == 0) return fail(TTVDM_ERR_SHAPE, "layernorm_flat: rows=%d n=%zu", rows, n);
  layernorm_flat_kernel<<<rows, 1024, 0, static_cast<cudaStream_t>(stream_)>>>(x, out, (long long)n, eps);
  TTVDM_CHECK_LAUNCH("layernorm_flat_kernel");
  return 0;
}
