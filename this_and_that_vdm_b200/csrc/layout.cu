// Layout helpers (stride-2 patch gather, nearest 2x upsample, residual axpy) and the fused sampler glue
// (CFG duplicate + sigma scaling + channel concat; CFG combine + Euler v-prediction step). All HBM-bound,
// 16-byte vectorised, grid-stride.
#include "common.h"
#include "ptx.cuh"

namespace ttvdm {

__global__ void im2col_s2_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int n_img, int H, int W,
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
    const int h = 2 * ho + tap / 3 - 1, w = 2 * wo + tap % 3 - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (h >= 0 && h < H && w >= 0 && w < W) v = __ldg(x + (((long long)n * H + h) * W + w) * C8 + c);
    out[i] = v;
  }
}

__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int n_img, int H, int W,
                                  int C8) {
  const long long total = (long long)n_img * H * 2 * W * 2 * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long r = i / C8;
    const int wo = (int)(r % (2 * W));
    r /= (2 * W);
    const int ho = (int)(r % (2 * H));
    const int n = (int)(r / (2 * H));
    out[i] = __ldg(x + (((long long)n * H + (ho >> 1)) * W + (wo >> 1)) * C8 + c);
  }
}

// out[n, 2y+py, 2x+px, :] = parts[2*py+px][n, y, x, :]  (16-byte pieces; the write side is the contiguous one)
__global__ void interleave2x_kernel(const uint4* __restrict__ parts, uint4* __restrict__ out, int n_img, int H, int W,
                                    int C8) {
  const long long plane = (long long)n_img * H * W * C8;
  const long long total = 4 * plane;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long r = i / C8;
    const int wo = (int)(r % (2 * W));
    r /= (2 * W);
    const int ho = (int)(r % (2 * H));
    const int n = (int)(r / (2 * H));
    const int p = ((ho & 1) << 1) | (wo & 1);
    out[i] = __ldg(parts + p * plane + (((long long)n * H + (ho >> 1)) * W + (wo >> 1)) * C8 + c);
  }
}

// no __restrict__ / __ldg: `out` may alias `a` (the residual merge runs in place)
__global__ void axpy_kernel(const uint4* a, const uint4* b, uint4* out,
                            float scale, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 ua = a[i], ub = b[i];
    const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wb[4] = {ub.x, ub.y, ub.z, ub.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = unpack_bf16(wa[j]), fb = unpack_bf16(wb[j]);
      o[j] = pack_bf16(fa.x + scale * fb.x, fa.y + scale * fb.y);
    }
    out[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// one thread per (b, f, y, x) pixel: writes c_pad bf16 channels
__global__ void sampler_prepare_kernel(const float* __restrict__ latents, const float* __restrict__ image_latents,
                                       const float* __restrict__ cond, __nv_bfloat16* __restrict__ model_in, int c_pad,
                                       int B_local, int batch_offset, int F, int h, int w, float inv_scale) {
  const long long hw = (long long)h * w;
  const long long total = (long long)B_local * F * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % hw;
    const int f = (int)((i / hw) % F);
    const int b = (int)(i / (hw * F));
    float v[12];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      v[c] = latents[((long long)f * 4 + c) * hw + pix] * inv_scale;
      v[4 + c] = image_latents[((long long)(b + batch_offset) * 4 + c) * hw + pix];
      v[8 + c] = cond ? cond[((long long)f * 4 + c) * hw + pix] : 0.f;
    }
    uint32_t* o = reinterpret_cast<uint32_t*>(model_in + i * c_pad);
#pragma unroll
    for (int j = 0; j < 6; ++j) o[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
    for (int j = 6; j < c_pad / 2; ++j) o[j] = 0u;
  }
}

__global__ void euler_step_kernel(float* __restrict__ latents, const float* __restrict__ eps_u,
                                  const float* __restrict__ eps_c, int ld_eps, const float* __restrict__ guidance,
                                  int F, int h, int w, float sigma, float sigma_next) {
  const long long hw = (long long)h * w;
  const long long total = (long long)F * 4 * hw;
  const float c_out = -sigma / sqrtf(sigma * sigma + 1.f);
  const float c_skip = 1.f / (sigma * sigma + 1.f);
  const float dt = sigma_next - sigma;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % hw;
    const int c = (int)((i / hw) % 4);
    const int f = (int)(i / (hw * 4));
    const long long e = ((long long)f * hw + pix) * ld_eps + c;
    const float u = eps_u[e], cc = eps_c[e];
    const float eps = u + guidance[f] * (cc - u);
    const float x = latents[i];
    const float x0 = eps * c_out + x * c_skip;
    const float d = (x - x0) / sigma;
    latents[i] = x + d * dt;
  }
}

__global__ void sinusoid_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out, int n, int dim) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half) return;
  const int row = i / half, j = i - row * half;
  const float w = expf(-9.210340371976184f * (float)j / (float)half);
  const float arg = t[row] * w;
  out[(long long)row * dim + j] = __float2bfloat16(cosf(arg));
  out[(long long)row * dim + half + j] = __float2bfloat16(sinf(arg));
}

static inline int grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  const long long cap = (long long)g_num_sms * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace ttvdm

using namespace ttvdm;

extern "C" int ttvdm_im2col_s2(const void* x, void* out, int n_img, int H, int W, int C, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!x || !out || C % 8 != 0 || H % 2 != 0 || W % 2 != 0 || n_img <= 0)
    return fail(TTVDM_ERR_SHAPE, "im2col_s2: need C%%8==0 and even H,W (got C=%d H=%d W=%d)", C, H, W);
  const long long total = (long long)n_img * (H / 2) * (W / 2) * 9 * (C / 8);
  im2col_s2_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(out), n_img, H, W, C / 8, H / 2, W / 2);
  TTVDM_CHECK_LAUNCH("im2col_s2_kernel");
  return 0;
}

extern "C" int ttvdm_upsample2x(const void* x, void* out, int n_img, int H, int W, int C, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!x || !out || C % 8 != 0 || n_img <= 0) return fail(TTVDM_ERR_SHAPE, "upsample2x: need C%%8==0 (C=%d)", C);
  const long long total = (long long)n_img * H * 2 * W * 2 * (C / 8);
  upsample2x_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(out), n_img, H, W, C / 8);
  TTVDM_CHECK_LAUNCH("upsample2x_kernel");
  return 0;
}

extern "C" int ttvdm_interleave2x(const void* parts, void* out, int n_img, int H, int W, int C, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!parts || !out || C % 8 != 0 || n_img <= 0) return fail(TTVDM_ERR_SHAPE, "interleave2x: need C%%8==0 (C=%d)", C);
  const long long total = (long long)n_img * H * 2 * W * 2 * (C / 8);
  interleave2x_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const uint4*>(parts), static_cast<uint4*>(out), n_img, H, W, C / 8);
  TTVDM_CHECK_LAUNCH("interleave2x_kernel");
  return 0;
}

extern "C" int ttvdm_sinusoid(const float* t, void* out, int n, int dim, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!t || !out || n <= 0 || dim <= 0 || dim % 2 != 0) return fail(TTVDM_ERR_SHAPE, "sinusoid: n=%d dim=%d", n, dim);
  const int total = n * (dim / 2);
  sinusoid_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      t, static_cast<__nv_bfloat16*>(out), n, dim);
  TTVDM_CHECK_LAUNCH("sinusoid_kernel");
  return 0;
}

extern "C" int ttvdm_axpy(const void* a, const void* b, void* out, float scale, size_t n, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!a || !b || !out || n % 8 != 0) return fail(TTVDM_ERR_SHAPE, "axpy: n %% 8 != 0");
  axpy_kernel<<<grid_for((long long)n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const uint4*>(a), static_cast<const uint4*>(b), static_cast<uint4*>(out), scale, (long long)n / 8);
  TTVDM_CHECK_LAUNCH("axpy_kernel");
  return 0;
}

extern "C" int ttvdm_sampler_prepare(const ttvdm_prepare_params* p, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!p || !p->latents || !p->image_latents || !p->model_in) return fail(TTVDM_ERR_SHAPE, "sampler_prepare: null");
  if (p->c_pad < 12 || p->c_pad % 2 != 0) return fail(TTVDM_ERR_SHAPE, "sampler_prepare: c_pad=%d", p->c_pad);
  const long long total = (long long)p->B_local * p->F * p->h * p->w;
  const float inv = 1.0f / sqrtf(p->sigma * p->sigma + 1.0f);
  sampler_prepare_kernel<<<grid_for(total, 128), 128, 0, static_cast<cudaStream_t>(stream_)>>>(
      p->latents, p->image_latents, p->cond, static_cast<__nv_bfloat16*>(p->model_in), p->c_pad, p->B_local,
      p->batch_offset, p->F, p->h, p->w, inv);
  TTVDM_CHECK_LAUNCH("sampler_prepare_kernel");
  return 0;
}

extern "C" int ttvdm_sampler_euler_step(const ttvdm_euler_params* p, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!p || !p->latents || !p->eps_u || !p->eps_c || !p->guidance) return fail(TTVDM_ERR_SHAPE, "euler: null");
  if (p->sigma <= 0.f) return fail(TTVDM_ERR_SHAPE, "euler: sigma must be > 0");
  const long long total = (long long)p->F * 4 * p->h * p->w;
  euler_step_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      p->latents, p->eps_u, p->eps_c, p->ld_eps, p->guidance, p->F, p->h, p->w, p->sigma, p->sigma_next);
  TTVDM_CHECK_LAUNCH("euler_step_kernel");
  return 0;
}
