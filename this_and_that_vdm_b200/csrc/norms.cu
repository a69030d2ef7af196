// HBM-bound kernels around the tensor-core ops: GroupNorm(32) (+SiLU, + channel concat), LayerNorm (+ frame
// positional embedding add). Channels-last bf16 in / out, fp32 (fp64 for the group sums) arithmetic.
#include "common.h"
#include "ptx.cuh"

namespace ttvdm {

// ------------------------------------------------------------------------------------------------ GroupNorm
// Thread t owns channel pairs {t, t+T, ...} (T = blockDim.x divides C/2) so its group ids are loop invariant.

__device__ __forceinline__ float2 load_pair(const __nv_bfloat16* x1, int c1, int ld1, const __nv_bfloat16* x2,
                                            int ld2, long long row, int pair) {
  const int c = pair * 2;
  const __nv_bfloat16* p = (c < c1) ? (x1 + row * ld1 + c) : (x2 + row * ld2 + (c - c1));
  return unpack_bf16(*reinterpret_cast<const uint32_t*>(p));
}

__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x1, int c1, int ld1,
                                const __nv_bfloat16* __restrict__ x2, int c2, int ld2, int rows_per_inst,
                                int rows_per_cta, double* __restrict__ sums) {
  __shared__ float bins[64];
  const int C = c1 + c2;
  const int cpg = C / 32;
  const int pairs = C / 2;
  const int inst = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, rows_per_inst);
  for (int i = threadIdx.x; i < 64; i += blockDim.x) bins[i] = 0.f;
  __syncthreads();
  for (int pair = threadIdx.x; pair < pairs; pair += blockDim.x) {
    float s = 0.f, ss = 0.f;
    const long long base = (long long)inst * rows_per_inst;
    for (int r = r0; r < r1; ++r) {
      const float2 v = load_pair(x1, c1, ld1, x2, ld2, base + r, pair);
      s += v.x + v.y;
      ss += v.x * v.x + v.y * v.y;
    }
    const int grp = (pair * 2) / cpg;  // cpg is even: both channels of a pair share the group
    atomicAdd(&bins[grp * 2], s);
    atomicAdd(&bins[grp * 2 + 1], ss);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64; i += blockDim.x) atomicAdd(&sums[inst * 64 + i], (double)bins[i]);
}

__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x1, int c1, int ld1,
                                const __nv_bfloat16* __restrict__ x2, int c2, int ld2, int rows_per_inst,
                                int rows_per_cta, const double* __restrict__ sums, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float eps, int silu, __nv_bfloat16* __restrict__ out,
                                int ldo) {
  __shared__ float s_mean[32], s_rstd[32];
  const int C = c1 + c2;
  const int cpg = C / 32;
  const int pairs = C / 2;
  const int inst = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, rows_per_inst);
  if (threadIdx.x < 32) {
    const double n = (double)rows_per_inst * cpg;
    const double m = sums[inst * 64 + threadIdx.x * 2] / n;
    double var = sums[inst * 64 + threadIdx.x * 2 + 1] / n - m * m;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = (float)m;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const long long base = (long long)inst * rows_per_inst;
  for (int pair = threadIdx.x; pair < pairs; pair += blockDim.x) {
    const int c = pair * 2;
    const int grp = c / cpg;
    const float rs = s_rstd[grp], mu = s_mean[grp];
    const float a0 = rs * gamma[c], a1 = rs * gamma[c + 1];
    const float b0 = beta[c] - mu * a0, b1 = beta[c + 1] - mu * a1;
    for (int r = r0; r < r1; ++r) {
      const float2 v = load_pair(x1, c1, ld1, x2, ld2, base + r, pair);
      float y0 = v.x * a0 + b0, y1 = v.y * a1 + b1;
      if (silu) {
        y0 = y0 / (1.f + __expf(-y0));
        y1 = y1 / (1.f + __expf(-y1));
      }
      *reinterpret_cast<uint32_t*>(out + (base + r) * ldo + c) = pack_bf16(y0, y1);
    }
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// One warp per row, row held in registers (C <= 2560).
template <int kPairsPerLane>
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int rows, int C,
                                 const float* __restrict__ addvec, int F, int S, __nv_bfloat16* __restrict__ sum_out,
                                 int ldsum, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                 __nv_bfloat16* __restrict__ out, int ldo) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const long long row = warp;
  const int pairs = C / 2;
  float2 v[kPairsPerLane];
  const float* av = nullptr;
  if (addvec != nullptr) av = addvec + (size_t)((row / S) % F) * C;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kPairsPerLane; ++i) {
    const int pair = lane + i * 32;
    if (pair < pairs) {
      v[i] = unpack_bf16(*reinterpret_cast<const uint32_t*>(x + row * ldx + pair * 2));
      if (av != nullptr) {
        const float2 e = *reinterpret_cast<const float2*>(av + pair * 2);
        v[i].x += e.x;
        v[i].y += e.y;
        if (sum_out != nullptr) {
          const uint32_t pk = pack_bf16(v[i].x, v[i].y);
          *reinterpret_cast<uint32_t*>(sum_out + row * ldsum + pair * 2) = pk;
          v[i] = unpack_bf16(pk);  // normalise exactly what downstream residuals will read
        }
      }
      s += v[i].x + v[i].y;
    } else {
      v[i] = make_float2(0.f, 0.f);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kPairsPerLane; ++i) {
    const int pair = lane + i * 32;
    if (pair < pairs) {
      const float d0 = v[i].x - mean, d1 = v[i].y - mean;
      ss += d0 * d0 + d1 * d1;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / C + eps);
#pragma unroll
  for (int i = 0; i < kPairsPerLane; ++i) {
    const int pair = lane + i * 32;
    if (pair < pairs) {
      const float2 g2 = *reinterpret_cast<const float2*>(gamma + pair * 2);
      const float2 b2 = *reinterpret_cast<const float2*>(beta + pair * 2);
      const float y0 = (v[i].x - mean) * rstd * g2.x + b2.x;
      const float y1 = (v[i].y - mean) * rstd * g2.y + b2.y;
      *reinterpret_cast<uint32_t*>(out + row * ldo + pair * 2) = pack_bf16(y0, y1);
    }
  }
}

static int gn_threads(int pairs) {
  // largest divisor of `pairs` that is a multiple of 32 and <= 512
  for (int t = 512; t >= 32; t -= 32)
    if (pairs % t == 0) return t;
  return 0;
}

}  // namespace ttvdm

using namespace ttvdm;

extern "C" int ttvdm_groupnorm(const ttvdm_groupnorm_params* p, void* stream_) {
  if (int rc = ensure_init()) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!p || !p->x1 || !p->out || !p->stats || !p->gamma || !p->beta) return fail(TTVDM_ERR_SHAPE, "groupnorm: null");
  const int c2 = p->x2 ? p->c2 : 0;
  const int C = p->c1 + c2;
  if (C % 64 != 0 || p->c1 % 2 != 0) return fail(TTVDM_ERR_SHAPE, "groupnorm: C=%d must be a multiple of 64", C);
  if (p->rows <= 0 || p->rows_per_inst <= 0 || p->rows % p->rows_per_inst != 0)
    return fail(TTVDM_ERR_SHAPE, "groupnorm: rows=%d rows_per_inst=%d", p->rows, p->rows_per_inst);
  const int n_inst = p->rows / p->rows_per_inst;
  const int threads = gn_threads(C / 2);
  if (threads == 0) return fail(TTVDM_ERR_SHAPE, "groupnorm: unsupported C=%d", C);
  // fixed chunk => the fp32 partial sums (and therefore the statistics) do not depend on how many sequences this
  // rank holds: a sharded half-pair reproduces the whole-pair result
  const int rows_per_cta = 32;
  const int chunks = (p->rows_per_inst + rows_per_cta - 1) / rows_per_cta;
  cudaError_t e = cudaMemsetAsync(p->stats, 0, (size_t)n_inst * 64 * sizeof(double), stream);
  if (e != cudaSuccess) return fail(TTVDM_ERR_CUDA, "groupnorm: memset: %s", cudaGetErrorString(e));
  dim3 grid(chunks, n_inst);
  const __nv_bfloat16* x1 = static_cast<const __nv_bfloat16*>(p->x1);
  const __nv_bfloat16* x2 = static_cast<const __nv_bfloat16*>(p->x2);
  gn_stats_kernel<<<grid, threads, 0, stream>>>(x1, p->c1, p->ld1, x2, c2, p->ld2, p->rows_per_inst, rows_per_cta,
                                                static_cast<double*>(p->stats));
  TTVDM_CHECK_LAUNCH("gn_stats_kernel");
  gn_apply_kernel<<<grid, threads, 0, stream>>>(x1, p->c1, p->ld1, x2, c2, p->ld2, p->rows_per_inst, rows_per_cta,
                                                static_cast<const double*>(p->stats), p->gamma, p->beta, p->eps,
                                                p->silu, static_cast<__nv_bfloat16*>(p->out), p->ldo);
  TTVDM_CHECK_LAUNCH("gn_apply_kernel");
  return 0;
}

extern "C" int ttvdm_layernorm(const ttvdm_layernorm_params* p, void* stream_) {
  if (int rc = ensure_init()) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!p || !p->x || !p->out || !p->gamma || !p->beta) return fail(TTVDM_ERR_SHAPE, "layernorm: null");
  if (p->C % 2 != 0 || p->C > 64 * 40) return fail(TTVDM_ERR_SHAPE, "layernorm: unsupported C=%d", p->C);
  if (p->rows <= 0) return fail(TTVDM_ERR_SHAPE, "layernorm: rows=%d", p->rows);
  if (p->addvec && (p->F <= 0 || p->S <= 0)) return fail(TTVDM_ERR_SHAPE, "layernorm: addvec needs F,S");
  const int ppl = (p->C / 2 + 31) / 32;
  const int threads = 256;
  const int grid = (p->rows + (threads / 32) - 1) / (threads / 32);
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(p->x);
  __nv_bfloat16* so = static_cast<__nv_bfloat16*>(p->sum_out);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p->out);
#define LN_LAUNCH(PPL)                                                                                             \
  layernorm_kernel<PPL><<<grid, threads, 0, stream>>>(x, p->ldx, p->rows, p->C, p->addvec, p->F, p->S, so, p->ldsum, \
                                                      p->gamma, p->beta, p->eps, o, p->ldo)
  if (ppl <= 5) LN_LAUNCH(5);
  else if (ppl <= 10) LN_LAUNCH(10);
  else if (ppl <= 20) LN_LAUNCH(20);
  else LN_LAUNCH(40);
#undef LN_LAUNCH
  TTVDM_CHECK_LAUNCH("layernorm_kernel");
  return 0;
}
