// HBM-bound kernels around the tensor-core ops: GroupNorm(32) (+SiLU, + channel concat), LayerNorm (+ frame
// positional embedding add). Channels-last bf16 in / out, fp32 (fp64 for the group sums) arithmetic.
#include "common.h"
#include "ptx.cuh"

namespace ttvdm {

// ------------------------------------------------------------------------------------------------ GroupNorm
// 16-byte vectorised. A CTA owns `rows_per_cta` consecutive rows of one group instance. Per source tensor (the
// optional second one is the up-block skip of torch.cat([h, skip], 1)) the first `active = (T / V) * V` threads
// (V = C_src / 8 vectors per row) keep a FIXED vector column, so group ids / affine coefficients are loop invariant
// and partial sums live in registers. Partials: fp32 per thread -> smem bins -> one fp64 atomic per (CTA, group).

struct GnSrc {
  const __nv_bfloat16* x;
  int c, ld, c_off;  // channels, row stride, channel offset inside the concatenated tensor
};

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

__global__ void __launch_bounds__(512)
gn_stats_kernel(GnSrc s1, GnSrc s2, int C, int rows_per_inst, int rows_per_cta, double* __restrict__ sums) {
  __shared__ double bins[64];  // fp64: the order of the atomics then only matters below ~1e-16 relative
  const int cpg = C / 32;
  const int inst = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, rows_per_inst);
  for (int i = threadIdx.x; i < 64; i += blockDim.x) bins[i] = 0.0;
  __syncthreads();
  const long long base = (long long)inst * rows_per_inst;
#pragma unroll 1
  for (int k = 0; k < 2; ++k) {
    const GnSrc s = k == 0 ? s1 : s2;
    if (s.c == 0) continue;
    const int V = s.c >> 3;
    const int rpi = blockDim.x / V;  // rows per iteration
    if ((int)threadIdx.x >= rpi * V) continue;
    const int col = threadIdx.x % V, rofs = threadIdx.x / V;
    float sm[8], sq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[i] = sq[i] = 0.f;
    // 4 independent 16-byte loads per thread and step, and the next step's loads are issued before the current
    // step's arithmetic (8 loads in flight): the kernel is latency-bound otherwise
    auto load4 = [&](uint4 (&u)[4], int r) {
#pragma unroll
      for (int b = 0; b < 4; ++b)
        u[b] = (r + b * rpi < r1) ? __ldg(reinterpret_cast<const uint4*>(s.x + (base + r + b * rpi) * s.ld) + col)
                                  : make_uint4(0, 0, 0, 0);
    };
    uint4 u[4], un[4];
    load4(u, r0 + rofs);
    for (int r = r0 + rofs; r < r1; r += 4 * rpi) {
      load4(un, r + 4 * rpi);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float f[8];
        unpack8(u[b], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          sm[i] += f[i];
          sq[i] += f[i] * f[i];
        }
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) u[b] = un[b];
    }
    const int c0 = s.c_off + col * 8;
    // merge the (at most few) groups this vector touches before going to shared memory
    int g_prev = c0 / cpg;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gi = (c0 + i) / cpg;
      if (gi != g_prev) {
        atomicAdd(&bins[g_prev * 2], (double)a);
        atomicAdd(&bins[g_prev * 2 + 1], (double)b);
        a = b = 0.f;
        g_prev = gi;
      }
      a += sm[i];
      b += sq[i];
    }
    atomicAdd(&bins[g_prev * 2], (double)a);
    atomicAdd(&bins[g_prev * 2 + 1], (double)b);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64; i += blockDim.x) atomicAdd(&sums[inst * 64 + i], bins[i]);
}

// `sums` (group sums of the sources that went through gn_stats_kernel; NULL if none did) plus, per source, the
// per-(instance, channel pair) sums the producing GEMM's epilogue accumulated (ttvdm_gemm gn_stats_out) — folded into
// the 32 group sums by every CTA (C / 2 <= 2560 doubles from L2: a few KB next to the >= 100 KB the CTA normalises).
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(512, 2)
gn_apply_kernel(GnSrc s1, GnSrc s2, int rows_per_inst, int rows_per_cta, const double* __restrict__ sums,
                const double* __restrict__ ps1, const double* __restrict__ ps2,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
                __nv_bfloat16* __restrict__ out, int ldo) {
  __shared__ float s_mean[32], s_rstd[32];
  __shared__ double s_bins[64];
  const int C = s1.c + s2.c;
  const int cpg = C / 32;
  const int inst = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, rows_per_inst);
  if (ps1 != nullptr || ps2 != nullptr) {
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_bins[i] = sums != nullptr ? sums[inst * 64 + i] : 0.0;
    __syncthreads();
    // pairs never straddle a group (cpg is even for every channel count of the path); host checks it
    for (int pi = threadIdx.x; pi < (C >> 1); pi += blockDim.x) {
      const int c = 2 * pi;
      const double* src = nullptr;
      if (c < s1.c) {
        if (ps1 != nullptr) src = ps1 + ((size_t)inst * (s1.c >> 1) + pi) * 2;
      } else if (ps2 != nullptr) {
        src = ps2 + ((size_t)inst * (s2.c >> 1) + (pi - (s1.c >> 1))) * 2;
      }
      if (src != nullptr) {
        const int gi = c / cpg;
        atomicAdd(&s_bins[gi * 2], src[0]);
        atomicAdd(&s_bins[gi * 2 + 1], src[1]);
      }
    }
    __syncthreads();
    sums = s_bins - inst * 64;  // same indexing below
  }
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
#pragma unroll 1
  for (int k = 0; k < 2; ++k) {
    const GnSrc s = k == 0 ? s1 : s2;
    if (s.c == 0) continue;
    const int V = s.c >> 3;
    const int rpi = blockDim.x / V;
    if ((int)threadIdx.x >= rpi * V) continue;
    const int col = threadIdx.x % V, rofs = threadIdx.x / V;
    const int c0 = s.c_off + col * 8;
    float ka[8], kb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int gi = (c0 + i) / cpg;
      ka[i] = s_rstd[gi] * gamma[c0 + i];
      kb[i] = beta[c0 + i] - s_mean[gi] * ka[i];
      if (silu) {
        ka[i] *= 0.5f;
        kb[i] *= 0.5f;
      }
    }
    auto load4 = [&](uint4 (&u)[4], int r) {
#pragma unroll
      for (int b = 0; b < 4; ++b)
        u[b] = (r + b * rpi < r1) ? __ldg(reinterpret_cast<const uint4*>(s.x + (base + r + b * rpi) * s.ld) + col)
                                  : make_uint4(0, 0, 0, 0);
    };
    uint4 u[4], un[4];
    load4(u, r0 + rofs);
    for (int r = r0 + rofs; r < r1; r += 4 * rpi) {
      load4(un, r + 4 * rpi);  // next step's loads are in flight while this step is normalised and stored
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (r + b * rpi >= r1) break;
        float f[8];
        unpack8(u[b], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // ka / kb carry a factor 1/2 when SiLU follows: y * sigmoid(y) = h + h * tanh(h) with h = y / 2 — one MUFU
          // operation per element instead of two (ex2 + rcp), which kept the XU pipe ~80 % busy at HBM speed
          const float h = fmaf(f[i], ka[i], kb[i]);
          f[i] = silu ? fmaf(h, tanh_approx(h), h) : h;
        }
        *reinterpret_cast<uint4*>(out + (base + r + b * rpi) * ldo + c0) =
            make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) u[b] = un[b];
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

// Vectorised variant for the channel counts of the SVD transformers (C = 40 * kLanes: 320 / 640 / 1280): kLanes lanes
// share a row, each lane owns five 16-byte vectors (lane l reads vectors l, l + kLanes, ...: every access of the lane
// group is a contiguous 128-byte line or more), 32 / kLanes rows per warp. Same arithmetic order per element as
// layernorm_kernel up to the reduction tree.
template <int kLanes>
__global__ void __launch_bounds__(256, 2)
layernorm_vec_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int rows, const float* __restrict__ addvec,
                     int F, int S, __nv_bfloat16* __restrict__ sum_out, int ldsum,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                     __nv_bfloat16* __restrict__ out, int ldo) {
  // Persistent: the grid is one wave of CTAs and every warp walks row groups gw, gw + n_warps, ... with the NEXT group's
  // five 16-byte loads issued before the current group is reduced (round 2: 8064 one-shot CTAs per level-0 call ran at
  // 4.3 TB/s; a warp's load -> shuffle -> load gamma/beta -> store chain was exposed once per CTA). Two CTAs per SM at
  // 128 registers measured best (level 0: 0.0705 -> 0.0676 ms; three CTAs at 80 registers spill: 0.098 ms).
  constexpr int kC = 40 * kLanes;
  constexpr int kRowsPerWarp = 32 / kLanes;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int sub = lane % kLanes;
  const long long n_groups = ((long long)rows + kRowsPerWarp - 1) / kRowsPerWarp;
  auto load_raw = [&](long long grp, uint4 (&r)[5]) {
    const long long row = grp * kRowsPerWarp + lane / kLanes;
    if (row < rows) {
#pragma unroll
      for (int i = 0; i < 5; ++i) r[i] = *reinterpret_cast<const uint4*>(x + row * ldx + (sub + i * kLanes) * 8);
    }
  };
  uint4 raw[5] = {};
  if (warp < n_groups) load_raw(warp, raw);
  for (long long grp = warp; grp < n_groups; grp += n_warps) {
    const long long row = grp * kRowsPerWarp + lane / kLanes;
    const bool live = row < rows;  // whole lane groups are live or not; dead groups still take part in the shuffles
    float v[5][8];
    float s = 0.f;
    if (live) {
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16(w[k]);
          v[i][2 * k] = f.x;
          v[i][2 * k + 1] = f.y;
        }
      }
    }
    // the raw registers are free again: the next group's loads fly under this group's reductions and stores
    if (grp + n_warps < n_groups) load_raw(grp + n_warps, raw);
    // loop-variant zero: keeps the 80 gamma / beta values out of registers (they are L1 hits; hoisted out of the loop they
    // cost two thirds of the occupancy)
    const int lv0 = (int)(grp >> 62);
    if (live) {
      const float* av = (addvec != nullptr) ? addvec + (size_t)((row / S) % F) * kC : nullptr;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        if (av != nullptr) {
          const float4 e0 = *reinterpret_cast<const float4*>(av + (sub + i * kLanes) * 8);
          const float4 e1 = *reinterpret_cast<const float4*>(av + (sub + i * kLanes) * 8 + 4);
          v[i][0] += e0.x; v[i][1] += e0.y; v[i][2] += e0.z; v[i][3] += e0.w;
          v[i][4] += e1.x; v[i][5] += e1.y; v[i][6] += e1.z; v[i][7] += e1.w;
          if (sum_out != nullptr) {
            uint32_t pk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              pk[k] = pack_bf16(v[i][2 * k], v[i][2 * k + 1]);
              const float2 f = unpack_bf16(pk[k]);  // normalise exactly what downstream residuals will read
              v[i][2 * k] = f.x;
              v[i][2 * k + 1] = f.y;
            }
            *reinterpret_cast<uint4*>(sum_out + row * ldsum + (sub + i * kLanes) * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[i][k];
      }
    }
#pragma unroll
    for (int o = kLanes / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / kC;
    float ss = 0.f;
    if (live) {
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float d = v[i][k] - mean;
          ss += d * d;
        }
    }
#pragma unroll
    for (int o = kLanes / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (live) {
      const float rstd = rsqrtf(ss / kC + eps);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int c0 = (sub + i * kLanes) * 8;
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + c0 + lv0), g1 = *reinterpret_cast<const float4*>(gamma + c0 + 4 + lv0);
        const float4 b0 = *reinterpret_cast<const float4*>(beta + c0 + lv0), b1 = *reinterpret_cast<const float4*>(beta + c0 + 4 + lv0);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint32_t pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          pk[k] = pack_bf16((v[i][2 * k] - mean) * rstd * gg[2 * k] + bb[2 * k],
                            (v[i][2 * k + 1] - mean) * rstd * gg[2 * k + 1] + bb[2 * k + 1]);
        *reinterpret_cast<uint4*>(out + row * ldo + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
}

}  // namespace ttvdm

using namespace ttvdm;

extern "C" int ttvdm_groupnorm(const ttvdm_groupnorm_params* p, void* stream_) {
  if (int rc = ensure_init()) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!p || !p->x1 || !p->out || !p->gamma || !p->beta) return fail(TTVDM_ERR_SHAPE, "groupnorm: null");
  if (!p->stats && (!p->pstats1 || (p->x2 && p->c2 && !p->pstats2)))
    return fail(TTVDM_ERR_SHAPE, "groupnorm: stats workspace is required unless every source has producer statistics");
  const int c2 = p->x2 ? p->c2 : 0;
  const int C = p->c1 + c2;
  if (C % 32 != 0 || p->c1 % 8 != 0 || c2 % 8 != 0 || p->c1 > 4096 || c2 > 4096)
    return fail(TTVDM_ERR_SHAPE, "groupnorm: channels %d+%d must be multiples of 8 (sum %% 32 == 0)", p->c1, c2);
  if (p->ld1 % 8 != 0 || (c2 && p->ld2 % 8 != 0) || p->ldo % 8 != 0)
    return fail(TTVDM_ERR_SHAPE, "groupnorm: row strides must be multiples of 8 elements");
  if (p->rows <= 0 || p->rows_per_inst <= 0 || p->rows % p->rows_per_inst != 0)
    return fail(TTVDM_ERR_SHAPE, "groupnorm: rows=%d rows_per_inst=%d", p->rows, p->rows_per_inst);
  const int n_inst = p->rows / p->rows_per_inst;
  // The chunking depends on rows_per_inst and the channel counts only (not on how many sequences this rank holds):
  // the fp32 partial sums, and so the statistics, are identical for a sharded half-pair and the whole pair.
  // A CTA walks its rows in steps of `granule` = 4 * (512 / vectors per row) rows; it gets 4..8 whole steps (setup —
  // group statistics, affine coefficients — is ~200 instructions per thread and must be amortised; too many rows per
  // CTA would leave the 5-D temporal norms with fewer CTAs than SMs).
  const int vmax0 = (p->c1 > c2 ? p->c1 : c2) / 8;  // >= 1 (channels are multiples of 8, C % 32 == 0)
  const int granule = 4 * (vmax0 < 512 ? 512 / vmax0 : 1);
  int steps = (p->rows_per_inst + 48 * granule - 1) / (48 * granule);
  if (steps < 4) steps = 4;
  if (steps > 8) steps = 8;
  int rows_per_cta = granule * steps;
  if (rows_per_cta > p->rows_per_inst) rows_per_cta = p->rows_per_inst;
  const int chunks = (p->rows_per_inst + rows_per_cta - 1) / rows_per_cta;
  const double* ps1 = static_cast<const double*>(p->pstats1);
  const double* ps2 = c2 ? static_cast<const double*>(p->pstats2) : nullptr;
  if ((ps1 || ps2) && (C / 32) % 2 != 0)
    return fail(TTVDM_ERR_SHAPE, "groupnorm: producer statistics need an even number of channels per group (C=%d)", C);
  // sources whose statistics came out of the producing GEMM's epilogue are not read by the statistics pass
  const bool need_pass = !ps1 || (c2 && !ps2);
  if (need_pass) {
    cudaError_t e = cudaMemsetAsync(p->stats, 0, (size_t)n_inst * 64 * sizeof(double), stream);
    if (e != cudaSuccess) return fail(TTVDM_ERR_CUDA, "groupnorm: memset: %s", cudaGetErrorString(e));
  }
  const int vmax = (p->c1 > c2 ? p->c1 : c2) / 8;
  int threads = 512;
  if (vmax > threads) return fail(TTVDM_ERR_SHAPE, "groupnorm: more than 4096 channels per source");
  if (vmax * rows_per_cta < threads) threads = ((vmax * rows_per_cta + 31) / 32) * 32;
  if (threads < vmax) threads = ((vmax + 31) / 32) * 32;
  dim3 grid(chunks, n_inst);
  GnSrc s1{static_cast<const __nv_bfloat16*>(p->x1), p->c1, p->ld1, 0};
  GnSrc s2{static_cast<const __nv_bfloat16*>(p->x2), c2, p->ld2, p->c1};
  if (need_pass) {
    GnSrc t1 = s1, t2 = s2;
    if (ps1) t1.c = 0;  // skipped by the kernel (c == 0); c_off of the other source is unaffected
    if (ps2) t2.c = 0;
    gn_stats_kernel<<<grid, threads, 0, stream>>>(t1, t2, C, p->rows_per_inst, rows_per_cta, static_cast<double*>(p->stats));
    TTVDM_CHECK_LAUNCH("gn_stats_kernel");
  }
  // The apply pass is elementwise, so its chunking is free: ONE wave of CTAs (as many as fit on the SMs at once, never
  // more), each walking a long row range — the per-CTA setup (group statistics from the producer sums, affine
  // coefficients) is ~2-3 us and was 35-40 % of a CTA's life with the statistics pass' 4-step chunks.
  static int n_sm = 0, occ512 = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ512, gn_apply_kernel, 512, 0) != cudaSuccess || occ512 < 1) occ512 = 1;
  }
  int a_rows_per_cta = rows_per_cta, a_chunks = chunks;
  if (threads == 512) {
    const int slots = n_sm * occ512;
    int want = slots / n_inst;
    if (want < 1) want = 1;
    a_rows_per_cta = ((p->rows_per_inst + want - 1) / want + granule - 1) / granule * granule;
    if (a_rows_per_cta < rows_per_cta) a_rows_per_cta = rows_per_cta;
    if (a_rows_per_cta > p->rows_per_inst) a_rows_per_cta = p->rows_per_inst;
    a_chunks = (p->rows_per_inst + a_rows_per_cta - 1) / a_rows_per_cta;
  }
  gn_apply_kernel<<<dim3(a_chunks, n_inst), threads, 0, stream>>>(s1, s2, p->rows_per_inst, a_rows_per_cta,
                                                need_pass ? static_cast<const double*>(p->stats) : nullptr, ps1, ps2,
                                                p->gamma, p->beta, p->eps, p->silu,
                                                static_cast<__nv_bfloat16*>(p->out), p->ldo);
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
  // 16-byte-vector kernel for C = 320 / 640 / 1280 when every row is 16-byte aligned
  const bool aligned = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(o) % 16 == 0) &&
                       (p->ldx % 8 == 0) && (p->ldo % 8 == 0) &&
                       (!so || (reinterpret_cast<uintptr_t>(so) % 16 == 0 && p->ldsum % 8 == 0)) &&
                       (!p->addvec || reinterpret_cast<uintptr_t>(p->addvec) % 16 == 0) &&
                       (reinterpret_cast<uintptr_t>(p->gamma) % 16 == 0) && (reinterpret_cast<uintptr_t>(p->beta) % 16 == 0);
  if (aligned && (p->C == 320 || p->C == 640 || p->C == 1280)) {
    const int lanes = p->C / 40;
    const int rows_per_warp = 32 / lanes;
    const int warps = (p->rows + rows_per_warp - 1) / rows_per_warp;
    int vgrid = (warps + (threads / 32) - 1) / (threads / 32);
    {
      // one wave: as many CTAs as the device holds at once (every warp then loops over its row groups)
      static int n_sm = 0, occ[3] = {0, 0, 0};
      if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], layernorm_vec_kernel<8>, threads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], layernorm_vec_kernel<16>, threads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], layernorm_vec_kernel<32>, threads, 0);
      }
      const int o = occ[lanes == 8 ? 0 : (lanes == 16 ? 1 : 2)];
      if (o > 0 && vgrid > n_sm * o) vgrid = n_sm * o;
    }
#define LNV_LAUNCH(L)                                                                                                \
  layernorm_vec_kernel<L><<<vgrid, threads, 0, stream>>>(x, p->ldx, p->rows, p->addvec, p->F, p->S, so, p->ldsum, p->gamma, \
                                                         p->beta, p->eps, o, p->ldo)
    if (lanes == 8) LNV_LAUNCH(8);
    else if (lanes == 16) LNV_LAUNCH(16);
    else LNV_LAUNCH(32);
#undef LNV_LAUNCH
    TTVDM_CHECK_LAUNCH("layernorm_vec_kernel");
    return 0;
  }
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
