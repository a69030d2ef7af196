// ttvdm_gemm — persistent, warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   warp 0      : TMA producer (A tile 128 x 64 bf16 + B tile BN x 64 bf16 per stage, 128B swizzle)
//   warp 1      : MMA issuer (one lane issues tcgen05.mma, M=128, N=BN, K=16; accumulators in TMEM)
//   warp 2      : TMEM allocator (512 columns = 2 accumulator stages x up to 256 columns)
//   warp 3      : idle
//   warps 4..19 : epilogue (TMEM -> registers -> fused bias / timestep shift / residuals / GEGLU -> global);
//                 16 warps because the epilogue is latency bound (dependent TMEM / global loads), not issue bound
//
// The 3x3 and temporal convolutions never materialise im2col: the producer shifts the TMA box coordinates per
// filter tap and lets TMA zero-fill the halo (out-of-bounds) elements.
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "ptx.cuh"

namespace ttvdm {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kMaxStages = 8;
constexpr int kNumEpiGroups = 4;                        // column groups of epilogue warps (x4 TMEM lane quarters)
constexpr int kNumThreads = 128 + 128 * kNumEpiGroups;   // 640
constexpr int kAccCols = 256;  // TMEM columns per accumulator stage
constexpr int kSmemBudget = 232448;
constexpr int kEpiWarps = 4 * kNumEpiGroups;
constexpr int kEpiBufBytes = 32 * 128;  // per epilogue warp: 32 rows x 64 bf16, 128B-swizzled (one TMA box)

struct GemmArgs {
  int mode;
  int M, N;
  int kc_a1;       // 64-wide K chunks served by tensor map A (rest of a tap: map A2)
  int kc_per_tap;  // K chunks per filter tap
  int taps;        // 1, 9 or 3
  int block_n, stages;
  int m_tiles, n_tiles;
  // conv / tconv geometry
  int H, W, TW, TH, tiles_w, tiles_h;
  // epilogue
  const float* bias;
  const float* rowvec;
  int rows_per_vec, ldrv;
  float s0, s1, s2;
  const __nv_bfloat16* res1;
  const __nv_bfloat16* res2;
  int ldr1, ldr2;
  int geglu;
  void* out;
  int ldo, out_fp32, act;
  int n_img;          // images (CONV) / batch (TCONV): tiles beyond it are padding of an odd CTA pair
  int tap_w, dy0, dx0; // CONV3X3: taps per window row (3, or 2 for the parity convolutions) and offset of the first tap
  int conv_stride;    // CONV3X3: 1 or 2 (input coordinates = stride * output coordinates + tap offset)
  int b_resident;     // 1: this CTA keeps ONE N tile of W (all K chunks) in smem and only streams A tiles
  int ctas_per_n;     // b_resident: CTAs sharing an N tile
  int tma_epi;        // 1: residual tile in / output tile out through per-warp smem + TMA (coalesced, asynchronous)
  int epi_box_w;      // CONV: pixels per image row covered by a warp's 32 tile rows (min(TW, 32))
  // normalisation fusions (see include/ttvdm.h)
  double* gn_stats;   // [M / gn_rpi][N / 2][2]: per (group instance, channel pair) sum / sum of squares of the stored bf16 values
  int gn_rpi;
  float* row_sums;    // [N / 32][M][2]: per row and 32-column chunk, sum / sum of squares of the stored bf16 values
                      // (LayerNorm of the consumer; plain stores, no atomics: every (chunk, row) has exactly one writer)
  int ln_parts;       // consumer: chunks per row in ln_rowsums (= K / 32)
  int contig;         // 1: a CTA walks a CONTIGUOUS range of tiles (few group instances per CTA -> few statistics flushes)
  int tiles_per_cta;
  const float* rs_addvec;  // [rs_add_mod][ld_rs_add] added to the stored values inside the row sums only
  int rs_add_rows, rs_add_mod, ld_rs_add;
  const float* ln_rowsums;  // consumer side: [M][2] of the A operand -> LayerNorm folded into this epilogue
  const float* ln_colsum;   // [N]
  const float* prevec;      // [prevec_mod][ldpv] added to the accumulator before the LayerNorm scale
  const float* ln_row_add;  // [prevec_mod][2]
  int prevec_rows, prevec_mod, ldpv;
  float ln_inv_k, ln_eps;
};

__device__ __forceinline__ void red_add_f64(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void red_add_f32x2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// Column sums over the 32 lanes of a warp: lane r holds v[0..32) = 32 consecutive columns of row r; on return v[0] of
// lane c is sum_r v_r[c]. Recursive halving: 16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h];
      const float keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0];
}

struct TileCoord {
  int n0;          // first output column
  int m0;          // LINEAR: first row
  int img, h0, w0; // CONV: image, first row/col of the TH x TW pixel box; TCONV: img=b, h0=f, w0=s0
};

__device__ __forceinline__ TileCoord tile_coord(const GemmArgs& g, int tile) {
  TileCoord t;
  const int mt = tile / g.n_tiles;
  t.n0 = (tile - mt * g.n_tiles) * g.block_n;
  t.m0 = mt * kBlockM;
  t.img = 0;
  t.h0 = 0;
  t.w0 = 0;
  if (g.mode == TTVDM_A_CONV3X3) {
    const int per_img = g.tiles_w * g.tiles_h;
    t.img = mt / per_img;
    const int r = mt - t.img * per_img;
    const int th = r / g.tiles_w;
    t.h0 = th * g.TH;
    t.w0 = (r - th * g.tiles_w) * g.TW;
  } else if (g.mode == TTVDM_A_TCONV3) {
    // tiles_w = tiles per frame along S; H = F; W = S
    const int per_img = g.tiles_w * g.H;
    t.img = mt / per_img;
    const int r = mt - t.img * per_img;
    t.h0 = r / g.tiles_w;
    t.w0 = (r - t.h0 * g.tiles_w) * kBlockM;
  }
  return t;
}

// exact-erf GELU (diffusers GEGLU uses F.gelu, erf form) with erf from Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7,
// two MUFU ops + 8 FMAs instead of libdevice erff's ~30 instructions; the GEGLU epilogue is instruction bound)
// Persistent tile schedule. Default: tile = blockIdx.x + i * gridDim.x over (m, n) with n fastest. B-resident: the CTA is
// pinned to N tile (blockIdx.x % n_tiles) and walks M tiles blockIdx.x / n_tiles + i * ctas_per_n.
template <int kCtas>
__device__ __forceinline__ int sched_tile(const GemmArgs& g, int i, int cta_rank) {
  if (kCtas == 2) {
    // CTA pair: both CTAs walk the same pair-tiles; CTA r owns M tile 2*pm + r (may be one past the end: all-padding tile)
    const int pm_tiles = (g.m_tiles + 1) >> 1;
    int pt;
    if (g.contig) {
      if (i >= g.tiles_per_cta) return -1;
      pt = ((int)blockIdx.x >> 1) * g.tiles_per_cta + i;
    } else {
      pt = ((int)blockIdx.x >> 1) + i * ((int)gridDim.x >> 1);
    }
    if (pt >= pm_tiles * g.n_tiles) return -1;
    const int pm = pt / g.n_tiles;
    return (2 * pm + cta_rank) * g.n_tiles + (pt - pm * g.n_tiles);
  }
  if (!g.b_resident) {
    int tile;
    if (g.contig) {
      if (i >= g.tiles_per_cta) return -1;
      tile = (int)blockIdx.x * g.tiles_per_cta + i;
    } else {
      tile = (int)blockIdx.x + i * (int)gridDim.x;
    }
    return tile < g.m_tiles * g.n_tiles ? tile : -1;
  }
  int mt;
  if (g.contig) {
    if (i >= g.tiles_per_cta) return -1;
    mt = ((int)blockIdx.x / g.n_tiles) * g.tiles_per_cta + i;
  } else {
    mt = (int)blockIdx.x / g.n_tiles + i * g.ctas_per_n;
  }
  return mt < g.m_tiles ? mt * g.n_tiles + ((int)blockIdx.x % g.n_tiles) : -1;
}

// Group instance of a tile if all of its rows lie in ONE instance (then its statistics go through the CTA's shared-memory
// accumulators), else -1 (LINEAR tiles at the low-resolution levels can straddle instances: direct global atomics).
__device__ __forceinline__ int tile_instance(const GemmArgs& g, const TileCoord& t) {
  if (g.mode == TTVDM_A_LINEAR) {
    const long long last = min((long long)t.m0 + kBlockM - 1, (long long)g.M - 1);
    const int a = t.m0 / g.gn_rpi, b = (int)(last / g.gn_rpi);
    return a == b ? a : -1;
  }
  if (t.img >= g.n_img) return -2;  // all-padding tile of an odd CTA pair: nothing to add
  if (g.mode == TTVDM_A_CONV3X3) return (int)(((long long)t.img * g.H * g.W) / g.gn_rpi);
  return (int)((((long long)t.img * g.H + t.h0) * g.W) / g.gn_rpi);
}

// Epilogue warps only (kEpiWarps * 32 threads, named barrier 1): add the CTA's shared-memory GroupNorm accumulators
// (4 copies, one per TMEM lane quarter, of [N / 2][2] floats: per channel pair sum / sum of squares over the tiles seen
// since the last flush; a copy is only ever touched by the warps of its quarter, each on its own columns, so plain
// read-modify-write is race free — shared-memory float atomics are CAS loops) to the global fp64 bins of `inst`.
__device__ __forceinline__ void gn_flush(const GemmArgs& g, float* acc, int inst, int epi_tid) {
  named_bar_sync(1, kEpiWarps * 32);
  if (inst >= 0) {
    for (int i = epi_tid; i < g.N; i += kEpiWarps * 32) {
      const float v = (acc[i] + acc[g.N + i]) + (acc[2 * g.N + i] + acc[3 * g.N + i]);  // the four lane-quarter copies
      if (v != 0.f) red_add_f64(g.gn_stats + (long long)inst * g.N + i, (double)v);
      acc[i] = acc[g.N + i] = acc[2 * g.N + i] = acc[3 * g.N + i] = 0.f;
    }
  }
  named_bar_sync(1, kEpiWarps * 32);
}

// Returns 2 * gelu(x) = x + |x| * erf(|x| / sqrt(2)) (the caller folds the 0.5 into its own scale factor).
// 12 FMA-pipe instructions + 2 MUFU (rcp, ex2): the GEGLU epilogue of the K = 320 level-0 GEMM is issue-bound, every
// instruction here is paid 330 M times per Euler step.
__device__ __forceinline__ float gelu_erf_x2(float x) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752f, ax, 1.0f)));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  p *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((x * x) * (-0.5f * 1.4426950408889634f)));  // exp(-z^2), z = |x|/sqrt(2)
  return fmaf(ax, fmaf(-p, e, 1.0f), x);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * gelu_erf_x2(x); }

// fp64 atomics of one 32-column chunk's per-pair GroupNorm sums: `a` = the bf16-rounded values of this lane's row
// (zeros for rows that do not exist), inst0..inst1 = group instances touched by the warp's 32 rows.
// GroupNorm sums of one 32-column chunk held one row per lane (`a` = the bf16-rounded values, anything for rows that do
// not exist). ALL 32 lanes must call this. acc != nullptr: the tile lies in one instance -> shared-memory accumulators;
// else one reduction per instance present in the warp's rows, straight to the global fp64 bins.
__device__ __forceinline__ void gn_stats_chunk(const GemmArgs& g, const float (&a)[32], bool valid, long long row, int col0,
                                               int lane, float* acc) {
  const int col = col0 + lane;
  if (acc != nullptr) {
    float q[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) q[i] = valid ? a[i] * a[i] : 0.f;
    const float cq = warp_colsum32(q, lane);
#pragma unroll
    for (int i = 0; i < 32; ++i) q[i] = valid ? a[i] : 0.f;
    const float cs = warp_colsum32(q, lane);
    // lane c holds column col0 + c; bins are per channel PAIR (sum, sum of squares): fold the odd column into the even lane
    const float cs2 = cs + __shfl_down_sync(0xffffffffu, cs, 1), cq2 = cq + __shfl_down_sync(0xffffffffu, cq, 1);
    if ((lane & 1) == 0 && col < g.N) {
      float2* d = reinterpret_cast<float2*>(acc + col);  // this quarter's copy; no other warp touches these columns now
      float2 o = *d;
      o.x += cs2;
      o.y += cq2;
      *d = o;
    }
    return;
  }
  const int mine = valid ? (int)(row / g.gn_rpi) : -1;
  int imin = valid ? mine : 0x7fffffff, imax = mine;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    imin = min(imin, __shfl_xor_sync(0xffffffffu, imin, o));
    imax = max(imax, __shfl_xor_sync(0xffffffffu, imax, o));
  }
  if (imax < 0) return;  // no valid row in this warp (warp-uniform)
  for (int inst = imin; inst <= imax; ++inst) {
    const bool in = valid && mine == inst;
    float q[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) q[i] = in ? a[i] * a[i] : 0.f;
    const float cq = warp_colsum32(q, lane);
#pragma unroll
    for (int i = 0; i < 32; ++i) q[i] = in ? a[i] : 0.f;
    const float cs = warp_colsum32(q, lane);
    if (col < g.N) {
      double* dst = g.gn_stats + (long long)inst * g.N + (col >> 1) * 2;
      red_add_f64(dst, (double)cs);
      red_add_f64(dst + 1, (double)cq);
    }
  }
}

// Direct epilogue of one 32-column chunk (one row per lane). Every lane of the warp calls it (rows that do not exist
// with valid = false): the GroupNorm statistics are a warp-wide reduction.
template <int kFeat>
__device__ __forceinline__ void epilogue_chunk(const GemmArgs& g, const uint32_t (&v)[32], long long out_row, int col0,
                                               const float* rv, bool valid, int lane, float* gn_acc) {
  constexpr bool kRs = kFeat == 2, kGn = kFeat == 3;
  float a[32];
  const bool full = (col0 + 32 <= g.N) && ((g.N & 7) == 0);
  const bool stats = (kRs || kGn) && full;  // host: statistics need bf16, non-GEGLU output with N % 32 == 0
  if (valid) {
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = __uint_as_float(v[i]);
    if (g.bias != nullptr) {
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + col0 + i));
          a[i] += b4.x; a[i + 1] += b4.y; a[i + 2] += b4.z; a[i + 3] += b4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (col0 + i < g.N) a[i] += g.bias[col0 + i];
      }
    }
    if (rv != nullptr) {
      if (full) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(rv + col0 + i));
          a[i] += b4.x; a[i + 1] += b4.y; a[i + 2] += b4.z; a[i + 3] += b4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (col0 + i < g.N) a[i] += rv[col0 + i];
      }
    }
    if (g.geglu) {
      // interleaved (hidden, gate) columns -> 16 outputs
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(g.out) + out_row * g.ldo + (col0 >> 1);
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float o0 = a[4 * j] * gelu_erf(a[4 * j + 1]);
        const float o1 = a[4 * j + 2] * gelu_erf(a[4 * j + 3]);
        pk[j] = pack_bf16(g.s0 * o0, g.s0 * o1);
      }
      if (full) {
        uint4* o4 = reinterpret_cast<uint4*>(o);
        o4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        o4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      } else {
        for (int j = 0; j < 16; ++j)
          if (col0 + 2 * j + 1 < g.N) {
            const float2 f = unpack_bf16(pk[j >> 1]);
            o[j] = __float2bfloat16((j & 1) ? f.y : f.x);
          }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] *= g.s0;
      if (g.res1 != nullptr) {
        const __nv_bfloat16* rp = g.res1 + out_row * g.ldr1 + col0;
        if (full) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 u = *(reinterpret_cast<const uint4*>(rp) + i);  // plain load: `out` may alias the residual
            const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = unpack_bf16(w4[j]);
              a[i * 8 + j * 2] += g.s1 * f.x;
              a[i * 8 + j * 2 + 1] += g.s1 * f.y;
            }
          }
        } else {
          for (int i = 0; i < 32; ++i)
            if (col0 + i < g.N) a[i] += g.s1 * __bfloat162float(rp[i]);
        }
      }
      if (g.res2 != nullptr) {
        const __nv_bfloat16* rp = g.res2 + out_row * g.ldr2 + col0;
        if (full) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 u = *(reinterpret_cast<const uint4*>(rp) + i);  // plain load: `out` may alias the residual
            const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = unpack_bf16(w4[j]);
              a[i * 8 + j * 2] += g.s2 * f.x;
              a[i * 8 + j * 2 + 1] += g.s2 * f.y;
            }
          }
        } else {
          for (int i = 0; i < 32; ++i)
            if (col0 + i < g.N) a[i] += g.s2 * __bfloat162float(rp[i]);
        }
      }
      if (g.act == 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = a[i] / (1.f + __expf(-a[i]));
      }
      if (g.out_fp32) {
        float* o = reinterpret_cast<float*>(g.out) + out_row * g.ldo + col0;
        if (full && (g.ldo & 3) == 0) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(o + i) = make_float4(a[i], a[i + 1], a[i + 2], a[i + 3]);
        } else {
          for (int i = 0; i < 32; ++i)
            if (col0 + i < g.N) o[i] = a[i];
        }
      } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(g.out) + out_row * g.ldo + col0;
        if (full) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(a[2 * i], a[2 * i + 1]);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *(reinterpret_cast<uint4*>(o) + i) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          if (stats) {
            // statistics of exactly what the consumer will read: the bf16-rounded values
            if (kRs) {
              uint64_t rs2 = 0ull, rq2 = 0ull;
              const float* av = g.rs_addvec != nullptr
                                    ? g.rs_addvec + (long long)((out_row / g.rs_add_rows) % g.rs_add_mod) * g.ld_rs_add + col0
                                    : nullptr;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                uint64_t f2 = pack_f32x2(a[2 * i], a[2 * i + 1]);
                if (av != nullptr) {
                  const float2 d = __ldg(reinterpret_cast<const float2*>(av) + i);
                  f2 = add_f32x2(f2, pack_f32x2(d.x, d.y));
                }
                rs2 = add_f32x2(rs2, f2);
                rq2 = fma_f32x2(f2, f2, rq2);
              }
              float sa, sb, qa, qb;
              unpack_f32x2(rs2, sa, sb);
              unpack_f32x2(rq2, qa, qb);
              *reinterpret_cast<float2*>(g.row_sums + ((long long)(col0 >> 5) * g.M + out_row) * 2) = make_float2(sa + sb, qa + qb);
            }
            if (kGn) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {  // GroupNorm sums are taken over the stored (rounded) values
                const float2 f = unpack_bf16(pk[i]);
                a[2 * i] = f.x;
                a[2 * i + 1] = f.y;
              }
            }
          }
        } else {
          for (int i = 0; i < 32; ++i)
            if (col0 + i < g.N) o[i] = __float2bfloat16(a[i]);
        }
      }
    }
  }
  if (kGn && stats) gn_stats_chunk(g, a, valid, out_row, col0, lane, gn_acc);  // whole warp
}

// kFeat selects ONE normalisation fusion so that every variant only carries its own epilogue code and registers:
//   0 plain   1 LayerNorm of A folded in (consumer)   2 LayerNorm row sums of the output (producer)
//   3 GroupNorm pair sums of the output (producer)
template <int kCtas, int kFeat>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
            const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
            const __grid_constant__ CUtensorMap tmRes, const GemmArgs g) {
  constexpr bool kLn = kFeat == 1, kRs = kFeat == 2, kGn = kFeat == 3;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte aligned carve-up (SWIZZLE_128B atoms are 1024 B)
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* res_bar = tempty_bar + 3;  // [kEpiWarps] residual tile landed in the warp's staging buffer
  uint64_t* bres_bar = res_bar + kEpiWarps;  // resident W tile landed
  uint8_t* tiles = smem + 1024;
  const int b_chunk_bytes = (g.block_n / kCtas) * kBlockK * 2;  // CTA pair: each CTA stages half of the W tile
  const int stage_bytes = g.b_resident ? kABytes : kABytes + b_chunk_bytes;
  const int cta_rank = kCtas == 2 ? (int)cluster_ctarank() : 0;
  // b_resident only: the pinned W tile sits behind the operand ring and the TMA-epilogue staging buffers
  uint8_t* bres = tiles + g.stages * stage_bytes + (g.tma_epi ? kEpiWarps * kEpiBufBytes : 0);
  // GroupNorm accumulators of this CTA ([N / 2][2] floats), behind everything else
  float* gn_acc_base = reinterpret_cast<float*>(bres + (g.b_resident ? g.taps * g.kc_per_tap * b_chunk_bytes : 0));
  if (kGn)
    for (int i = threadIdx.x; i < 4 * g.N; i += kNumThreads) gn_acc_base[i] = 0.f;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kc = g.taps * g.kc_per_tap;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    if (g.tma_epi) {
      tma_prefetch_desc(&tmOut);
      tma_prefetch_desc(&tmRes);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < g.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    mbar_init(bres_bar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4 * kNumEpiGroups * kCtas);  // one arrival per epilogue warp (of both CTAs of a pair)
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    if (kCtas == 2) tmem_alloc_pair<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (kCtas == 2) cluster_sync_all();  // the peer's barriers are initialised before anything can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // one thread, chosen with elect.sync (NOT `lane == 0`: only then does ptxas keep tensor-map / descriptor operands in
    // uniform registers instead of wrapping every TMA / MMA instruction in an ELECT + R2UR + branch sequence)
    if (elect_one()) {
    int stage = 0;
    uint32_t phase = 0;
    uint8_t* sa_ring = tiles;
    const uint32_t tx_bytes = kCtas == 1 ? (uint32_t)(g.b_resident ? kABytes : stage_bytes) : 2u * (uint32_t)stage_bytes;
    if (g.b_resident) {
      // the CTA's N tile of W: all K chunks, once
      const int n0 = (blockIdx.x % g.n_tiles) * g.block_n;
      mbar_expect_tx(bres_bar, num_kc * b_chunk_bytes);
      for (int kc = 0; kc < num_kc; ++kc) tma_load_2d(bres + kc * b_chunk_bytes, &tmB, bres_bar, kc * kBlockK, n0);
    }
    for (int ti = 0;; ++ti) {
      const int tile = sched_tile<kCtas>(g, ti, cta_rank);
      if (tile < 0) break;
      const TileCoord t = tile_coord(g, tile);
      // tap / channel-chunk counters instead of kc / kc_per_tap: this is ONE thread feeding the whole operand ring, it runs
      // at one instruction per several cycles next to the busy epilogue warps, and a stage of the level-0 shapes lasts
      // only ~320 tensor-pipe cycles — every division here shows up in the GEMM time (a runtime `tap / tap_w` cost the
      // level-0 / level-1 convolutions 20 %)
      int kc = 0;
      for (int tap = 0; tap < g.taps; ++tap) {
        int ty, tx;
        if (g.tap_w == 3) {
          ty = tap / 3;
          tx = tap - 3 * ty;
        } else {
          ty = tap >> 1;
          tx = tap & 1;
        }
        const int cs = g.conv_stride;
        const int ax = g.mode == TTVDM_A_CONV3X3 ? cs * t.w0 + tx + g.dx0 : t.w0;
        const int ay = g.mode == TTVDM_A_CONV3X3 ? cs * t.h0 + ty + g.dy0 : t.h0 + tap - 1;
        for (int cc = 0; cc < g.kc_per_tap; ++cc, ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = sa_ring;
          uint8_t* sb = sa + kABytes;
          if (kCtas == 1 || cta_rank == 0) mbar_expect_tx(&full_bar[stage], tx_bytes);  // pair: both CTAs' bytes land on the leader's barrier
          if (g.mode == TTVDM_A_LINEAR) {
            const CUtensorMap* ma = (cc < g.kc_a1) ? &tmA : &tmA2;
            const int ck = (cc < g.kc_a1 ? cc : cc - g.kc_a1) * kBlockK;
            if (kCtas == 1) tma_load_2d(sa, ma, &full_bar[stage], ck, t.m0);
            else tma_load_2d_pair(sa, ma, &full_bar[stage], ck, t.m0);
          } else {
            if (kCtas == 1) tma_load_4d(sa, &tmA, &full_bar[stage], cc * kBlockK, ax, ay, t.img);
            else tma_load_4d_pair(sa, &tmA, &full_bar[stage], cc * kBlockK, ax, ay, t.img);
          }
          if (kCtas == 2) tma_load_2d_pair(sb, &tmB, &full_bar[stage], kc * kBlockK, t.n0 + cta_rank * (g.block_n >> 1));
          else if (!g.b_resident) tma_load_2d(sb, &tmB, &full_bar[stage], kc * kBlockK, t.n0);
          sa_ring += stage_bytes;
          if (++stage == g.stages) {
            stage = 0;
            phase ^= 1;
            sa_ring = tiles;
          }
        }
      }
    }
    }
    __syncwarp();
  } else if (warp == 1 && cta_rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA of a pair), one elected thread
    if (elect_one()) {
    const uint32_t idesc = make_idesc_bf16(kBlockM * kCtas, g.block_n);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    if (g.b_resident) mbar_wait(bres_bar, 0);
    for (;; ++it) {
      const int tile = sched_tile<kCtas>(g, it, cta_rank);
      if (tile < 0) break;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kAccCols;
      for (int kc = 0; kc < num_kc; ++kc) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        {
          const uint32_t sa = smem_u32(tiles + stage * stage_bytes);
          const uint32_t sb = g.b_resident ? smem_u32(bres + kc * b_chunk_bytes) : sa + kABytes;
          const uint64_t a_desc = make_sdesc_sw128(sa, 16, 1024);
          const uint64_t b_desc = make_sdesc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
            if (kCtas == 1) tc_mma_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | k) != 0);
            else tc_mma_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | k) != 0);
          }
          if (kCtas == 1) {
            tc_commit(&empty_bar[stage]);
            if (kc == num_kc - 1) tc_commit(&tfull_bar[acc]);
          } else {  // multicast: frees the stage / publishes the accumulator in BOTH CTAs
            tc_commit_pair(&empty_bar[stage]);
            if (kc == num_kc - 1) tc_commit_pair(&tfull_bar[acc]);
          }
        }
        if (++stage == g.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int grp = (warp - 4) >> 2;  // which 32-column chunks (ch % kNumEpiGroups == grp)
    const int chunks = g.block_n / 32;
    const int epi_tid = threadIdx.x - 128;
    int cur_inst = -1;  // group instance whose sums sit in gn_acc_base (same value in every epilogue thread)
    if (g.tma_epi) {
      // ---- TMA epilogue: each warp owns a 32-row x 64-column strip per tile. The residual strip is fetched by TMA
      // into the warp's 128B-swizzled staging buffer while the tile's MMAs run, the fp32 epilogue math happens in
      // registers (one row per lane), and the bf16 strip leaves through a TMA store — every global access is a full
      // 128-byte line and none of them occupies the LSU.
      const int ew = warp - 4;
      uint8_t* sbuf = tiles + g.stages * stage_bytes + ew * kEpiBufBytes;
      uint64_t* rbar = &res_bar[ew];
      const bool has_res = g.res1 != nullptr;
      const int strip = grp * 64;  // first accumulator column of this warp's strip inside the tile
      uint32_t rphase = 0;         // parity of rbar: flips only on tiles where this warp actually fetched a residual
      int it = 0;
      for (;; ++it) {
        const int tile = sched_tile<kCtas>(g, it, cta_rank);
        if (tile < 0) break;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const TileCoord t = tile_coord(g, tile);
        const int col0 = t.n0 + strip;
        const bool active = (strip < g.block_n) && (col0 < g.N);
        float* gn_acc = nullptr;
        if (kGn) {
          const int ti = tile_instance(g, t);
          if (ti >= 0) {
            if (ti != cur_inst) {
              if (cur_inst >= 0) gn_flush(g, gn_acc_base, cur_inst, epi_tid);
              cur_inst = ti;
            }
            gn_acc = gn_acc_base + q * g.N;  // this lane quarter's copy
          }
        }
        // box coordinates of the warp's 32 tile rows
        int c1, c2 = 0, c3 = 0;
        if (g.mode == TTVDM_A_LINEAR) {
          c1 = t.m0 + q * 32;
        } else if (g.mode == TTVDM_A_CONV3X3) {
          const int r0 = q * 32;
          c1 = t.w0 + (r0 % g.TW);  // TW >= 32: column offset inside the pixel-box row; TW < 32: 0
          c2 = t.h0 + r0 / g.TW;
          c3 = t.img;
        } else {
          c1 = t.w0 + q * 32;
          c2 = t.h0;
          c3 = t.img;
        }
        long long my_row;
        bool my_valid;
        {
          const int r = q * 32 + lane;
          if (g.mode == TTVDM_A_LINEAR) {
            my_row = t.m0 + r;
            my_valid = my_row < g.M;
          } else if (g.mode == TTVDM_A_CONV3X3) {
            const int th = r / g.TW, tw = r - th * g.TW;
            my_valid = (t.h0 + th < g.H) && (t.w0 + tw < g.W) && (t.img < g.n_img);
            my_row = ((long long)t.img * g.H + t.h0 + th) * g.W + t.w0 + tw;
          } else {
            my_valid = ((t.w0 + r) < g.W) && (t.img < g.n_img);
            my_row = ((long long)t.img * g.H + t.h0) * g.W + t.w0 + r;
          }
        }
        if (active) {
          if (lane == 0) {
            tma_store_wait_read();  // the previous tile's store has finished reading the staging buffer
            if (has_res) {
              mbar_expect_tx(rbar, kEpiBufBytes);
              if (g.mode == TTVDM_A_LINEAR) tma_load_2d(sbuf, &tmRes, rbar, col0, c1);
              else tma_load_4d(sbuf, &tmRes, rbar, col0, c1, c2, c3);
            }
          }
          __syncwarp();
        }
        const float* rv = nullptr;
        if (g.rowvec != nullptr && my_valid) rv = g.rowvec + (my_row / g.rows_per_vec) * g.ldrv;
        float ln_nrm = 0.f, ln_rstd = 1.f;  // ln_nrm = -rstd * mean
        const float* pv = nullptr;
        if (kLn && my_valid) {
          float2 rs = make_float2(0.f, 0.f);
          // [K / 32][M][2], coalesced across the warp's 32 rows; 10 loads in flight at a time (a one-at-a-time loop is a
          // chain of L2 latencies at the head of every tile's epilogue: measured +80 % on the level-0 qkv GEMM)
          for (int j0 = 0; j0 < g.ln_parts; j0 += 10) {
            float2 pj[10];
#pragma unroll
            for (int u = 0; u < 10; ++u)
              pj[u] = (j0 + u < g.ln_parts)
                          ? __ldg(reinterpret_cast<const float2*>(g.ln_rowsums) + (long long)(j0 + u) * g.M + my_row)
                          : make_float2(0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 10; ++u) {
              rs.x += pj[u].x;
              rs.y += pj[u].y;
            }
          }
          if (g.prevec_mod > 0) {
            const int pi = (int)((my_row / g.prevec_rows) % g.prevec_mod);
            if (g.prevec != nullptr) pv = g.prevec + (long long)pi * g.ldpv;
            if (g.ln_row_add != nullptr) {
              const float2 ra = __ldg(reinterpret_cast<const float2*>(g.ln_row_add) + pi);
              rs.x += ra.x;
              rs.y += ra.y;
            }
          }
          const float ln_mean = rs.x * g.ln_inv_k;
          ln_rstd = rsqrtf(fmaxf(fmaf(-ln_mean, ln_mean, rs.y * g.ln_inv_k), 0.f) + g.ln_eps);
          ln_nrm = -ln_rstd * ln_mean;
        }
        uint64_t row_s2 = 0ull, row_q2 = 0ull;  // LayerNorm sums of this lane's row over one chunk (two f32x2 lanes each)
        const float* rs_av = (kRs && g.rs_addvec != nullptr && my_valid)
                                 ? g.rs_addvec + (long long)((my_row / g.rs_add_rows) % g.rs_add_mod) * g.ld_rs_add
                                 : nullptr;
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        if (active) {
          if (has_res) {
            mbar_wait(rbar, rphase);
            rphase ^= 1;
          }
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int ch_col = strip + cc * 32;
            if (ch_col >= g.block_n) break;  // warp-uniform (BN = 160 / 96 / 32: last strip is one chunk wide)
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + acc * kAccCols + ch_col, v);
            tmem_ld_wait();
            const int gc = t.n0 + ch_col;  // global column of v[0]
            float a[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = __uint_as_float(v[i]);
            if (kLn && gc + 32 <= g.N) {
              // LayerNorm of the A operand folded in:  rstd * (acc + prevec) + (bias - rstd * mean * colsum)
              if (pv != nullptr) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(pv + gc + i));
                  a[i] += b4.x; a[i + 1] += b4.y; a[i + 2] += b4.z; a[i + 3] += b4.w;
                }
              }
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(g.ln_colsum + gc + i));
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + gc + i));
                a[i] = fmaf(a[i], ln_rstd, fmaf(c4.x, ln_nrm, b4.x));
                a[i + 1] = fmaf(a[i + 1], ln_rstd, fmaf(c4.y, ln_nrm, b4.y));
                a[i + 2] = fmaf(a[i + 2], ln_rstd, fmaf(c4.z, ln_nrm, b4.z));
                a[i + 3] = fmaf(a[i + 3], ln_rstd, fmaf(c4.w, ln_nrm, b4.w));
              }
            } else
            if (gc + 32 <= g.N) {
              if (g.bias != nullptr) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + gc + i));
                  a[i] += b4.x; a[i + 1] += b4.y; a[i + 2] += b4.z; a[i + 3] += b4.w;
                }
              }
              if (rv != nullptr) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(rv + gc + i));
                  a[i] += b4.x; a[i + 1] += b4.y; a[i + 2] += b4.z; a[i + 3] += b4.w;
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (gc + i < g.N) a[i] += (g.bias ? g.bias[gc + i] : 0.f) + (rv ? rv[gc + i] : 0.f);
            }
            if (g.geglu) {
              // interleaved (hidden, gate) columns -> 16 outputs = 32 bytes of the 64-byte staging row (no swizzle)
              uint32_t pk[8];
              const float hs = 0.5f * g.s0;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float o0 = (hs * a[4 * j]) * gelu_erf_x2(a[4 * j + 1]);
                const float o1 = (hs * a[4 * j + 2]) * gelu_erf_x2(a[4 * j + 3]);
                pk[j] = pack_bf16(o0, o1);
              }
              uint4* sp = reinterpret_cast<uint4*>(sbuf + lane * 64 + cc * 32);
              sp[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              sp[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
              continue;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] *= g.s0;
            if (g.res2 != nullptr && my_valid && gc + 32 <= g.N) {
              const __nv_bfloat16* rp = g.res2 + my_row * g.ldr2 + gc;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 u = *(reinterpret_cast<const uint4*>(rp) + i);  // plain load: `out` may alias the residual
                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = unpack_bf16(w4[j]);
                  a[i * 8 + j * 2] += g.s2 * f.x;
                  a[i * 8 + j * 2 + 1] += g.s2 * f.y;
                }
              }
            }
            // staging row = lane; 16-byte pieces cc*4 .. cc*4+3, XOR-swizzled with (row & 7) (TMA SWIZZLE_128B)
            uint8_t* srow = sbuf + lane * 128;
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
              uint4* sp = reinterpret_cast<uint4*>(srow + (((cc * 4 + pc) ^ (lane & 7)) << 4));
              float o[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] = a[pc * 8 + k];
              if (has_res) {
                const uint4 u = *sp;
                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float2 f = unpack_bf16(w4[k]);
                  o[2 * k] += g.s1 * f.x;
                  o[2 * k + 1] += g.s1 * f.y;
                }
              }
              if (g.act == 1) {
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = o[k] / (1.f + __expf(-o[k]));
              }
              const uint4 pk4 = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
              *sp = pk4;
              if (kRs) {
                // LayerNorm sums from the fp32 values (the bf16 rounding of the stored copy averages out over a row);
                // packed f32x2 adds / FMAs: the level-0 epilogues are issue bound, every instruction counts
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  uint64_t f2 = pack_f32x2(o[2 * k], o[2 * k + 1]);
                  if (rs_av != nullptr) {
                    const float2 d = __ldg(reinterpret_cast<const float2*>(rs_av + gc + pc * 8) + k);
                    f2 = add_f32x2(f2, pack_f32x2(d.x, d.y));
                  }
                  row_s2 = add_f32x2(row_s2, f2);
                  row_q2 = fma_f32x2(f2, f2, row_q2);
                }
              }
            }
            if (kRs && my_valid && gc + 32 <= g.N) {
              // one (32-column chunk, row) slot per lane: 32 consecutive rows = one 256-byte store, no atomics
              float sa, sb, qa, qb;
              unpack_f32x2(row_s2, sa, sb);
              unpack_f32x2(row_q2, qa, qb);
              *reinterpret_cast<float2*>(g.row_sums + ((long long)(gc >> 5) * g.M + my_row) * 2) = make_float2(sa + sb, qa + qb);
            }
            if (kRs) row_s2 = row_q2 = 0ull;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {  // accumulator stage is free for the next tile's MMAs (the leader's MMA warp counts both CTAs)
          if (kCtas == 1) mbar_arrive(&tempty_bar[acc]);
          else mbar_arrive_cluster(&tempty_bar[acc], 0);
        }
        if (active) {
          fence_async_smem();  // every lane: its staging writes become visible to the async proxy
          __syncwarp();
          if (lane == 0) {
            if (g.mode == TTVDM_A_LINEAR) tma_store_2d(&tmOut, sbuf, g.geglu ? (col0 >> 1) : col0, c1);
            else tma_store_4d(&tmOut, sbuf, col0, c1, c2, c3);
            tma_store_commit();
          }
          if (kGn) {
            // GroupNorm statistics of the strip from the staged bf16 values: lane l owns the channel pair (2l, 2l+1) of
            // the strip and walks the 32 rows (one conflict-free 128-byte row read per step: the 16-byte piece holding
            // the pair sits at ((l >> 2) ^ (r & 7))).
            const uint32_t vmask = __ballot_sync(0xffffffffu, my_valid);
            const int pair_col = col0 + 2 * lane;
            const bool col_ok = pair_col < g.N && (2 * lane < g.block_n - strip);
            const uint8_t* sb = sbuf + (lane & 3) * 4;
            if (gn_acc != nullptr) {
              // the whole tile lies in one instance: sums go to this lane quarter's shared-memory accumulators (plain
              // read-modify-write: no other warp of the quarter owns these columns), flushed on instance change
              float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
              if (vmask == 0xffffffffu) {
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                  const float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(sb + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4)));
                  s0 += f.x; s1 += f.y;
                  q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
                }
              } else {
                for (int r = 0; r < 32; ++r) {
                  if (!((vmask >> r) & 1u)) continue;  // warp-uniform
                  const float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(sb + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4)));
                  s0 += f.x; s1 += f.y;
                  q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
                }
              }
              if (col_ok) {
                float2* d = reinterpret_cast<float2*>(gn_acc + pair_col);
                float2 o = *d;
                o.x += s0 + s1;
                o.y += q0 + q1;
                *d = o;
              }
            } else {
              // LINEAR tile straddling group instances (low-resolution levels): rows ascend inside the strip, so the
              // instances form runs; each run goes straight to the global fp64 bins
              const int inst_l = my_valid ? (int)(my_row / g.gn_rpi) : -1;
              float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
              int cur = -1;
              for (int r = 0; r < 32; ++r) {
                const int ir = __shfl_sync(0xffffffffu, inst_l, r);
                if (!((vmask >> r) & 1u)) continue;  // warp-uniform
                if (ir != cur) {
                  if (cur >= 0 && col_ok) {
                    double* dst = g.gn_stats + ((long long)cur * (g.N >> 1) + (pair_col >> 1)) * 2;
                    red_add_f64(dst, (double)(s0 + s1));
                    red_add_f64(dst + 1, (double)(q0 + q1));
                  }
                  s0 = s1 = q0 = q1 = 0.f;
                  cur = ir;
                }
                const float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(sb + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4)));
                s0 += f.x; s1 += f.y;
                q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
              }
              if (cur >= 0 && col_ok) {
                double* dst = g.gn_stats + ((long long)cur * (g.N >> 1) + (pair_col >> 1)) * 2;
                red_add_f64(dst, (double)(s0 + s1));
                red_add_f64(dst + 1, (double)(q0 + q1));
              }
            }
            __syncwarp();  // every lane has read the strip before lane 0 may let the next tile's residual land in it
          }
        }
      }
      if (lane == 0) tma_store_wait_read();
      if (kGn) gn_flush(g, gn_acc_base, cur_inst, epi_tid);
    } else {
    int it = 0;
    for (;; ++it) {
      const int tile = sched_tile<kCtas>(g, it, cta_rank);
      if (tile < 0) break;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const TileCoord t = tile_coord(g, tile);
      const int r = q * 32 + lane;
      long long out_row;
      bool valid;
      if (g.mode == TTVDM_A_LINEAR) {
        out_row = t.m0 + r;
        valid = out_row < g.M;
      } else if (g.mode == TTVDM_A_CONV3X3) {
        const int th = r / g.TW, tw = r - th * g.TW;
        const int h = t.h0 + th, w = t.w0 + tw;
        valid = (h < g.H) && (w < g.W) && (t.img < g.n_img);
        out_row = ((long long)t.img * g.H + h) * g.W + w;
      } else {
        const int s = t.w0 + r;
        valid = (s < g.W) && (t.img < g.n_img);
        out_row = ((long long)t.img * g.H + t.h0) * g.W + s;
      }
      const float* rv = nullptr;
      if (g.rowvec != nullptr && valid) rv = g.rowvec + (out_row / g.rows_per_vec) * g.ldrv;
      float* gn_acc = nullptr;
      if (kGn) {
        const int ti = tile_instance(g, t);
        if (ti >= 0) {
          if (ti != cur_inst) {
            if (cur_inst >= 0) gn_flush(g, gn_acc_base, cur_inst, epi_tid);
            cur_inst = ti;
          }
          gn_acc = gn_acc_base + q * g.N;  // this lane quarter's copy
        }
      }

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      for (int ch = grp; ch < chunks; ch += kNumEpiGroups) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + acc * kAccCols + ch * 32, v);
        tmem_ld_wait();
        const int col0 = t.n0 + ch * 32;
        if (col0 < g.N) epilogue_chunk<kFeat>(g, v, out_row, col0, rv, valid, lane, gn_acc);  // col0 is warp-uniform
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCtas == 1) mbar_arrive(&tempty_bar[acc]);
        else mbar_arrive_cluster(&tempty_bar[acc], 0);
      }
    }
    if (kGn) gn_flush(g, gn_acc_base, cur_inst, epi_tid);
      }
  }

  tc_fence_before();
  __syncthreads();
  if (kCtas == 2) cluster_sync_all();  // neither CTA may exit (or free TMEM) while the other can still signal it
  if (warp == 2) {
    if (kCtas == 2) tmem_dealloc_pair<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

// GroupNorm-statistics producers whose natural tile (e.g. BN = 160 for N = 320) is not a multiple of 64 columns would use
// the direct epilogue, where the column sums are a register butterfly (+35 % on the level-0 temporal conv, which has
// little MMA time to hide it under). Up to this K they take a padded 64-column-granular tile instead so that the TMA
// epilogue (column sums read from the staged strip) applies (measured: level-0 temporal conv +0.010 ms instead of +0.075,
// level-1 +0.012 instead of +0.037); longer K is MMA bound and keeps the exact tile.
static int gn_tma_max_k() {
  static const int v = [] {
    const char* e = getenv("TTVDM_GN_TMA_MAXK");
    return e ? atoi(e) : 2048;
  }();
  return v;
}

static int pick_block_n(int N) {
  // largest UMMA-legal tile (multiple of 32 for the epilogue chunks, <= 256) that divides N; else 128/64/32
  const int cands[] = {256, 192, 160, 128, 96, 64, 32};
  for (int c : cands)
    if (N % c == 0) return c;
  if (N >= 128) return 128;
  if (N >= 64) return 64;
  return 32;
}

}  // namespace ttvdm

using namespace ttvdm;

template <int kCtas, int kFeat>
static cudaError_t launch_variant(const cudaLaunchConfig_t* cfg, const CUtensorMap& tmA, const CUtensorMap& tmA2,
                                  const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmRes,
                                  const GemmArgs& g) {
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<kCtas, kFeat>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  return cudaLaunchKernelEx(cfg, gemm_kernel<kCtas, kFeat>, tmA, tmA2, tmB, tmOut, tmRes, g);
}

extern "C" int ttvdm_gemm(const ttvdm_gemm_params* p, void* stream_) {
  if (int rc = ensure_init()) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!p || !p->a || !p->w || !p->out) return fail(TTVDM_ERR_SHAPE, "gemm: null pointer");
  if (p->M <= 0 || p->N <= 0) return fail(TTVDM_ERR_SHAPE, "gemm: empty problem M=%d N=%d", p->M, p->N);
  const int k2 = p->a2 ? p->k2 : 0;
  if (p->k1 % kBlockK != 0 || k2 % kBlockK != 0)
    return fail(TTVDM_ERR_SHAPE, "gemm: k1=%d / k2=%d must be multiples of 64", p->k1, k2);
  if (p->a2 && p->mode != TTVDM_A_LINEAR) return fail(TTVDM_ERR_SHAPE, "gemm: a2 only in LINEAR mode");
  if (p->geglu && (p->N % 2 != 0 || p->res1 || p->res2 || p->out_fp32))
    return fail(TTVDM_ERR_SHAPE, "gemm: geglu epilogue excludes residuals / fp32 out");
  if (p->rowvec && p->rows_per_vec <= 0) return fail(TTVDM_ERR_SHAPE, "gemm: rows_per_vec must be > 0");
  if (p->rowvec && p->ldrv > 0 && (p->ldrv % 4 != 0 || (reinterpret_cast<uintptr_t>(p->rowvec) & 15)))
    return fail(TTVDM_ERR_SHAPE, "gemm: rowvec must be 16 B aligned with ldrv %% 4 == 0");
  if (p->act != 0 && p->act != 1) return fail(TTVDM_ERR_SHAPE, "gemm: unknown act %d", p->act);
  if (p->gn_stats_out || p->row_sums_out) {
    if (p->geglu || p->out_fp32 || p->N % 32 != 0)
      return fail(TTVDM_ERR_SHAPE, "gemm: gn_stats_out / row_sums_out need a bf16, non-GEGLU output with N %% 32 == 0");
    if (p->gn_stats_out) {
      const long long unit = p->mode == TTVDM_A_CONV3X3 ? (long long)p->H * p->W : (p->mode == TTVDM_A_TCONV3 ? p->W : 1);
      if (p->gn_rows_per_inst <= 0 || p->M % p->gn_rows_per_inst != 0 || p->gn_rows_per_inst % unit != 0)
        return fail(TTVDM_ERR_SHAPE, "gemm: gn_rows_per_inst=%d must divide M=%d and be a multiple of one image / frame",
                    p->gn_rows_per_inst, p->M);
    }
  }
  if (p->gn_stats_out && p->N > 2560) return fail(TTVDM_ERR_SHAPE, "gemm: gn_stats_out supports N <= 2560");
  if (p->ln_rowsums && (!p->ln_colsum || !p->bias || p->rowvec || p->N % 32 != 0 || p->out_fp32 || (p->k1 + k2) % 32 != 0))
    return fail(TTVDM_ERR_SHAPE, "gemm: LayerNorm fold needs ln_colsum, the folded bias, no rowvec, a bf16 output and N %% 32 == 0");
  if (p->ln_rowsums && (p->prevec || p->ln_row_add) && (p->prevec_rows <= 0 || p->prevec_mod <= 0))
    return fail(TTVDM_ERR_SHAPE, "gemm: prevec needs prevec_rows > 0 and prevec_mod > 0");
  const int feat = p->ln_rowsums ? 1 : (p->row_sums_out ? 2 : (p->gn_stats_out ? 3 : 0));
  if ((p->ln_rowsums != nullptr) + (p->row_sums_out != nullptr) + (p->gn_stats_out != nullptr) > 1)
    return fail(TTVDM_ERR_SHAPE, "gemm: ln_rowsums, row_sums_out and gn_stats_out are mutually exclusive (one fusion per launch)");
  if (p->prevec && (p->ldpv % 4 != 0 || (reinterpret_cast<uintptr_t>(p->prevec) & 15)))
    return fail(TTVDM_ERR_SHAPE, "gemm: prevec must be 16 B aligned with ldpv %% 4 == 0");

  GemmArgs g;
  g.conv_stride = 1;
  g.tap_w = 3;
  g.dy0 = g.dx0 = -1;
  memset(&g, 0, sizeof(g));
  g.mode = p->mode;
  g.M = p->M;
  g.N = p->N;
  g.kc_a1 = p->k1 / kBlockK;
  g.kc_per_tap = (p->k1 + k2) / kBlockK;
  g.taps = p->mode == TTVDM_A_CONV3X3 ? 9 : (p->mode == TTVDM_A_TCONV3 ? 3 : 1);
  g.tap_w = 3;
  g.dy0 = g.dx0 = -1;
  if (p->mode == TTVDM_A_CONV3X3 && p->conv_taps == 4) {
    if (p->conv_stride == 2) return fail(TTVDM_ERR_SHAPE, "conv3x3: the 2 x 2 window is a stride-1 mode");
    g.taps = 4;
    g.tap_w = 2;
    g.dy0 = p->conv_dy0;
    g.dx0 = p->conv_dx0;
  } else if (p->mode == TTVDM_A_CONV3X3 && p->conv_taps != 0 && p->conv_taps != 9) {
    return fail(TTVDM_ERR_SHAPE, "conv3x3: conv_taps %d (0 / 9 or 4)", p->conv_taps);
  }
  g.block_n = pick_block_n(p->N);
  const int ktot_pre = g.taps * (p->k1 + k2);
  int want_tma = 0;
  // long-K GEMMs are MMA bound: they keep the deeper operand pipeline (no staging buffers) and the direct epilogue
  if (g.block_n % 64 == 0 && p->N >= 64 && (ktot_pre <= 2048 || p->geglu)) {
    want_tma = 1;
  } else if ((ktot_pre <= 640 || (p->gn_stats_out && ktot_pre <= gn_tma_max_k())) && p->N >= 64) {
    // short-K GEMMs are epilogue / memory bound: take a 64-column-granular tile (fewest N tiles, then least padding)
    // so the TMA epilogue applies, even if that pads the last N tile
    int best = 0, best_tiles = 1 << 30;
    double best_eff = -1.0;
    const int cands[] = {256, 192, 128, 64};
    for (int c : cands) {
      const int nt = (p->N + c - 1) / c;
      const double eff = (double)p->N / ((double)nt * c);
      if (nt < best_tiles || (nt == best_tiles && eff > best_eff + 1e-9)) {
        best = c;
        best_tiles = nt;
        best_eff = eff;
      }
    }
    g.block_n = best;
    want_tma = 1;
  }
  g.n_tiles = (p->N + g.block_n - 1) / g.block_n;
  const int stage_bytes = kABytes + g.block_n * kBlockK * 2;
  // TMA epilogue: bf16 output, no GEGLU, 16-byte aligned rows and bases (TMA global-memory constraints)
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  g.tma_epi = (want_tma && !p->out_fp32 && (!p->geglu || (p->mode == TTVDM_A_LINEAR && p->N % 16 == 0)) && p->N % 8 == 0 && p->ldo % 8 == 0 && al16(p->out) &&
               (!p->res1 || (p->ldr1 % 8 == 0 && al16(p->res1))) && (!p->res2 || (p->res1 && p->ldr2 % 8 == 0 && al16(p->res2) && p->N % 32 == 0)))
                  ? 1 : 0;  // res2 is applied per whole 32-column chunk there: a ragged last chunk takes the direct epilogue
  {
    static const int force_direct = getenv("TTVDM_DIRECT_EPILOGUE") != nullptr;  // A/B switch for profiling only
    if (force_direct) g.tma_epi = 0;
  }
  // CTA-level GroupNorm accumulators ([N / 2][2] floats) live behind the operand ring / staging / resident W tile
  const int gn_bytes = p->gn_stats_out ? ((p->N * 16 + 127) & ~127) : 0;  // 4 copies (one per TMEM lane quarter)
  const int epi_bytes = (g.tma_epi ? kEpiWarps * kEpiBufBytes : 0) + gn_bytes;
  // B-resident schedule for short-K, many-M-tile GEMMs (the K = 320 level-0 linears): re-streaming the W tile from L2 for
  // every 128 rows makes them L2->SM bandwidth bound; pinning one N tile per CTA cuts the operand traffic per tile from
  // (128 + BN) * K to 128 * K elements. Needs the W tile, the epilogue staging and >= 3 A stages in shared memory.
  const int b_res_bytes = g.taps * ((p->k1 + k2) / kBlockK) * g.block_n * kBlockK * 2;
  int stage_b = stage_bytes;
  g.b_resident = 0;
  if (p->mode == TTVDM_A_LINEAR && b_res_bytes + epi_bytes + 3 * kABytes + 2048 <= kSmemBudget && g.n_tiles <= g_num_sms / 4 &&
      (p->M + kBlockM - 1) / kBlockM >= 8 * (g_num_sms / g.n_tiles)) {
    g.b_resident = 1;
    g.ctas_per_n = g_num_sms / g.n_tiles;
    stage_b = kABytes;
  }
  g.stages = (kSmemBudget - 2048 - epi_bytes - (g.b_resident ? b_res_bytes : 0)) / stage_b;
  if (g.stages < 2) {
    g.b_resident = 0;
    stage_b = stage_bytes;
    g.stages = (kSmemBudget - 2048 - epi_bytes) / stage_b;
  }
  if (g.stages > kMaxStages) g.stages = kMaxStages;
  const int ktot = g.taps * (p->k1 + k2);

  CUtensorMap tmA, tmA2, tmB;
  int rc;
  if (p->mode == TTVDM_A_LINEAR) {
    g.m_tiles = (p->M + kBlockM - 1) / kBlockM;
    uint64_t dims[2] = {(uint64_t)p->k1, (uint64_t)p->M};
    uint64_t str[1] = {(uint64_t)p->lda * 2};
    uint32_t box[2] = {kBlockK, kBlockM};
    if ((rc = make_tmap_bf16(&tmA, p->a, 2, dims, str, box))) return rc;
    if (p->a2) {
      uint64_t dims2[2] = {(uint64_t)k2, (uint64_t)p->M};
      uint64_t str2[1] = {(uint64_t)p->lda2 * 2};
      if ((rc = make_tmap_bf16(&tmA2, p->a2, 2, dims2, str2, box))) return rc;
    } else {
      tmA2 = tmA;
    }
  } else if (p->mode == TTVDM_A_CONV3X3) {
    if ((long long)p->n_img * p->H * p->W != p->M) return fail(TTVDM_ERR_SHAPE, "conv3x3: M != n_img*H*W");
    // pick the TW x TH (=128) pixel box wasting the fewest MMA rows
    int best_tw = 128;
    double best_eff = -1.0;
    for (int tw = 128; tw >= 1; tw >>= 1) {
      const int th = 128 / tw;
      const double cover = (double)((p->W + tw - 1) / tw * tw) * ((p->H + th - 1) / th * th);
      const double eff = (double)p->W * p->H / cover;
      if (eff > best_eff + 1e-9) {
        best_eff = eff;
        best_tw = tw;
      }
    }
    g.TW = best_tw;
    g.TH = 128 / best_tw;
    g.H = p->H;
    g.W = p->W;
    g.tiles_w = (p->W + g.TW - 1) / g.TW;
    g.tiles_h = (p->H + g.TH - 1) / g.TH;
    g.m_tiles = p->n_img * g.tiles_w * g.tiles_h;
    const int cs = p->conv_stride == 2 ? 2 : 1;
    if (p->conv_stride != 0 && p->conv_stride != 1 && p->conv_stride != 2)
      return fail(TTVDM_ERR_SHAPE, "conv3x3: stride %d (1 or 2)", p->conv_stride);
    g.conv_stride = cs;
    // stride 2: the input is [n_img, 2H, 2W, k1]; a box of (2 TW) x (2 TH) positions with element stride 2 delivers the
    // TW x TH pixels one tap needs (negative / past-the-end coordinates are zero-filled: the padding of 1)
    const uint64_t Wi = (uint64_t)p->W * cs, Hi = (uint64_t)p->H * cs;
    uint64_t dims[4] = {(uint64_t)p->k1, Wi, Hi, (uint64_t)p->n_img};
    uint64_t str[3] = {(uint64_t)p->lda * 2, (uint64_t)p->lda * 2 * Wi, (uint64_t)p->lda * 2 * Wi * Hi};
    uint32_t box[4] = {kBlockK, (uint32_t)(g.TW * cs), (uint32_t)(g.TH * cs), 1};
    uint32_t est[4] = {1, (uint32_t)cs, (uint32_t)cs, 1};
    if ((rc = make_tmap_bf16(&tmA, p->a, 4, dims, str, box, true, est))) return rc;
    tmA2 = tmA;
  } else if (p->mode == TTVDM_A_TCONV3) {
    if ((long long)p->n_img * p->H * p->W != p->M) return fail(TTVDM_ERR_SHAPE, "tconv3: M != B*F*S");
    g.H = p->H;  // F
    g.W = p->W;  // S
    g.tiles_w = (p->W + kBlockM - 1) / kBlockM;
    g.m_tiles = p->n_img * p->H * g.tiles_w;
    uint64_t dims[4] = {(uint64_t)p->k1, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->n_img};
    uint64_t str[3] = {(uint64_t)p->lda * 2, (uint64_t)p->lda * 2 * p->W, (uint64_t)p->lda * 2 * p->W * p->H};
    uint32_t box[4] = {kBlockK, kBlockM, 1, 1};
    if ((rc = make_tmap_bf16(&tmA, p->a, 4, dims, str, box))) return rc;
    tmA2 = tmA;
  } else {
    return fail(TTVDM_ERR_SHAPE, "gemm: unknown mode %d", p->mode);
  }
  {
    uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)p->N};
    uint64_t str[1] = {(uint64_t)ktot * 2};
    uint32_t box[2] = {kBlockK, (uint32_t)g.block_n};
    if ((rc = make_tmap_bf16(&tmB, p->w, 2, dims, str, box))) return rc;
  }
  CUtensorMap tmOut = tmB, tmRes = tmB;
  if (g.tma_epi) {
    // one box = the 32 tile rows of an epilogue warp x 64 columns
    auto make_io = [&](CUtensorMap* m, const void* base, int ld) -> int {
      if (p->mode == TTVDM_A_LINEAR) {
        uint64_t dims[2] = {(uint64_t)p->N, (uint64_t)p->M};
        uint64_t str[1] = {(uint64_t)ld * 2};
        uint32_t box[2] = {64, 32};
        return make_tmap_bf16(m, base, 2, dims, str, box);
      }
      uint64_t dims[4] = {(uint64_t)p->N, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->n_img};
      uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * p->W, (uint64_t)ld * 2 * p->W * p->H};
      uint32_t bw = 32, bh = 1;
      if (p->mode == TTVDM_A_CONV3X3 && g.TW < 32) {
        bw = (uint32_t)g.TW;
        bh = 32u / (uint32_t)g.TW;
      }
      uint32_t box[4] = {64, bw, bh, 1};
      return make_tmap_bf16(m, base, 4, dims, str, box);
    };
    if (p->geglu) {
      // GEGLU halves the columns: a warp's 64 accumulator columns become a 32-column (64-byte) box, no swizzle
      uint64_t dims[2] = {(uint64_t)(p->N / 2), (uint64_t)p->M};
      uint64_t str[1] = {(uint64_t)p->ldo * 2};
      uint32_t box[2] = {32, 32};
      if ((rc = make_tmap_bf16(&tmOut, p->out, 2, dims, str, box, false))) return rc;
    } else if ((rc = make_io(&tmOut, p->out, p->ldo))) {
      return rc;
    }
    if (p->res1 && (rc = make_io(&tmRes, p->res1, p->ldr1))) return rc;
  }
  g.n_img = p->mode == TTVDM_A_LINEAR ? 1 : p->n_img;
  g.bias = p->bias;
  g.rowvec = p->rowvec;
  g.rows_per_vec = p->rows_per_vec > 0 ? p->rows_per_vec : 1;
  g.ldrv = p->ldrv > 0 ? p->ldrv : p->N;
  g.act = p->act;
  g.s0 = p->s0;
  g.s1 = p->s1;
  g.s2 = p->s2;
  g.res1 = static_cast<const __nv_bfloat16*>(p->res1);
  g.res2 = static_cast<const __nv_bfloat16*>(p->res2);
  g.ldr1 = p->ldr1;
  g.ldr2 = p->ldr2;
  g.geglu = p->geglu;
  g.out = p->out;
  g.ldo = p->ldo;
  g.out_fp32 = p->out_fp32;
  g.gn_stats = static_cast<double*>(p->gn_stats_out);
  g.gn_rpi = p->gn_rows_per_inst > 0 ? p->gn_rows_per_inst : 1;
  g.row_sums = p->row_sums_out;
  g.ln_parts = (p->k1 + k2) / 32;
  {
    // contiguous tile ranges per CTA keep a CTA inside one or two group instances, so the shared-memory GroupNorm
    // accumulators are flushed (fp64 atomics) a handful of times per CTA instead of once per tile
    static const char* e = getenv("TTVDM_GEMM_CONTIG");  // 0 / 1 forces it for every launch (A/B switch), unset = with stats
    g.contig = e ? (atoi(e) != 0) : (p->gn_stats_out != nullptr);
  }
  g.rs_addvec = p->row_sums_out ? p->rs_addvec : nullptr;
  g.rs_add_rows = p->rs_add_rows > 0 ? p->rs_add_rows : 1;
  g.rs_add_mod = p->rs_add_mod > 0 ? p->rs_add_mod : 1;
  g.ld_rs_add = p->ld_rs_add > 0 ? p->ld_rs_add : p->N;
  if (g.rs_addvec && (g.ld_rs_add % 2 != 0 || (reinterpret_cast<uintptr_t>(g.rs_addvec) & 7)))
    return fail(TTVDM_ERR_SHAPE, "gemm: rs_addvec must be 8 B aligned with an even row stride");
  g.ln_rowsums = p->ln_rowsums;
  g.ln_colsum = p->ln_colsum;
  g.ln_eps = p->ln_eps;
  g.ln_inv_k = 1.0f / (float)(p->k1 + k2);
  g.prevec = p->ln_rowsums ? p->prevec : nullptr;
  g.ln_row_add = p->ln_rowsums ? p->ln_row_add : nullptr;
  g.prevec_rows = p->prevec_rows;
  g.prevec_mod = (p->ln_rowsums && (p->prevec || p->ln_row_add)) ? p->prevec_mod : 0;
  g.ldpv = p->ldpv > 0 ? p->ldpv : p->N;
  if (g.ln_rowsums && !g.tma_epi)
    return fail(TTVDM_ERR_SHAPE, "gemm: LayerNorm fold needs the TMA epilogue (16 B aligned bf16 out / res1, N %% 64 == 0 or K <= 640)");
  // vector epilogue needs 16 B aligned rows; otherwise the scalar path is taken only for ragged columns
  if (!p->out_fp32 && ((p->ldo % 8) != 0 && p->N >= 32)) return fail(TTVDM_ERR_SHAPE, "gemm: ldo %% 8 != 0");
  if ((p->res1 && p->ldr1 % 8) || (p->res2 && p->ldr2 % 8)) return fail(TTVDM_ERR_SHAPE, "gemm: ldr %% 8 != 0");

  // CTA pairs (cta_group::2): each CTA stages its own 128 A rows but only HALF of the W tile, the leader issues
  // M = 256 MMAs for both -> the L2 -> SM operand traffic per FLOP drops by a third (A + B/2 instead of A + B)
  int pair = 0;
  {
    static const char* e = getenv("TTVDM_GEMM_PAIR");  // 0 = never, 1 = whenever legal, unset = heuristic
    const bool legal = !g.b_resident && g.m_tiles >= 2 && g.block_n % 32 == 0 && (g.block_n / 2) % 8 == 0;
    if (legal) {
      if (e) pair = atoi(e) != 0;
      // measured on B200 (tools/gpu_kernel_check.py --time): +5-10 % for K >= 640 with N >= 640 and for the long-K
      // 3x3 convolutions of any width (level-0 conv K = 2880, N = 320: 0.413 -> 0.381 ms; level-3 conv: 0.133 -> 0.125 ms),
      // neutral or negative for the K = 320 level-0 linears and GEGLU (epilogue / HBM bound)
      else pair = ((ktot_pre >= 640 && p->N >= 640) || ktot_pre >= 2048) && ((long long)g.m_tiles * g.n_tiles >= g_num_sms / 2);
    }
  }
  const int num_tiles = g.m_tiles * g.n_tiles;
  cudaError_t le;
  if (pair) {
    const int stage_pair = kABytes + (g.block_n / 2) * kBlockK * 2;
    g.stages = (kSmemBudget - 2048 - epi_bytes) / stage_pair;
    if (g.stages > kMaxStages) g.stages = kMaxStages;
    const size_t smem = 2048 + (size_t)g.stages * stage_pair + epi_bytes;
    // B tensor map: box of BN/2 rows
    {
      uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)p->N};
      uint64_t str[1] = {(uint64_t)ktot * 2};
      uint32_t box[2] = {kBlockK, (uint32_t)(g.block_n / 2)};
      if ((rc = make_tmap_bf16(&tmB, p->w, 2, dims, str, box))) return rc;
    }
    const int pair_tiles = ((g.m_tiles + 1) / 2) * g.n_tiles;
    const int max_pairs = g_num_sms / 2;
    const int pairs = pair_tiles < max_pairs ? pair_tiles : max_pairs;
    if (g.contig) g.tiles_per_cta = (pair_tiles + pairs - 1) / pairs;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    switch (feat) {
      case 1: le = launch_variant<2, 1>(&cfg, tmA, tmA2, tmB, tmOut, tmRes, g); break;
      case 2: le = launch_variant<2, 2>(&cfg, tmA, tmA2, tmB, tmOut, tmRes, g); break;
      case 3: le = launch_variant<2, 3>(&cfg, tmA, tmA2, tmB, tmOut, tmRes, g); break;
      default: le = launch_variant<2, 0>(&cfg, tmA, tmA2, tmB, tmOut, tmRes, g); break;
    }
    if (le != cudaSuccess) return fail(TTVDM_ERR_CUDA, "gemm_kernel<2>: %s", cudaGetErrorString(le));
  } else {
    const size_t smem = 2048 + (size_t)g.stages * stage_b + epi_bytes + (g.b_resident ? b_res_bytes : 0);
    const int grid = g.b_resident ? g.ctas_per_n * g.n_tiles : (num_tiles < g_num_sms ? num_tiles : g_num_sms);
    if (g.contig) {
      const int per = g.b_resident ? g.ctas_per_n : grid;       // CTAs sharing the tile list
      const int total = g.b_resident ? g.m_tiles : num_tiles;   // tiles in that list
      g.tiles_per_cta = (total + per - 1) / per;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    switch (feat) {
      case 1: le = launch_variant<1, 1>(&cfg, tmA, tmA2, tmB, tmOut, tmRes, g); break;
      case 2: le = launch_variant<1, 2>(&cfg, tmA, tmA2, tmB, tmOut, tmRes, g); break;
      case 3: le = launch_variant<1, 3>(&cfg, tmA, tmA2, tmB, tmOut, tmRes, g); break;
      default: le = launch_variant<1, 0>(&cfg, tmA, tmA2, tmB, tmOut, tmRes, g); break;
    }
    if (le != cudaSuccess) return fail(TTVDM_ERR_CUDA, "gemm_kernel<1>: %s", cudaGetErrorString(le));
  }
  TTVDM_CHECK_LAUNCH("gemm_kernel");
  return 0;
}
