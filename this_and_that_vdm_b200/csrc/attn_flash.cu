// K5 / K6 / K8 — flash-style attention on tcgen05 for head_dim 64.
//
//   S = Q K^T   : tcgen05.mma M=128 N=128 K=64, Q/K tiles staged by TMA (128B swizzle), S in TMEM cols [0,128)
//   softmax     : 128 threads (one query row each) read S with tcgen05.ld, online max/sum in fp32 registers,
//                 write P (bf16) into a 128B-swizzled K-major smem tile
//   O += P V    : tcgen05.mma M=128 N=64 K=128, V tile used in place as the MN-major B operand, O accumulates in TMEM
//                 cols [128,192) across KV tiles; the reference max is only moved when it grows by more than 2^8
//                 (lazy rescale), in which case the softmax warps scale their O rows in TMEM (tcgen05.ld/st).
//
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM allocator   warps 4..7: softmax / epilogue
//
// 112 KB smem + 256 TMEM columns per CTA -> two CTAs per SM, so one CTA's softmax overlaps the other's MMAs.
// The same kernel serves spatial self-attention (KV = the image's own tokens), spatial cross-attention (one KV
// tile = the <=128 context tokens of the image's batch element) and temporal cross-attention, where the
// reference's context-selection quirk (row (b,s) reads context (b*S+s) mod B,
// svd/diffusion_arch/transformer_temporal.py:310-319) becomes a per-row mask over one KV tile per context.
#include "common.h"
#include "ptx.cuh"

namespace ttvdm {

constexpr int kQT = 128;   // query rows per CTA
constexpr int kKT = 128;   // keys per KV tile
constexpr int kD = 64;     // head dim
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KB
constexpr int kAttnThreads = 256;
constexpr int kAttnSmem = kTileBytes * (1 + 2 + 2 + 2) + 256;

enum { KV_SELF = 0, KV_CROSS_SPATIAL = 1, KV_CROSS_TEMPORAL = 2 };

struct AttnArgs {
  int kv_mode;
  int seq_q;      // query rows per unit (image)
  int seq_kv;     // SELF: keys per unit; CROSS: L
  int q_tiles;    // ceil(seq_q / 128)
  int heads;
  int F, S;       // rows ordered (b, f, s); unit = (b, f)
  int n_ctx, batch_offset;
  float scale_log2;  // scale * log2(e)
  __nv_bfloat16* out;
  int ldo;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attn_flash_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTileBytes;          // 2 stages
  uint8_t* sV = smem + 3 * kTileBytes;      // 2 stages
  uint8_t* sP = smem + 5 * kTileBytes;      // 2 K-blocks of 64 keys
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * kTileBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* v_full = bars + 3;   // [2]
  uint64_t* k_empty = bars + 5;  // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* p_ready = bars + 10;
  uint64_t* o_done = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  // ---- which tile am I
  const int qt = blockIdx.x;
  const int head = blockIdx.y;
  const int unit = blockIdx.z;
  const int q_row0 = unit * g.seq_q + qt * kQT;  // global query row of tile row 0
  const int q_valid = min(kQT, g.seq_q - qt * kQT);
  int n_kv_tiles, kv_row_base;
  if (g.kv_mode == KV_SELF) {
    n_kv_tiles = (g.seq_kv + kKT - 1) / kKT;
    kv_row_base = unit * g.seq_kv;
  } else if (g.kv_mode == KV_CROSS_SPATIAL) {
    n_kv_tiles = 1;
    kv_row_base = (g.batch_offset + unit / g.F) * g.seq_kv;
  } else {
    n_kv_tiles = g.n_ctx;
    kv_row_base = 0;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_done, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, kTileBytes);
      tma_load_2d(sQ, &tmQ, q_full, head * kD, q_row0);
    }
    for (int j = 0; j < n_kv_tiles; ++j) {
      const int st = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int kv_row = (g.kv_mode == KV_CROSS_TEMPORAL) ? j * g.seq_kv : kv_row_base + j * kKT;
      mbar_wait(&k_empty[st], ph ^ 1);
      if (lane == 0) {
        mbar_expect_tx(&k_full[st], kTileBytes);
        tma_load_2d(sK + st * kTileBytes, &tmK, &k_full[st], head * kD, kv_row);
      }
      mbar_wait(&v_empty[st], ph ^ 1);
      if (lane == 0) {
        mbar_expect_tx(&v_full[st], kTileBytes);
        tma_load_2d(sV + st * kTileBytes, &tmV, &v_full[st], head * kD, kv_row);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major
    const uint64_t q_desc = make_sdesc_sw128(smem_u32(sQ), 16, 1024);
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    if (lane == 0) {
      const uint64_t k_desc = make_sdesc_sw128(smem_u32(sK), 16, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_ss(tmem_S, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k != 0);
      tc_commit(&k_empty[0]);
      tc_commit(s_full);
    }
    __syncwarp();
    for (int j = 0; j < n_kv_tiles; ++j) {
      const int st = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(p_ready, j & 1);
      mbar_wait(&v_full[st], ph);
      tc_fence_after();
      if (lane == 0) {
        // O_j = P (128 x 128, K-major, two 64-key blocks) * V (128 keys x 64, MN-major: 8-key groups 1024 B apart)
        const uint32_t sp = smem_u32(sP);
        const uint32_t sv = smem_u32(sV + st * kTileBytes);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t p_desc = make_sdesc_sw128(sp + (k >> 2) * kTileBytes + (k & 3) * 32, 16, 1024);
          const uint64_t v_desc = make_sdesc_sw128(sv + k * 2048, 16, 1024);
          tc_mma_ss(tmem_O, p_desc, v_desc, idesc_pv, (j | k) != 0);  // O accumulates across KV tiles
        }
        tc_commit(&v_empty[st]);
        tc_commit(o_done);
      }
      __syncwarp();
      if (j + 1 < n_kv_tiles) {
        const int st2 = (j + 1) & 1;
        const uint32_t ph2 = ((j + 1) >> 1) & 1;
        mbar_wait(&k_full[st2], ph2);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t k_desc = make_sdesc_sw128(smem_u32(sK + st2 * kTileBytes), 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_ss(tmem_S, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k != 0);
          tc_commit(&k_empty[st2]);
          tc_commit(s_full);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax + output
    const int qd = warp & 3;
    const int r = qd * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_addr = uint32_t(qd * 32) << 16;
    int my_ctx = -1;
    if (g.kv_mode == KV_CROSS_TEMPORAL) {
      // global row -> (b, f, s); temporal batch row (b, s) reads context (b*S + s) mod n_ctx  [reference quirk]
      const long long row = (long long)q_row0 + r;
      const int s = (int)(row % g.S);
      const int b = (int)(row / ((long long)g.F * g.S)) + g.batch_offset;
      my_ctx = (int)(((long long)b * g.S + s) % g.n_ctx);
    }
    const float c2 = g.scale_log2;
    float m_used = -INFINITY;  // (stale) row max the exponentials are taken against
    float l_run = 0.f;
    uint8_t* const prow0 = sP + r * 128;
    const int rx = r & 7;
    for (int j = 0; j < n_kv_tiles; ++j) {
      int kv_valid = kKT;
      if (g.kv_mode == KV_SELF) kv_valid = min(kKT, g.seq_kv - j * kKT);
      else kv_valid = g.seq_kv;
      const bool row_off = (g.kv_mode == KV_CROSS_TEMPORAL) && (my_ctx != j);
      const bool masked = (kv_valid < kKT) || (g.kv_mode == KV_CROSS_TEMPORAL);  // CTA-uniform
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // ---- pass 1: row max of this tile
      float mx = -INFINITY;
      if (!masked) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_S + lane_addr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_S + lane_addr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        if (row_off) mx = -INFINITY;
      }
      // ---- lazy rescale: only move the reference max when it grows by more than 2^8 (exp2 domain)
      const bool grow = (mx - m_used) * c2 > 8.0f;  // j == 0: m_used = -inf -> true (NaN if both -inf -> false)
      const float m_new = grow ? mx : m_used;
      if (j > 0) {
        // O_{0..j-1} has been accumulated in TMEM; PV_{j-1} must be complete before O is touched / P is overwritten
        mbar_wait(o_done, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? ex2((m_used - m_new) * c2) : 1.0f;  // m_used = -inf -> 0 (row still empty)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_O + lane_addr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32(tmem_O + lane_addr + c * 32, v);
          }
          tmem_st_wait();
          l_run *= alpha;
        }
      }
      m_used = m_new;
      const float ms = (m_used == -INFINITY) ? 0.f : m_used * c2;
      // ---- pass 2: probabilities -> 128B-swizzled smem tile (bf16), row sum
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_S + lane_addr + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
        if (!masked) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2(fmaf(__uint_as_float(v[i]), c2, -ms));
            const float p1 = ex2(fmaf(__uint_as_float(v[i + 1]), c2, -ms));
            sum0 += p0;
            sum1 += p1;
            pk[i >> 1] = pack_bf16(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0 = ex2(fmaf(__uint_as_float(v[i]), c2, -ms));
            float p1 = ex2(fmaf(__uint_as_float(v[i + 1]), c2, -ms));
            if (row_off || c * 32 + i >= kv_valid) p0 = 0.f;
            if (row_off || c * 32 + i + 1 >= kv_valid) p1 = 0.f;
            sum0 += p0;
            sum1 += p1;
            pk[i >> 1] = pack_bf16(p0, p1);
          }
        }
        // keys [c*32, c*32+32) live in K-block (c>>1), 16-byte chunks ((c&1)*4 .. +3), XOR-swizzled by (row & 7)
        uint8_t* prow = prow0 + (c >> 1) * kTileBytes;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = ((c & 1) * 4 + ch) ^ rx;
          *reinterpret_cast<uint4*>(prow + chunk * 16) =
              make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        }
      }
      l_run += sum0 + sum1;
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    // ---- epilogue: O / l
    mbar_wait(o_done, (n_kv_tiles - 1) & 1);
    tc_fence_after();
    const float inv = (l_run > 0.f) ? 1.f / l_run : 0.f;
    __nv_bfloat16* orow = g.out + (long long)(q_row0 + r) * g.ldo + head * kD;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_O + lane_addr + c * 32, v);
      tmem_ld_wait();
      if (r < q_valid) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t w4[4];
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2)
            w4[k2] = pack_bf16(__uint_as_float(v[i * 8 + k2 * 2]) * inv, __uint_as_float(v[i * 8 + k2 * 2 + 1]) * inv);
          *(reinterpret_cast<uint4*>(orow + c * 32) + i) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<256>(tmem_base);
}

static int launch_attn(const void* q, int ldq, long long q_rows, const void* k, int ldk, const void* v, int ldv,
                       long long kv_rows, const AttnArgs& g, int units, cudaStream_t stream) {
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  const uint32_t box[2] = {kD, 128};
  {
    uint64_t dims[2] = {(uint64_t)g.heads * kD, (uint64_t)q_rows};
    uint64_t str[1] = {(uint64_t)ldq * 2};
    if ((rc = make_tmap_bf16(&tmQ, q, 2, dims, str, box))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)g.heads * kD, (uint64_t)kv_rows};
    uint64_t strk[1] = {(uint64_t)ldk * 2};
    uint64_t strv[1] = {(uint64_t)ldv * 2};
    if ((rc = make_tmap_bf16(&tmK, k, 2, dims, strk, box))) return rc;
    if ((rc = make_tmap_bf16(&tmV, v, 2, dims, strv, box))) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_flash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e != cudaSuccess) return fail(TTVDM_ERR_CUDA, "attn: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  if (g.heads > 65535 || units > 65535) return fail(TTVDM_ERR_SHAPE, "attn: grid too large");
  dim3 grid(g.q_tiles, g.heads, units);
  attn_flash_kernel<<<grid, kAttnThreads, kAttnSmem, stream>>>(tmQ, tmK, tmV, g);
  TTVDM_CHECK_LAUNCH("attn_flash_kernel");
  return 0;
}

}  // namespace ttvdm

using namespace ttvdm;

extern "C" int ttvdm_attn_spatial(const ttvdm_attn_params* p, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!p || !p->q || !p->k || !p->v || !p->out) return fail(TTVDM_ERR_SHAPE, "attn_spatial: null");
  if (p->n_img <= 0 || p->heads <= 0 || p->seq <= 0) return fail(TTVDM_ERR_SHAPE, "attn_spatial: empty");
  if ((p->ldq | p->ldk | p->ldv | p->ldo) % 8 != 0) return fail(TTVDM_ERR_SHAPE, "attn_spatial: ld %% 8 != 0");
  AttnArgs g{};
  g.kv_mode = KV_SELF;
  g.seq_q = p->seq;
  g.seq_kv = p->seq;
  g.q_tiles = (p->seq + kQT - 1) / kQT;
  g.heads = p->heads;
  g.F = 1;
  g.S = p->seq;
  g.n_ctx = 1;
  g.batch_offset = 0;
  g.scale_log2 = p->scale * 1.4426950408889634f;
  g.out = static_cast<__nv_bfloat16*>(p->out);
  g.ldo = p->ldo;
  const long long rows = (long long)p->n_img * p->seq;
  return launch_attn(p->q, p->ldq, rows, p->k, p->ldk, p->v, p->ldv, rows, g, p->n_img,
                     static_cast<cudaStream_t>(stream_));
}

extern "C" int ttvdm_attn_cross(const ttvdm_xattn_params* p, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!p || !p->q || !p->kc || !p->vc || !p->out) return fail(TTVDM_ERR_SHAPE, "attn_cross: null");
  if (p->L <= 0 || p->L > kKT) return fail(TTVDM_ERR_SHAPE, "attn_cross: L=%d (1..128)", p->L);
  if (p->F <= 0 || p->S <= 0 || p->rows <= 0 || p->rows % (p->F * p->S) != 0)
    return fail(TTVDM_ERR_SHAPE, "attn_cross: rows=%d not a multiple of F*S=%d", p->rows, p->F * p->S);
  if ((p->ldq | p->ldo) % 8 != 0) return fail(TTVDM_ERR_SHAPE, "attn_cross: ld %% 8 != 0");
  const int b_local = p->rows / (p->F * p->S);
  if (p->n_ctx <= 0 || (!p->temporal && p->batch_offset + b_local > p->n_ctx))
    return fail(TTVDM_ERR_SHAPE, "attn_cross: batch %d+%d exceeds n_ctx=%d", p->batch_offset, b_local, p->n_ctx);
  AttnArgs g{};
  g.kv_mode = p->temporal ? KV_CROSS_TEMPORAL : KV_CROSS_SPATIAL;
  g.seq_q = p->S;
  g.seq_kv = p->L;
  g.q_tiles = (p->S + kQT - 1) / kQT;
  g.heads = p->heads;
  g.F = p->F;
  g.S = p->S;
  g.n_ctx = p->n_ctx;
  g.batch_offset = p->batch_offset;
  g.scale_log2 = p->scale * 1.4426950408889634f;
  g.out = static_cast<__nv_bfloat16*>(p->out);
  g.ldo = p->ldo;
  const int C = p->heads * kD;
  return launch_attn(p->q, p->ldq, p->rows, p->kc, C, p->vc, C, (long long)p->n_ctx * p->L, g, b_local * p->F,
                     static_cast<cudaStream_t>(stream_));
}
