// K5 / K6 / K8 — flash-style attention on tcgen05 for head_dim 64.
//
//   S = Q K^T   : tcgen05.mma M=128 N=128 K=64, Q/K tiles staged by TMA (128B swizzle), S in TMEM cols [0,128)
//   softmax     : 128 threads (one query row each) read S with tcgen05.ld, online max/sum in fp32 registers,
//                 write P (bf16) into a 128B-swizzled K-major smem tile
//   O += P V    : tcgen05.mma M=128 N=64 K=128, V tile used in place as the MN-major B operand, O accumulates in TMEM
//                 cols [128,192) across KV tiles; the reference max is only moved when it grows by more than 2^8
//                 (lazy rescale), in which case the softmax warps scale their O rows in TMEM (tcgen05.ld/st).
//
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM allocator   warps 4..7 / 8..11: softmax of Q tile 0 / 1
//
// Each CTA owns TWO 128-row Q tiles that share every K/V stage (half the K/V traffic per FLOP). Their chains
// (softmax_t -> P_t -> [O_t += P_t V_j ; S_t = Q_t K_{j+1}^T] -> softmax_t) run out of phase, so while one tile waits
// for its MMAs the other tile's softmax keeps the MUFU busy. 192 KB smem, 384 TMEM columns, one CTA per SM; the
// kernel is bound by MUFU.EX2 (16/clk/SM: 1024 cycles per 128x128 tile vs 512 cycles of MMA).
// The same kernel serves spatial self-attention (KV = the image's own tokens), spatial cross-attention (one KV
// tile = the <=128 context tokens of the image's batch element) and temporal cross-attention, where the
// reference's context-selection quirk (row (b,s) reads context (b*S+s) mod B,
// svd/diffusion_arch/transformer_temporal.py:310-319) becomes a per-row mask over one KV tile per context.
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace ttvdm {

constexpr int kQT = 128;      // query rows per Q tile
constexpr int kQTiles = 2;    // Q tiles per CTA (they share every K/V tile)
constexpr int kKT = 128;      // keys per KV tile
constexpr int kD = 64;        // head dim
constexpr int kKV = 3;        // K/V ring depth
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KB
constexpr int kAttnThreads = 128 + 128 * kQTiles;  // 4 service warps + 4 softmax warps per Q tile
constexpr int kAttnSmem = kTileBytes * (kQTiles + 2 * kKV + 2 * kQTiles) + 256;

enum { KV_SELF = 0, KV_CROSS_SPATIAL = 1, KV_CROSS_TEMPORAL = 2 };

struct AttnArgs {
  int kv_mode;
  int seq_q;      // query rows per unit (image)
  int seq_kv;     // SELF: keys per unit; CROSS: L
  int q_tiles;    // CTAs per (unit, head) = ceil(seq_q / 256)
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

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_flash_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                                       // [kQTiles]
  uint8_t* sK = sQ + kQTiles * kTileBytes;                  // [kKV]
  uint8_t* sV = sK + kKV * kTileBytes;                      // [kKV]
  uint8_t* sP = sV + kKV * kTileBytes;                      // [kQTiles] x (2 K-blocks of 64 keys)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kQTiles * kTileBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;              // [kKV]
  uint64_t* v_full = k_full + kKV;          // [kKV]
  uint64_t* k_empty = v_full + kKV;         // [kKV]
  uint64_t* v_empty = k_empty + kKV;        // [kKV]
  uint64_t* s_full = v_empty + kKV;         // [kQTiles]  S_t = Q_t K_j^T is in TMEM
  uint64_t* p_ready = s_full + kQTiles;     // [kQTiles]  P_t written (and S_t consumed) by the softmax warps of tile t
  uint64_t* o_done = p_ready + kQTiles;     // [kQTiles]  O_t += P_t V_j complete
  uint64_t* s_free = o_done + kQTiles;      // [kQTiles]  S_t copied to registers: the next Q_t K^T may overwrite it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + kQTiles);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  // ---- which tiles am I
  const int qt = blockIdx.x;
  const int head = blockIdx.y;
  const int unit = blockIdx.z;
  const int q_row0 = unit * g.seq_q + qt * (kQT * kQTiles);  // global query row of Q tile 0, row 0
  const int q_left = g.seq_q - qt * (kQT * kQTiles);          // valid query rows in this CTA (> 0)
  const int n_qt = q_left > kQT ? 2 : 1;                      // Q tiles that carry at least one valid row
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
    for (int i = 0; i < kKV; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < kQTiles; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&o_done[i], 1);
      mbar_init(&s_free[i], 128);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S_0 [0,128)  S_1 [128,256)  O_0 [256,320)  O_1 [320,384)

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, n_qt * kTileBytes);
      for (int t = 0; t < n_qt; ++t) tma_load_2d(sQ + t * kTileBytes, &tmQ, q_full, head * kD, q_row0 + t * kQT);
    }
    for (int j = 0; j < n_kv_tiles; ++j) {
      const int st = j % kKV;
      const uint32_t ph = (j / kKV) & 1;
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
    // Per KV tile j and Q tile t:  O_t += P_t(j) V_j  as soon as the softmax warps of tile t publish P_t(j), then
    // S_t = Q_t K_{j+1}^T straight away (S_t was consumed before P_t was published). While tile t waits for these
    // MMAs the other tile's softmax keeps the MUFU busy; both tiles share every K/V stage.
    const uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    if (lane == 0) {
      const uint64_t k_desc = make_sdesc_sw128(smem_u32(sK), 16, 1024);
      for (int t = 0; t < n_qt; ++t) {
        const uint64_t q_desc = make_sdesc_sw128(smem_u32(sQ + t * kTileBytes), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_ss(tmem_base + t * 128, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k != 0);
        tc_commit(&s_full[t]);
      }
      tc_commit(&k_empty[0]);
    }
    __syncwarp();
    for (int j = 0; j < n_kv_tiles; ++j) {
      const int st = j % kKV;
      const uint32_t ph = (j / kKV) & 1;
      const int st2 = (j + 1) % kKV;
      const uint32_t ph2 = ((j + 1) / kKV) & 1;
      // (a) S_t = Q_t K_{j+1}^T as soon as the softmax warps hold S_t(j) in registers: it is ready long before they
      //     finish exponentiating tile j, so the softmax chain never waits for the tensor core
      if (j + 1 < n_kv_tiles) {
        mbar_wait(&k_full[st2], ph2);
        for (int t = 0; t < n_qt; ++t) {
          mbar_wait(&s_free[t], j & 1);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t q_desc = make_sdesc_sw128(smem_u32(sQ + t * kTileBytes), 16, 1024);
            const uint64_t k_desc = make_sdesc_sw128(smem_u32(sK + st2 * kTileBytes), 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_ss(tmem_base + t * 128, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k != 0);
            tc_commit(&s_full[t]);
            if (t == n_qt - 1) tc_commit(&k_empty[st2]);
          }
          __syncwarp();
        }
      }
      // (b) O_t += P_t(j) V_j
      mbar_wait(&v_full[st], ph);
      for (int t = 0; t < n_qt; ++t) {
        mbar_wait(&p_ready[t], j & 1);
        tc_fence_after();
        if (lane == 0) {
          // P_t (128 x 128, K-major, two 64-key blocks) * V_j (128 keys x 64, MN-major: 8-key groups 1024 B apart)
          const uint32_t sp = smem_u32(sP + t * 2 * kTileBytes);
          const uint32_t sv = smem_u32(sV + st * kTileBytes);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t p_desc = make_sdesc_sw128(sp + (k >> 2) * kTileBytes + (k & 3) * 32, 16, 1024);
            const uint64_t v_desc = make_sdesc_sw128(sv + k * 2048, 16, 1024);
            tc_mma_ss(tmem_base + 256 + t * 64, p_desc, v_desc, idesc_pv, (j | k) != 0);
          }
          tc_commit(&o_done[t]);
          if (t == n_qt - 1) tc_commit(&v_empty[st]);
        }
        __syncwarp();
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ------------------------------------------------------------------ softmax + output (one thread per query row)
    const int t = (warp - 4) >> 2;  // Q tile of this warp
    const int qd = warp & 3;        // TMEM lane quarter
    const int r = qd * 32 + lane;   // query row within the Q tile == TMEM lane
    if (t < n_qt) {
      const uint32_t lane_addr = uint32_t(qd * 32) << 16;
      const uint32_t tmem_S = tmem_base + t * 128 + lane_addr;
      const uint32_t tmem_O = tmem_base + 256 + t * 64 + lane_addr;
      const int q_valid = min(kQT, q_left - t * kQT);
      int my_ctx = -1;
      if (g.kv_mode == KV_CROSS_TEMPORAL) {
        // global row -> (b, f, s); temporal batch row (b, s) reads context (b*S + s) mod n_ctx  [reference quirk]
        const long long row = (long long)q_row0 + t * kQT + r;
        const int s = (int)(row % g.S);
        const int b = (int)(row / ((long long)g.F * g.S)) + g.batch_offset;
        my_ctx = (int)(((long long)b * g.S + s) % g.n_ctx);
      }
      const float c2 = g.scale_log2;
      float m_used = -INFINITY;  // (stale) row max the exponentials are taken against
      float l_run = 0.f;
      uint8_t* const prow0 = sP + t * 2 * kTileBytes + r * 128;
      const int rx = r & 7;
      for (int j = 0; j < n_kv_tiles; ++j) {
        int kv_valid = kKT;
        if (g.kv_mode == KV_SELF) kv_valid = min(kKT, g.seq_kv - j * kKT);
        else kv_valid = g.seq_kv;
        const bool row_off = (g.kv_mode == KV_CROSS_TEMPORAL) && (my_ctx != j);
        const bool masked = (kv_valid < kKT) || (g.kv_mode == KV_CROSS_TEMPORAL);  // CTA-uniform
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        // ---- the whole S row (128 fp32) goes to registers: four tcgen05.ld in flight, one wait (~1 TMEM latency)
        uint32_t v[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32(tmem_S + c * 32, v[c]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[t]);  // S_t(j) is in registers
        if (masked) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (row_off || c * 32 + i >= kv_valid) v[c][i] = 0xff800000u;  // -inf
        }
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1])));
        // ---- lazy rescale: only move the reference max when it grows by more than 2^8 (exp2 domain)
        const bool grow = (mx - m_used) * c2 > 8.0f;  // j == 0: m_used = -inf -> true (NaN if both -inf -> false)
        const float m_new = grow ? mx : m_used;
        if (j > 0) {
          // P_t(j-1) V_{j-1} must be complete before P_t is overwritten and before O_t is rescaled
          mbar_wait(&o_done[t], (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, grow)) {
            const float alpha = grow ? ex2((m_used - m_new) * c2) : 1.0f;  // m_used = -inf -> 0 (row still empty)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t o[32];
              tmem_ld_32x32(tmem_O + c * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x32(tmem_O + c * 32, o);
            }
            tmem_st_wait();
            l_run *= alpha;
          }
        }
        m_used = m_new;
        const float ms = (m_used == -INFINITY) ? 0.f : m_used * c2;
        // ---- probabilities -> 128B-swizzled smem tile (bf16), row sum
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2(fmaf(__uint_as_float(v[c][i]), c2, -ms));  // exp2(-inf) = 0 for masked keys
            const float p1 = ex2(fmaf(__uint_as_float(v[c][i + 1]), c2, -ms));
            sum0 += p0;
            sum1 += p1;
            pk[i >> 1] = pack_bf16(p0, p1);
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
        mbar_arrive(&p_ready[t]);
      }
      // ---- epilogue: O / l
      mbar_wait(&o_done[t], (n_kv_tiles - 1) & 1);
      tc_fence_after();
      const float inv = (l_run > 0.f) ? 1.f / l_run : 0.f;
      __nv_bfloat16* orow = g.out + (long long)(q_row0 + t * kQT + r) * g.ldo + head * kD;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_O + c * 32, v);
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
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
  g.q_tiles = (p->seq + kQT * kQTiles - 1) / (kQT * kQTiles);
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
  g.q_tiles = (p->S + kQT * kQTiles - 1) / (kQT * kQTiles);
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
