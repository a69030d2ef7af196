// K5 — flash-style spatial self-attention on tcgen05 for head_dim 64.
//
//   S = Q K^T   : tcgen05.mma (SS) M=128 N=128 K=16 x 4, Q/K tiles staged by TMA (128B swizzle), S_t in TMEM
//   softmax     : 128 threads per Q tile (one query row each) pull the whole S row into registers with four
//                 tcgen05.ld in flight, online max/sum in fp32 (four chains of 3-input max), and write P (bf16 pairs)
//                 straight back into TMEM with tcgen05.st — P never touches shared memory
//   O += P V    : tcgen05.mma (TS: A = P from TMEM, B = the V tile in place as MN-major smem operand) M=128 N=64 K=16
//                 x 8, O_t accumulates in TMEM across KV tiles; the reference max is only moved when it grows by more
//                 than 2^8 (lazy rescale), in which case the softmax warps scale their O rows in TMEM.
//
//   warp 0: TMA producer   warp 1 / 2: MMA issuer of Q tile 0 / 1 (warp 2 also owns the TMEM allocation)
//   warps 4..7 / 8..11: softmax of Q tile 0 / 1
//
// Each CTA owns TWO 128-row Q tiles that share every K/V stage (half the K/V traffic per FLOP) and are otherwise
// independent: own S / O / P columns, own barriers, own issuer warp. Their chains
// (S_t(j) -> registers -> [S_t = Q_t K_{j+1}^T] -> exp2 -> P_t(j) -> [O_t += P_t(j) V_j]) overlap freely, so one tile's
// MUFU-free phases (TMEM round trips, max, publishing P) can fall under the other tile's exponentials.
// What bounds the kernel (measured with the clock64 event trace behind -DTTVDM_ATTN_TRACE, tools/attn_trace.py, and
// tools/microbench/mma_rate.cu / mma_group.cu):
//   * MUFU.EX2: 16/clk/SM = 1024 cycles per 128 x 128 tile;
//   * the tensor pipe runs one accumulation chain at a time and every chain pays ~330 cycles of latency: 4 x N128
//     (Q K^T) = 507 cycles, 8 x N64 (P V) = 652 — 2318 cycles per KV tile for both Q tiles, although the MMAs themselves
//     are only 1024 cycles of work. The kernel needs ~2550 cycles per KV tile: within 10 % of the chain bound;
//   * the ISSUING THREAD bounded round 1's kernel: (a) under `if (lane == 0)` ptxas wraps every tcgen05.mma in an
//     ELECT / 4 x R2UR / branch sequence (30-60 cycles per MMA) — under elect.sync the descriptors stay in uniform
//     registers and MMAs issue back to back; (b) one thread is slow at everything else (an mbarrier wait is a
//     ~130-cycle shared-memory round trip, scalar code runs at one instruction per 6-10 cycles next to two busy softmax
//     warps), so one issuer serving both tiles spent ~800 cycles per chain and forced the tiles into a fixed order;
//     one issuer per tile with a fixed sequence of blocking waits removed that; (c) P through shared memory cost
//     64 KB of smem writes + reads per KV tile.
//   Tried and measured slower on the same box: a polling scheduler over both tiles, 256-key KV tiles with two key
//   groups per row (shared or split accumulators), refilling S registers under the exponentials (software pipelining),
//   a forced half-phase skew between the tiles, and moving a quarter of the exp2 to an FMA-pipe polynomial.
// 192 KB smem (two Q buffers + 4-deep K/V ring), all 512 TMEM columns, one persistent CTA per SM (see the kernel).
// Cross-attention against the 77-token CLIP context has its own kernel (attn_cross.cu): one small K/V tile per
// (context, head) does not need this machinery.
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace ttvdm {

constexpr int kQT = 128;      // query rows per Q tile
constexpr int kQTiles = 2;    // Q tiles per CTA (they share every K/V tile)
constexpr int kKT = 128;      // keys per KV tile
constexpr int kD = 64;        // head dim
constexpr int kKV = 4;        // K/V ring depth
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KB
constexpr int kAttnThreads = 128 + 128 * kQTiles;  // 4 service warps + 4 softmax warps per Q tile
constexpr int kAttnPolyDefault = 13;  // 3 of every 8 column pairs, spread (TTVDM_ATTN_POLY: 0..4 = first n of 8, 12..14 = n - 10 spread)
constexpr int kAttnSmem = kTileBytes * (2 * kQTiles + 2 * kKV) + 512;
// TMEM columns: S_0 [0,128)  S_1 [128,256)  O_0 [256,320)  O_1 [320,384)  P_0 [384,448)  P_1 [448,512)
constexpr uint32_t kColS = 0, kColO = 256, kColP = 384;


struct AttnArgs {
  int seq_q;      // query rows per unit (image)
  int seq_kv;     // keys per unit (= seq_q: self-attention)
  int q_tiles;    // work items per (unit, head) = ceil(seq_q / 256)
  int heads;
  int units;      // images
  float scale_log2;  // scale * log2(e)
  __nv_bfloat16* out;
  int ldo;
  // kCross only (cross-attention against one <= 128-key context tile, include/ttvdm.h ttvdm_xattn_params): rows are
  // addressed as (unit = (b_local, f), s); a work item takes 256 rows s = s0 + q_stride * i of one unit
  int q_stride;      // 1 (spatial) or n_ctx (temporal: the rows of a unit that read one context)
  int S, F;          // rows per unit in memory, frames per batch element
  int n_units;       // b_local * F
  int n_ctx, batch_offset, temporal;
#ifdef TTVDM_ATTN_TRACE
  long long* trace;  // [3 classes][kTraceCap] (clock << 8 | tag) event log of one CTA (debug builds only)
#endif
};

#ifdef TTVDM_ATTN_TRACE
constexpr int kTraceCap = 4096;
#define TR_DECL(cls, on) long long* tr_p = (g.trace && (on) && blockIdx.x == 3) ? g.trace + (cls) * kTraceCap : nullptr; int tr_n = 0
#define TR(tag) do { if (tr_p && tr_n < kTraceCap) tr_p[tr_n++] = (clock64() << 8) | (tag); } while (0)
#else
#define TR_DECL(cls, on)
#define TR(tag)
#endif

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 of two exponents (x <= ~8) WITHOUT the MUFU: Cody-Waite split on the FMA pipe in packed f32x2 arithmetic.
// t = x + 1.5 * 2^23 leaves n = round(x) in the low mantissa bits of t; f = x - n in [-0.5, 0.5]; 2^f by a degree-3
// minimax polynomial (relative error 7.5e-5, far below the 2^-9 of the bf16 rounding P gets anyway); the result is
// the polynomial's bit pattern with n added to the exponent field (one IMAD per element). x is clamped at -126 first
// (masked keys are -inf; a finite key may sit hundreds of units below the row maximum).
__device__ __forceinline__ void exp2_poly_pair(float x0, float x1, float& r0, float& r1) {
  const uint64_t x2 = pack_f32x2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const uint64_t t = add_f32x2(x2, pack_f32x2(12582912.f, 12582912.f));
  const uint64_t n = add_f32x2(t, pack_f32x2(-12582912.f, -12582912.f));
  const uint64_t f = fma_f32x2(n, pack_f32x2(-1.f, -1.f), x2);
  uint64_t q = fma_f32x2(f, pack_f32x2(0.055171459913253784f, 0.055171459913253784f),
                         pack_f32x2(0.2426108568906784f, 0.2426108568906784f));
  q = fma_f32x2(q, f, pack_f32x2(0.6932609677314758f, 0.6932609677314758f));
  q = fma_f32x2(q, f, pack_f32x2(0.9999281167984009f, 0.9999281167984009f));
  float q0, q1, t0, t1;
  unpack_f32x2(q, q0, q1);
  unpack_f32x2(t, t0, t1);
  r0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  r1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// kPoly of every 8 column pairs take their exponential on the FMA pipe (exp2_poly_pair), the rest on the MUFU: the
// softmax of two 128 x 128 tiles is 2048 MUFU cycles per KV tile on one SM (16 results per clock) against 1024 cycles of
// tensor-pipe work, so the MUFU is what bounds the kernel at d = 64 unless part of the exponentials leaves it.
// kCross: cross-attention against the short CLIP context (K6 / K8) on the same pipeline. The whole K / V of a
// (context, head) is ONE ragged KV tile (L <= 128 keys: the masked path), the Q tiles of a work item are 2 x 128 rows of
// one (b, f) unit fetched by a 3-D TMA box whose row dimension has element stride q_stride — for the temporal layers
// that gathers exactly the rows s = s0 + n_ctx * i that read this context (the reference's quirk,
// svd/diffusion_arch/transformer_temporal.py:310-319), so nothing is masked or computed twice.
template <int kPoly, bool kCross, bool kSpread = false>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_flash_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                                       // [2 buffers][kQTiles]
  uint8_t* sK = sQ + 2 * kQTiles * kTileBytes;              // [kKV]
  uint8_t* sV = sK + kKV * kTileBytes;                      // [kKV]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kKV * kTileBytes);
  uint64_t* q_full = bars + 0;              // [2]        the Q tiles of a work item have landed
  uint64_t* q_empty = q_full + 2;           // [2]        both issuers are done with the Q buffer
  uint64_t* k_full = q_empty + 2;           // [kKV]
  uint64_t* v_full = k_full + kKV;          // [kKV]
  uint64_t* k_empty = v_full + kKV;         // [kKV]      released by the issuer of every Q tile
  uint64_t* v_empty = k_empty + kKV;        // [kKV]
  uint64_t* s_full = v_empty + kKV;         // [kQTiles]  S_t = Q_t K_j^T is in TMEM
  uint64_t* p_ready = s_full + kQTiles;     // [kQTiles]  P_t written by the softmax warps of tile t
  uint64_t* o_done = p_ready + kQTiles;     // [kQTiles]  O_t += P_t V_j complete
  uint64_t* s_free = o_done + kQTiles;      // [kQTiles]  S_t copied to registers: the next Q_t K^T may overwrite it
  uint64_t* o_free = s_free + kQTiles;      // [kQTiles]  O_t of the finished work item is in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + kQTiles);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  // ---- persistent CTA: work item w = (q tile pair, head, unit), w = blockIdx.x, blockIdx.x + gridDim.x, ...
  // Barriers, TMEM and the pipeline state live across work items (all parities come from running counters), so the
  // next item's Q/K/V loads and its first Q K^T overlap the previous item's last P V and its output stores. This is
  // what keeps the low-resolution levels (few KV tiles per item) efficient: with one CTA per item they paid ~5 us of
  // set-up and exposed latency per item.
  const int total_works = g.q_tiles * g.heads * g.units;
  struct Work {
    long long q_row0;  // global query row of Q tile 0, row 0 (row i of the item is q_row0 + i * q_stride)
    int q_left, head, n_kv_tiles, kv_row_base;
    int unit, s_first;  // kCross: TMA coordinates of the item's first row
  };
  auto decode = [&](int w) -> Work {
    Work k;
    const int qt = w % g.q_tiles;
    const int hu = w / g.q_tiles;
    k.head = hu % g.heads;
    if (!kCross) {
      const int unit = hu / g.heads;
      k.q_row0 = (long long)unit * g.seq_q + qt * (kQT * kQTiles);
      k.q_left = g.seq_q - qt * (kQT * kQTiles);  // valid query rows of this item (> 0)
      k.n_kv_tiles = (g.seq_kv + kKT - 1) / kKT;
      k.kv_row_base = unit * g.seq_kv;
      k.unit = unit;
      k.s_first = 0;
    } else {
      // units = n_units (spatial: context = global batch index of the unit) or n_ctx * n_units (temporal)
      const int uc = hu / g.heads;
      const int unit = g.temporal ? uc % g.n_units : uc;
      const int b_glob = g.batch_offset + unit / g.F;
      const int ctx = g.temporal ? uc / g.n_units : b_glob;
      int s0 = 0;
      if (g.temporal) s0 = (int)((((long long)ctx - (long long)b_glob * g.S) % g.n_ctx + g.n_ctx) % g.n_ctx);
      const int rows = (g.S - s0 + g.q_stride - 1) / g.q_stride;  // rows of this unit that read this context
      k.unit = unit;
      k.s_first = s0 + g.q_stride * qt * (kQT * kQTiles);
      k.q_row0 = (long long)unit * g.S + k.s_first;
      k.q_left = rows - qt * (kQT * kQTiles);  // may be <= 0 for the last item of a unit with s0 > 0
      k.n_kv_tiles = 1;
      k.kv_row_base = ctx * g.seq_kv;
    }
    return k;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], kQTiles);
    }
    for (int i = 0; i < kKV; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], kQTiles);
      mbar_init(&v_empty[i], kQTiles);
    }
    for (int i = 0; i < kQTiles; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&o_done[i], 1);
      mbar_init(&s_free[i], 128);
      mbar_init(&o_free[i], 128);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected thread)
    if (elect_one()) {
      uint32_t kt = 0;  // K/V tiles loaded so far (ring position)
      int wi = 0;
      for (int w = blockIdx.x; w < total_works; w += gridDim.x, ++wi) {
        const Work k = decode(w);
        const int qb = wi & 1;
        mbar_wait(&q_empty[qb], ((wi >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[qb], kQTiles * kTileBytes);
        // both Q tiles are always loaded: a tile past the item's rows holds the next image's rows or TMA zero fill
        // and is computed but never stored
        for (int t = 0; t < kQTiles; ++t) {
          if (!kCross) tma_load_2d(sQ + (qb * kQTiles + t) * kTileBytes, &tmQ, &q_full[qb], k.head * kD, (int)k.q_row0 + t * kQT);
          else tma_load_3d(sQ + (qb * kQTiles + t) * kTileBytes, &tmQ, &q_full[qb], k.head * kD, k.s_first + t * kQT * g.q_stride, k.unit);
        }
        for (int j = 0; j < k.n_kv_tiles; ++j, ++kt) {
          const int st = kt % kKV;
          const uint32_t ph = (kt / kKV) & 1;
          const int kv_row = k.kv_row_base + j * kKT;
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_expect_tx(&k_full[st], kTileBytes);
          tma_load_2d(sK + st * kTileBytes, &tmK, &k_full[st], k.head * kD, kv_row);
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_expect_tx(&v_full[st], kTileBytes);
          tma_load_2d(sV + st * kTileBytes, &tmV, &v_full[st], k.head * kD, kv_row);
        }
      }
    }
    __syncwarp();
  } else if (warp - 1 < kQTiles) {
    // ------------------------------------------------------------------ MMA issuers: warp 1 + t drives Q tile t
    // Per tile the events come in a fixed order (S_t(j) pulled into registers -> P_t(j) published), so each issuer runs
    // a fixed sequence with blocking waits: S_t = Q_t K_{j+1}^T as soon as S_t(j) is in registers, O_t += P_t(j) V_j as
    // soon as P_t(j) is published. The issuing thread is chosen with elect.sync, NOT `lane == 0` (see the header).
    const int t = warp - 1;
    if (elect_one()) {
      TR_DECL(0, t == 0);
      const uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major
      const uint32_t d_s = tmem_base + kColS + t * 128, d_o = tmem_base + kColO + t * 64, a_p = tmem_base + kColP + t * 64;
      uint32_t kt = 0;  // K/V tiles consumed before this work item (ring position)
      uint32_t c = 0;   // KV tiles this Q tile has been through (parity of s_full / s_free / p_ready / o_done)
      int wi = 0;
      for (int w = blockIdx.x; w < total_works; w += gridDim.x, ++wi) {
        const Work k = decode(w);
        const int qb = wi & 1;
        const uint64_t q_desc = make_sdesc_sw128(smem_u32(sQ + (qb * kQTiles + t) * kTileBytes), 16, 1024);
        auto issue_qk = [&](uint32_t tile) {
          const int st = tile % kKV;
          const uint64_t k_desc = make_sdesc_sw128(smem_u32(sK + st * kTileBytes), 16, 1024);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma_ss(d_s, q_desc + 2 * kk, k_desc + 2 * kk, idesc_qk, kk != 0);
          tc_commit(&s_full[t]);
          tc_commit(&k_empty[st]);  // one of kQTiles arrivals: this tile has read K
        };
        mbar_wait(&q_full[qb], (wi >> 1) & 1);
        mbar_wait(&k_full[kt % kKV], (kt / kKV) & 1);
        if (c > 0) mbar_wait(&s_free[t], (c - 1) & 1);  // S_t of the previous item's last tile is in registers
        tc_fence_after();
        issue_qk(kt);
        if (k.n_kv_tiles == 1) tc_commit(&q_empty[qb]);
        else mbar_wait(&k_full[(kt + 1) % kKV], ((kt + 1) / kKV) & 1);
        for (int j = 0; j < k.n_kv_tiles; ++j) {
          const uint32_t tile = kt + j;
          const int st = tile % kKV;
          if (j + 1 < k.n_kv_tiles) {
            mbar_wait(&s_free[t], (c + j) & 1);  // the softmax warps hold S_t(j) in registers (K_{j+1} has landed)
            tc_fence_after();
            TR(0x10 + t);
            issue_qk(tile + 1);
            if (j + 2 == k.n_kv_tiles) tc_commit(&q_empty[qb]);  // last Q K^T of the item: one of kQTiles arrivals
            TR(0x18 + t);
          }
          mbar_wait(&v_full[st], (tile / kKV) & 1);  // landed long ago; this round trip hides under the softmax
          mbar_wait(&p_ready[t], (c + j) & 1);
          if (j == 0 && wi > 0) mbar_wait(&o_free[t], (wi - 1) & 1);  // the previous item's O_t has been read out
          tc_fence_after();
          TR(0x20 + t);
          // P_t (128 x 128 bf16 pairs in TMEM) * V_j (128 keys x 64, MN-major: 8-key groups 1024 B apart)
          const uint32_t sv = smem_u32(sV + st * kTileBytes);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            tc_mma_ts(d_o, a_p + kk * 8, make_sdesc_sw128(sv + kk * 2048, 16, 1024), idesc_pv, (j | kk) != 0);
          tc_commit(&o_done[t]);
          tc_commit(&v_empty[st]);  // one of kQTiles arrivals
          TR(0x28 + t);
          if (j + 2 < k.n_kv_tiles) mbar_wait(&k_full[(tile + 2) % kKV], ((tile + 2) / kKV) & 1);
        }
        kt += k.n_kv_tiles;
        c += k.n_kv_tiles;
      }
    }
    __syncwarp();
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ------------------------------------------------------------------ softmax + output (one thread per query row)
    const int t = (warp - 4) >> 2;  // Q tile of this warp
    const int qd = warp & 3;        // TMEM lane quarter
    const int r = qd * 32 + lane;   // query row within the Q tile == TMEM lane
    const uint32_t lane_addr = uint32_t(qd * 32) << 16;
    const uint32_t tmem_S = tmem_base + kColS + t * 128 + lane_addr;
    const uint32_t tmem_O = tmem_base + kColO + t * 64 + lane_addr;
    const uint32_t tmem_P = tmem_base + kColP + t * 64 + lane_addr;
    const float c2 = g.scale_log2;
    uint32_t c = 0;  // KV tiles this Q tile has been through
    TR_DECL(1 + t, (threadIdx.x & 127) == 0);
    for (int w = blockIdx.x; w < total_works; w += gridDim.x) {
      const Work k = decode(w);
      const int q_valid = min(kQT, k.q_left - t * kQT);  // <= 0: this tile carries no row of the item
      float m_used = -INFINITY;  // (stale) row max the exponentials are taken against
      float l_run = 0.f;
      for (int j = 0; j < k.n_kv_tiles; ++j, ++c) {
        const int kv_valid = min(kKT, g.seq_kv - j * kKT);
        const bool masked = kv_valid < kKT;  // CTA-uniform: ragged last KV tile
        TR(1);
        mbar_wait(&s_full[t], c & 1);
        tc_fence_after();
        TR(2);
        // ---- the whole S row (128 fp32) goes to registers: four tcgen05.ld in flight, one wait (~1 TMEM latency)
        uint32_t v[4][32];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) tmem_ld_32x32(tmem_S + cc * 32, v[cc]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[t]);  // S_t(j) is in registers: the next Q_t K^T may overwrite it
        TR(3);
        if (masked) {
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (cc * 32 + i >= kv_valid) v[cc][i] = 0xff800000u;  // -inf
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent chains of 3-input max
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            mx4[(i >> 1) & 3] = fmax3(mx4[(i >> 1) & 3], __uint_as_float(v[cc][i]), __uint_as_float(v[cc][i + 1]));
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // ---- lazy rescale: only move the reference max when it grows by more than 2^8 (exp2 domain)
        const bool grow = (mx - m_used) * c2 > 8.0f;  // j == 0: m_used = -inf -> true (NaN if both -inf -> false)
        const float m_new = grow ? mx : m_used;
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
          // P_t(j-1) V_{j-1} must be complete before O_t is rescaled
          mbar_wait(&o_done[t], (c - 1) & 1);
          tc_fence_after();
          const float alpha = grow ? ex2((m_used - m_new) * c2) : 1.0f;  // m_used = -inf -> 0 (row still empty)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t o[32];
            tmem_ld_32x32(tmem_O + cc * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(tmem_O + cc * 32, o);
          }
          tmem_st_wait();
          l_run *= alpha;
        }
        m_used = m_new;
        const float ms = (m_used == -INFINITY) ? 0.f : m_used * c2;
        TR(4);
        // ---- probabilities: bf16 pairs (P column c = keys 2c, 2c+1), row sum (packed f32x2 scale / sum)
        const uint64_t c22 = pack_f32x2(c2, c2), nms2 = pack_f32x2(-ms, -ms);
        uint64_t sum_a = pack_f32x2(0.f, 0.f), sum_b = sum_a;
        uint32_t pk[2][32];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float x0, x1, p0, p1;
            unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(v[cc][i]), __uint_as_float(v[cc][i + 1])), c22, nms2), x0, x1);
            // a ragged (masked) KV tile is rare and CTA-uniform: it keeps every exponential on the MUFU
            // kSpread: the kPoly polynomial pairs of every 8 are spread evenly (Bresenham: 3 of 8 = pairs 0, 3, 6) instead of
            // being the first kPoly — MUFU and FMA-pipe work alternate at a finer grain in the instruction stream
            if ((kSpread ? ((((i >> 1) & 7) * kPoly) & 7) : ((i >> 1) & 7)) < kPoly && !masked) {
              exp2_poly_pair(x0, x1, p0, p1);
            } else {
              p0 = ex2(x0);  // exp2(-inf) = 0 for masked keys
              p1 = ex2(x1);
            }
            const uint64_t p2 = pack_f32x2(p0, p1);
            if (i & 2) sum_b = add_f32x2(sum_b, p2); else sum_a = add_f32x2(sum_a, p2);
            pk[cc >> 1][(cc & 1) * 16 + (i >> 1)] = pack_bf16(p0, p1);
          }
        float sum0, sum1;
        unpack_f32x2(add_f32x2(sum_a, sum_b), sum0, sum1);
        l_run += sum0 + sum1;
        TR(5);
        if (j > 0) {
          // the tensor core must have finished reading P_t(j-1) (issued one tile ago); for j == 0 the previous item's
          // epilogue below has already waited for its last P V
          mbar_wait(&o_done[t], (c - 1) & 1);
          tc_fence_after();
        }
        tmem_st_32x32(tmem_P, pk[0]);
        tmem_st_32x32(tmem_P + 32, pk[1]);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_ready[t]);
        TR(6);
      }
      // ---- epilogue: O / l
      mbar_wait(&o_done[t], (c - 1) & 1);
      tc_fence_after();
      const float inv = (l_run > 0.f) ? 1.f / l_run : 0.f;
      uint32_t o[2][32];
      tmem_ld_32x32(tmem_O, o[0]);
      tmem_ld_32x32(tmem_O + 32, o[1]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&o_free[t]);  // O_t is in registers: the next item's first P V may overwrite it
      if (r < q_valid) {
        __nv_bfloat16* orow = g.out + (k.q_row0 + (long long)(t * kQT + r) * (kCross ? g.q_stride : 1)) * g.ldo + k.head * kD;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint32_t w4[4];
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2)
              w4[k2] = pack_bf16(__uint_as_float(o[cc][i * 8 + k2 * 2]) * inv, __uint_as_float(o[cc][i * 8 + k2 * 2 + 1]) * inv);
            *(reinterpret_cast<uint4*>(orow + cc * 32) + i) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

static int launch_attn(const void* q, int ldq, long long q_rows, const void* k, int ldk, const void* v, int ldv,
                       long long kv_rows, const AttnArgs& g, int units, cudaStream_t stream, bool cross = false) {
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  const uint32_t box[2] = {kD, 128};
  if (!cross) {
    uint64_t dims[2] = {(uint64_t)g.heads * kD, (uint64_t)q_rows};
    uint64_t str[1] = {(uint64_t)ldq * 2};
    if ((rc = make_tmap_bf16(&tmQ, q, 2, dims, str, box))) return rc;
  } else {
    // (channel, s, unit) with element stride q_stride along s: a box of 128 * q_stride positions delivers 128 rows
    uint64_t dims[3] = {(uint64_t)g.heads * kD, (uint64_t)g.S, (uint64_t)g.n_units};
    uint64_t str[2] = {(uint64_t)ldq * 2, (uint64_t)g.S * ldq * 2};
    uint32_t box3[3] = {kD, (uint32_t)(128 * g.q_stride), 1};
    uint32_t est[3] = {1, (uint32_t)g.q_stride, 1};
    if ((rc = make_tmap_bf16(&tmQ, q, 3, dims, str, box3, true, est))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)g.heads * kD, (uint64_t)kv_rows};
    uint64_t strk[1] = {(uint64_t)ldk * 2};
    uint64_t strv[1] = {(uint64_t)ldv * 2};
    if ((rc = make_tmap_bf16(&tmK, k, 2, dims, strk, box))) return rc;
    if ((rc = make_tmap_bf16(&tmV, v, 2, dims, strv, box))) return rc;
  }
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnArgs);
  static const Kern kerns[9] = {attn_flash_kernel<0, false>, attn_flash_kernel<1, false>, attn_flash_kernel<2, false>,
                                attn_flash_kernel<3, false>, attn_flash_kernel<4, false>, attn_flash_kernel<0, true>,
                                attn_flash_kernel<2, false, true>, attn_flash_kernel<3, false, true>,
                                attn_flash_kernel<4, false, true>};
  static int poly = -1;  // column pairs of every 8 whose exponential runs on the FMA pipe (TTVDM_ATTN_POLY: A/B runs)
  if (poly < 0) {
    const char* e = getenv("TTVDM_ATTN_POLY");
    int v = e ? atoi(e) : kAttnPolyDefault;  // 0..4: first v pairs of every 8; 12..14: v - 10 pairs spread over every 8
    if (!((v >= 0 && v <= 4) || (v >= 12 && v <= 14))) v = kAttnPolyDefault;
    if (v >= 12) v = v - 12 + 6;  // index into kerns[]
    for (int i = 0; i < 9; ++i) {
      cudaError_t ce = cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
      if (ce != cudaSuccess) return fail(TTVDM_ERR_CUDA, "attn: cudaFuncSetAttribute: %s", cudaGetErrorString(ce));
    }
    poly = v;
  }
  AttnArgs ga = g;
  ga.units = units;
  const long long works = (long long)g.q_tiles * g.heads * units;
  if (works > 0x7fffffffLL) return fail(TTVDM_ERR_SHAPE, "attn: too many work items");
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = (int)(works < n_sm ? works : n_sm);  // persistent: one CTA per SM
  kerns[cross ? 5 : poly]<<<grid, kAttnThreads, kAttnSmem, stream>>>(tmQ, tmK, tmV, ga);
  TTVDM_CHECK_LAUNCH("attn_flash_kernel");
  return 0;
}

int launch_attn_cross_tc(const ttvdm_xattn_params* p, cudaStream_t stream) {
  const int st = p->temporal ? p->n_ctx : 1;
  if (st > 2 || p->L > kKT) return -1;  // a 256-position TMA box holds 128 rows only up to stride 2
  // Work items are 256 rows of ONE (b, f) unit: below ~2 items per unit the padding of the last item and the per-item
  // pipeline latency outweigh the tensor-core path (measured on B200, tools/xattn_ab.py: S = 576 temporal 0.068 vs
  // 0.057 ms for the warp-level kernel, S = 576 spatial 0.047 vs 0.051, S >= 2304 25-30 % faster)
  if ((p->S + st - 1) / st < 512) return -1;
  const int b_local = p->rows / (p->F * p->S);
  AttnArgs g{};
  g.seq_q = 0;
  g.seq_kv = p->L;
  g.q_stride = st;
  g.S = p->S;
  g.F = p->F;
  g.n_units = b_local * p->F;
  g.n_ctx = p->n_ctx;
  g.batch_offset = p->batch_offset;
  g.temporal = p->temporal ? 1 : 0;
  const int rows_per_unit = (p->S + st - 1) / st;
  g.q_tiles = (rows_per_unit + kQT * kQTiles - 1) / (kQT * kQTiles);
  g.heads = p->heads;
  g.scale_log2 = p->scale * 1.4426950408889634f;
  g.out = static_cast<__nv_bfloat16*>(p->out);
  g.ldo = p->ldo;
#ifdef TTVDM_ATTN_TRACE
  g.trace = nullptr;
#endif
  const int ldkv = p->heads * kD;
  const int units = g.temporal ? p->n_ctx * g.n_units : g.n_units;
  return launch_attn(p->q, p->ldq, p->rows, p->kc, ldkv, p->vc, ldkv, (long long)p->n_ctx * p->L, g, units, stream, true);
}

}  // namespace ttvdm

using namespace ttvdm;

extern "C" int ttvdm_attn_spatial(const ttvdm_attn_params* p, void* stream_) {
  if (int rc = ensure_init()) return rc;
  if (!p || !p->q || !p->k || !p->v || !p->out) return fail(TTVDM_ERR_SHAPE, "attn_spatial: null");
  if (p->n_img <= 0 || p->heads <= 0 || p->seq <= 0) return fail(TTVDM_ERR_SHAPE, "attn_spatial: empty");
  if ((p->ldq | p->ldk | p->ldv | p->ldo) % 8 != 0) return fail(TTVDM_ERR_SHAPE, "attn_spatial: ld %% 8 != 0");
  AttnArgs g{};
  g.seq_q = p->seq;
  g.seq_kv = p->seq;
  g.q_tiles = (p->seq + kQT * kQTiles - 1) / (kQT * kQTiles);
  g.heads = p->heads;
  g.scale_log2 = p->scale * 1.4426950408889634f;
  g.out = static_cast<__nv_bfloat16*>(p->out);
  g.ldo = p->ldo;
#ifdef TTVDM_ATTN_TRACE
  { const char* e = getenv("TTVDM_ATTN_TRACE"); g.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 10)) : nullptr; }
#endif
  const long long rows = (long long)p->n_img * p->seq;
  return launch_attn(p->q, p->ldq, rows, p->k, p->ldk, p->v, p->ldv, rows, g, p->n_img,
                     static_cast<cudaStream_t>(stream_));
}
