// Flash-style attention on tcgen05 for sm_100a: O = softmax(Q K^T * scale) V, no S x S matrix in HBM.
//
// One CTA = 128 queries of one (sample, head). Per 128-key block:
//   warp 0   TMA: K block [128 keys x d] and V^T block [d x 128 keys] into a 2-stage ring
//   warp 1   tcgen05.mma: S = Q K^T into TMEM (double buffered), then O += P V with P read from TMEM
//   warps 2-5  one query row per thread: tcgen05.ld S -> mask -> online softmax (exp2, fp32) ->
//              bf16 P written back over S with tcgen05.st; O is rescaled in TMEM only when the running
//              maximum grows by more than 2^8 (lazy rescale)
// Q and K are read in place from the projection output ([tokens, heads*d], any row stride) through
// 4-D tensor maps whose innermost extent is the head dim, so d = 40/80/160 needs no padding in HBM:
// the TMA zero-fills up to the 64-element swizzle atom. V arrives transposed ([heads*d, NB, Skv_pad])
// from a swapped-operand projection GEMM, which makes it a K-major B operand.
//
// Replaces sd/attention.py:55-76 (SelfAttention, incl. the causal mask of :58-62) and :219-234
// (CrossAttention).
#include "common.cuh"
#include "host.h"
#include "../../include/sdb200.h"

namespace sdb {

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 128;
constexpr int ATT_THREADS = 192;
constexpr int ATT_CHUNK_BYTES = 128 * 128;  // 128 rows x 64 bf16

struct AttnParams {
  CUtensorMap map_q;
  CUtensorMap map_k;
  CUtensorMap map_vt;
  __nv_bfloat16* out;
  long long ldo;
  int S, Skv, d, heads, NB;
  int causal;
  float scale_log2;
  int dchunks;   // ceil(d / 64)
  int dk_steps;  // ceil(d / 16)
  int dv_pad;    // V^T rows per head fed to the P.V MMA (multiple of 16, >= d)
  int vt_rows;   // V^T rows per head in memory (d, or dv_pad when the ones row is present)
  int sum_col;   // column of O that accumulates the softmax denominator (V^T ones row), or -1
  int p_f16;     // P is stored as f16 (exponentials taken two at a time in f16x2) instead of bf16
  int stages;    // K/V ring depth of the two-tile kernel (2 or 3)
  int tmem_cols; // one-tile kernel: TMEM columns to allocate (256 lets two CTAs share an SM)
  int o_col;     // one-tile kernel: first TMEM column of O
  int exp_poly;  // two-tile kernel: every fourth exponential on the FMA pipe (exp2_poly)
  int prescaled; // scale_log2 == 1: scores arrive in log2 units
  int tiles_per_cta;  // one-tile kernel, single key block, no mask: query tiles one CTA walks (K / V^T, TMEM and
                      // barriers set up once); 1 = one tile per CTA
  int qk_fold;   // two-tile kernel: the row offset of the softmax rides in Q.K^T (column d of K is all ones, column d
                 // of the Q tile in shared memory gets -round(row maximum of key block 0)), see the softmax loop
  int stagger;   // two-tile kernel, separate P: clocks tile 1 starts after tile 0 (0 = fixed issue order, lock step)
  int two_issuers;  // two-tile kernel: one MMA issuer warp per query tile (default) instead of one for both
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // pure: let the compiler schedule it freely
  return y;
}
// 2^x without the MUFU: x = n + f, n = round(x), f in [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial (max
// relative error 7.5e-5 - P is rounded to bf16, 2e-3, right after), 2^n by adding n to the exponent field.
// n comes out of the magic-number addition (1.5 * 2^23 + x has round(x) in its low mantissa bits, and
// (bits << 23) is exactly n << 23 because the constant's lowest set bit is bit 22). x >= -126 after the clamp
// (masked scores are -inf), x <= 8 by the lazy-rescale bound.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.0551716685f, f, 0.2426111400f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// two exponentials per MUFU operation: (lo, hi) fp32 -> f16x2 -> 2^x in f16x2
__device__ __forceinline__ uint32_t ex2_f16x2(float lo, float hi) {
  uint32_t h, y;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(hi), "f"(lo));
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(h));
  return y;
}

// Up to two CTAs per SM: with a single key block (cross-attention over the 77 CLIP tokens) S needs no
// double buffer, the CTA allocates 256 TMEM columns and a second CTA hides the load -> MMA -> softmax ->
// MMA -> store latency chain of the first.
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();
  const int q0 = blockIdx.x * ATT_BQ;
  const int h = blockIdx.y;
  const int n = blockIdx.z;

  const int q_bytes = p.dchunks * ATT_CHUNK_BYTES;
  const int k_bytes = p.dchunks * ATT_CHUNK_BYTES;
  const int v_chunk_bytes = p.dv_pad * 128;
  const int stage_bytes = k_bytes + 2 * v_chunk_bytes;
  uint8_t* q_smem = smem;
  uint8_t* kv_smem = smem + q_bytes;
  const int kv_stages = (p.tmem_cols == 256) ? 1 : 2;      // a single key block needs one stage
  const int T = p.tiles_per_cta;
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_smem + kv_stages * stage_bytes);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;    // [2]
  uint64_t* p_full = bars + 7;    // [2]
  uint64_t* o_done = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  uint64_t* q_empty = bars + 11;  // multi-tile form: Q consumed by its Q.K^T, the next tile's Q may be loaded
  uint64_t* o_free = bars + 12;   // O (and S/P) read out: the next tile's Q.K^T may start

  int nkv = (p.Skv + ATT_BKV - 1) / ATT_BKV;
  if (p.causal) {
    const int lim = (min(q0 + ATT_BQ, p.S) + ATT_BKV - 1) / ATT_BKV;
    nkv = min(nkv, lim);
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_q);
    tma_prefetch_desc(&p.map_k);
    tma_prefetch_desc(&p.map_vt);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
    }
    mbar_init(o_done, 1);
    mbar_init(q_empty, 1);
    mbar_init(o_free, 128);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();          // nothing above touched global data (PDL: the predecessor grid may still be running)
  const uint32_t tmem_o = tmem_base + (uint32_t)p.o_col;

  if (T > 1) {
    // ===================== multi-tile form (cross-attention over the 77 CLIP tokens): one key block, no mask.
    // The CTA walks T query tiles of its (sample, head): K / V^T are loaded once, TMEM and the barriers are set
    // up once, the next Q tile is loaded as soon as its predecessor's Q.K^T has completed (it has that tile's
    // softmax, P.V and store to arrive: S / P share TMEM columns, so the next Q.K^T waits for them anyway). Per tile the chain is
    // Q.K^T -> softmax -> P.V -> store (~3 000 clk) instead of a whole CTA life (~9 600 clk); two CTAs per SM
    // interleave their chains as before.
    const int tile0 = blockIdx.x * T;
    const int ntile = min(T, (p.S + ATT_BQ - 1) / ATT_BQ - tile0);
    if (warp == 0) {
      if (elect_one()) {
        uint8_t* kd = kv_smem;
        uint8_t* vd = kd + k_bytes;
        mbar_arrive_expect_tx(&kv_full[0], (uint32_t)stage_bytes);
        for (int c = 0; c < p.dchunks; ++c)
          tma_load_4d(&p.map_k, &kv_full[0], kd + c * ATT_CHUNK_BYTES, c * 64, h, 0, n);
        for (int c = 0; c < 2; ++c)
          tma_load_3d(&p.map_vt, &kv_full[0], vd + c * v_chunk_bytes, c * 64, n, h * p.d);
      }
      __syncwarp();
      for (int i = 0; i < ntile; ++i) {
        if (i >= 1) mbar_wait(q_empty, (uint32_t)((i - 1) & 1), 41);
        if (elect_one()) {
          mbar_arrive_expect_tx(q_full, (uint32_t)q_bytes);
          for (int c = 0; c < p.dchunks; ++c)
            tma_load_4d(&p.map_q, q_full, q_smem + c * ATT_CHUNK_BYTES, c * 64, h, (tile0 + i) * ATT_BQ, n);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      const uint32_t idesc_qk = make_idesc_bf16(ATT_BQ, ATT_BKV);
      const uint32_t idesc_pv = make_idesc_bf16(ATT_BQ, (uint32_t)p.dv_pad);
      const uint32_t k_addr = smem_u32(kv_smem);
      const uint32_t v_addr = k_addr + (uint32_t)k_bytes;
      mbar_wait(&kv_full[0], 0, 42);
      for (int i = 0; i < ntile; ++i) {
        mbar_wait(q_full, (uint32_t)(i & 1), 43);
        if (i > 0) mbar_wait(o_free, (uint32_t)((i - 1) & 1), 44);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t q_addr = smem_u32(q_smem);
          for (int ks = 0; ks < p.dk_steps; ++ks) {
            const uint32_t off = (uint32_t)((ks >> 2) * ATT_CHUNK_BYTES + (ks & 3) * 32);
            mma_ss(tmem_base, make_kmajor_sw128_desc(q_addr + off), make_kmajor_sw128_desc(k_addr + off), idesc_qk,
                   ks > 0 ? 1u : 0u);
          }
          tc_commit(&s_full[0]);
          tc_commit(q_empty);
        }
        __syncwarp();
        mbar_wait(&p_full[0], (uint32_t)(i & 1), 45);
        tc_fence_after();
        if (elect_one()) {
          for (int ks = 0; ks < ATT_BKV / 16; ++ks) {
            const uint32_t off = (uint32_t)((ks >> 2) * v_chunk_bytes + (ks & 3) * 32);
            mma_ts(tmem_o, tmem_base + (uint32_t)(ks * 8), make_kmajor_sw128_desc(v_addr + off), idesc_pv,
                   ks > 0 ? 1u : 0u);
          }
          tc_commit(o_done);
        }
        __syncwarp();
      }
    } else {
      const int quad = warp & 3;
      const int row = quad * 32 + lane;
      const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
      const float sl2 = p.scale_log2;
      const uint32_t s_addr = tmem_base + lane_addr;
      for (int i = 0; i < ntile; ++i) {
        const int qrow = (tile0 + i) * ATT_BQ + row;
        mbar_wait(&s_full[0], (uint32_t)(i & 1), 46);
        tc_fence_after();
        uint32_t sv[128];
        tmem_ld32(s_addr + 0, sv + 0);
        tmem_ld32(s_addr + 32, sv + 32);
        tmem_ld32(s_addr + 64, sv + 64);
        tmem_ld32(s_addr + 96, sv + 96);
        tmem_ld_wait();
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int k = 0; k < 128; k += 2) {
          if (k >= p.Skv) sv[k] = 0xff800000u;
          if (k + 1 >= p.Skv) sv[k + 1] = 0xff800000u;
          m0 = fmaxf(m0, __uint_as_float(sv[k]));
          m1 = fmaxf(m1, __uint_as_float(sv[k + 1]));
        }
        const float mx = fmaxf(m0, m1);
        const float mb = (mx == -INFINITY) ? 0.f : mx * sl2;
        float l0 = 0.f, l1 = 0.f;
        uint32_t pk[64];
#pragma unroll
        for (int k = 0; k < 128; k += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(sv[k]), sl2, -mb));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sv[k + 1]), sl2, -mb));
          l0 += p0; l1 += p1;
          pk[k >> 1] = pack_bf16x2(p0, p1);
        }
        tmem_st32(s_addr + 0, pk + 0);
        tmem_st32(s_addr + 32, pk + 32);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[0]);
        const float lsum = l0 + l1;
        const float inv = (lsum > 0.f) ? 1.0f / lsum : 0.f;
        mbar_wait(o_done, (uint32_t)(i & 1), 47);
        tc_fence_after();
        __nv_bfloat16* orow = p.out + ((long long)n * p.S + qrow) * p.ldo + (long long)h * p.d;
        for (int c = 0; c < p.dv_pad; c += 16) {
          uint32_t ov[16];
          tmem_ld16(tmem_o + lane_addr + (uint32_t)c, ov);
          tmem_ld_wait();
          if (qrow < p.S) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int col = c + g * 8;
              if (col + 8 <= p.d) {
                uint4 u;
                u.x = pack_bf16x2(__uint_as_float(ov[g * 8 + 0]) * inv, __uint_as_float(ov[g * 8 + 1]) * inv);
                u.y = pack_bf16x2(__uint_as_float(ov[g * 8 + 2]) * inv, __uint_as_float(ov[g * 8 + 3]) * inv);
                u.z = pack_bf16x2(__uint_as_float(ov[g * 8 + 4]) * inv, __uint_as_float(ov[g * 8 + 5]) * inv);
                u.w = pack_bf16x2(__uint_as_float(ov[g * 8 + 6]) * inv, __uint_as_float(ov[g * 8 + 7]) * inv);
                *reinterpret_cast<uint4*>(orow + col) = u;
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(o_free);
      }
    }
  } else if (warp == 0) {
    // ===================== TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, (uint32_t)q_bytes);
      for (int c = 0; c < p.dchunks; ++c)
        tma_load_4d(&p.map_q, q_full, q_smem + c * ATT_CHUNK_BYTES, c * 64, h, q0, n);
    }
    __syncwarp();
    for (int j = 0; j < nkv; ++j) {
      const int s = j & 1;
      if (j >= 2) mbar_wait(&kv_empty[s], ((j >> 1) - 1) & 1, 11);
      if (elect_one()) {
        uint8_t* kd = kv_smem + s * stage_bytes;
        uint8_t* vd = kd + k_bytes;
        mbar_arrive_expect_tx(&kv_full[s], (uint32_t)stage_bytes);
        for (int c = 0; c < p.dchunks; ++c)
          tma_load_4d(&p.map_k, &kv_full[s], kd + c * ATT_CHUNK_BYTES, c * 64, h, j * ATT_BKV, n);
        for (int c = 0; c < 2; ++c)
          tma_load_3d(&p.map_vt, &kv_full[s], vd + c * v_chunk_bytes, j * ATT_BKV + c * 64, n, h * p.d);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer
    const uint32_t idesc_qk = make_idesc_bf16(ATT_BQ, ATT_BKV);
    const uint32_t idesc_pv = make_idesc_bf16(ATT_BQ, (uint32_t)p.dv_pad);
    const uint32_t q_addr = smem_u32(q_smem);
    mbar_wait(q_full, 0, 12);
    auto issue_qk = [&](int j) {
      const int s = j & 1;
      mbar_wait(&kv_full[s], (j >> 1) & 1, 13);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t k_addr = smem_u32(kv_smem + s * stage_bytes);
        for (int ks = 0; ks < p.dk_steps; ++ks) {
          const uint32_t off = (uint32_t)((ks >> 2) * ATT_CHUNK_BYTES + (ks & 3) * 32);
          mma_ss(tmem_base + (uint32_t)(s * 128), make_kmajor_sw128_desc(q_addr + off),
                 make_kmajor_sw128_desc(k_addr + off), idesc_qk, ks > 0 ? 1u : 0u);
        }
        tc_commit(&s_full[s]);
      }
      __syncwarp();
    };
    if (nkv > 0) issue_qk(0);
    for (int j = 0; j < nkv; ++j) {
      const int s = j & 1;
      if (j + 1 < nkv) issue_qk(j + 1);
      mbar_wait(&p_full[s], (j >> 1) & 1, 14);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t v_addr = smem_u32(kv_smem + s * stage_bytes + k_bytes);
        for (int ks = 0; ks < ATT_BKV / 16; ++ks) {
          const uint32_t off = (uint32_t)((ks >> 2) * v_chunk_bytes + (ks & 3) * 32);
          mma_ts(tmem_o, tmem_base + (uint32_t)(s * 128 + ks * 8), make_kmajor_sw128_desc(v_addr + off),
                 idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
        }
        tc_commit(&kv_empty[s]);
        tc_commit(o_done);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax / correction / output: one query row per thread
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int qrow = q0 + row;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    float m_ref = -INFINITY;  // maximum the exponentials are currently taken against
    float l_run = 0.f;
    const float sl2 = p.scale_log2;
    for (int j = 0; j < nkv; ++j) {
      const int s = j & 1;
      mbar_wait(&s_full[s], (j >> 1) & 1, 15);
      tc_fence_after();
      uint32_t sv[128];
      const uint32_t s_addr = tmem_base + lane_addr + (uint32_t)(s * 128);
      tmem_ld32(s_addr + 0, sv + 0);
      tmem_ld32(s_addr + 32, sv + 32);
      tmem_ld32(s_addr + 64, sv + 64);
      tmem_ld32(s_addr + 96, sv + 96);
      tmem_ld_wait();
      const int key0 = j * ATT_BKV;
      int visible = p.Skv - key0;                       // keys [0, visible) of this block are real
      if (p.causal) visible = min(visible, qrow - key0 + 1);
      float mloc = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; ++i) {
        float v = __uint_as_float(sv[i]);
        v = (i < visible) ? v : -INFINITY;
        sv[i] = __float_as_uint(v);
        mloc = fmaxf(mloc, v);
      }
      const float m_new = fmaxf(m_ref, mloc);
      // lazy rescale: keep the old reference maximum while the new one is within 2^8 of it
      const bool need = (m_new > m_ref) && ((m_new - m_ref) * sl2 > 8.0f);
      const bool any_need = __any_sync(0xffffffffu, need);
      if (any_need && j > 0) {
        mbar_wait(o_done, (j - 1) & 1, 16);
        tc_fence_after();
        const float factor = need ? ex2_approx((m_ref - m_new) * sl2) : 1.0f;
        for (int c = 0; c < p.dv_pad; c += 16) {
          uint32_t ov[16];
          tmem_ld16(tmem_o + lane_addr + (uint32_t)c, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * factor);
          tmem_st16(tmem_o + lane_addr + (uint32_t)c, ov);
        }
        tmem_st_wait();
        l_run *= factor;
      }
      if (need) m_ref = m_new;
      const float mb = (m_ref == -INFINITY) ? 0.f : m_ref * sl2;
      float lsum = 0.f;
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 128; i += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(sv[i]), sl2, -mb));
        const float p1 = ex2_approx(fmaf(__uint_as_float(sv[i + 1]), sl2, -mb));
        lsum += p0 + p1;
        pk[i >> 1] = pack_bf16x2(p0, p1);
      }
      l_run += lsum;
      tmem_st32(s_addr + 0, pk + 0);
      tmem_st32(s_addr + 32, pk + 32);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[s]);
    }
    // ---- epilogue: O / l -> bf16
    if (nkv > 0) {
      mbar_wait(o_done, (nkv - 1) & 1, 17);
      tc_fence_after();
    }
    const float inv = (l_run > 0.f) ? 1.0f / l_run : 0.f;
    __nv_bfloat16* orow = p.out + ((long long)n * p.S + qrow) * p.ldo + (long long)h * p.d;
    for (int c = 0; c < p.dv_pad; c += 16) {
      uint32_t ov[16];
      if (nkv > 0) {
        tmem_ld16(tmem_o + lane_addr + (uint32_t)c, ov);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) ov[i] = 0u;
      }
      if (qrow < p.S) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int col = c + g * 8;
          if (col + 8 <= p.d) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(ov[g * 8 + 0]) * inv, __uint_as_float(ov[g * 8 + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(ov[g * 8 + 2]) * inv, __uint_as_float(ov[g * 8 + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(ov[g * 8 + 4]) * inv, __uint_as_float(ov[g * 8 + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(ov[g * 8 + 6]) * inv, __uint_as_float(ov[g * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + col) = u;
          }
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}


// ------------------------------------------------------------------------------------------------
// Two-tile variant (head dim <= 128, no causal mask): one CTA owns 256 queries as two 128-query tiles
// that share every K / V^T block. While the softmax warpgroup of one tile works on S_t, the tensor
// pipe runs the other tile's P.V and the next Q.K^T, so exponentials (MUFU) and MMAs overlap:
//   warp 0     TMA: both Q tiles once, then K block + V^T block per 128 keys into a 3-stage ring
//   warp 1     tcgen05.mma issuer (one elected lane), order per key block j:
//                P0.V(j)  Q0.K(j+1)  P1.V(j)  Q1.K(j+1)
//   warps 2-5  softmax of tile 0, warps 6-9 softmax of tile 1 (one query row per thread)
// TMEM: S0/P0 at columns [0,128), S1/P1 at [128,256), O0 at [256,384), O1 at [384,512).
constexpr int ATT2_THREADS = 352;
// Warp roles of the two-tile kernel: softmax warps 0-3 (tile 0) and 4-7 (tile 1) - TMEM lane quadrant = warp % 4 -,
// then the TMA producer and the MMA issuer (highest warp id on its sub-partition; measured neutral against the
// issuer as warp 1).
constexpr int ATT2_TMA_WARP = 8;
constexpr int ATT2_MMA_WARP = 9;
// Second MMA issuer. A tcgen05.mma costs its ISSUING WARP ~60 ns (117 clk at 1.96 GHz) whatever its N - measured with
// tools/micro/mma_issuers.cu: one warp 117 clk/MMA for N = 48 ... 128; two warps issuing independent streams 58.6
// aggregate; four warps 29 (TS operands: the pipe floor N/2 = 24 is finally in sight). With d = 40 the 22 small MMAs
// of a key block (2 x [3 Q.K^T + 8 P.V]) held ONE issuer for 2 600 of the ~2 950 clocks a key block took; with one
// issuer warp per query tile the two tiles' chains run side by side.
constexpr int ATT2_MMA_WARP2 = 10;
constexpr int ATT2_MAX_STAGES = 3;   // K/V ring depth (4 and 6 measured no faster)

__global__ void __launch_bounds__(ATT2_THREADS, 1)
attn2_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();
  const int q0 = blockIdx.x * (2 * ATT_BQ);
  const int h = blockIdx.y;
  const int n = blockIdx.z;

  const int q_bytes = p.dchunks * ATT_CHUNK_BYTES;          // one 128-query tile
  const int k_bytes = p.dchunks * ATT_CHUNK_BYTES;
  const int v_chunk_bytes = p.dv_pad * 128;
  const int stage_bytes = k_bytes + 2 * v_chunk_bytes;
  uint8_t* q_smem = smem;
  uint8_t* kv_smem = smem + 2 * q_bytes;
  const int nstages = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_smem + nstages * stage_bytes);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;                      // [ATT2_MAX_STAGES]
  uint64_t* kv_empty = kv_full + ATT2_MAX_STAGES;    // [ATT2_MAX_STAGES]
  uint64_t* s_full = kv_empty + ATT2_MAX_STAGES;     // [2]
  uint64_t* p_full = s_full + 2;                 // [2]
  uint64_t* o_done = p_full + 2;                 // [2]
  uint64_t* s_free = o_done + 2;                 // [2] softmax has S_t in registers (separate-P layout)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);
  uint64_t* q_ready = s_free + 3;                // [2] qk_fold: the rows' offsets are in the Q tile
  // TMEM columns. Narrow heads (dv_pad <= 64) keep P apart from S, so Q.K^T of the next key block can
  // overwrite S_t while the softmax of the current one is still computing and P_t waits for its P.V:
  //   separate:  S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O0 [384,448) O1 [448,512)
  //   aliased:   S0/P0 [0,128) S1/P1 [128,256) O0 [256,384) O1 [384,512)
  const bool p_sep = p.dv_pad <= 64;
  const uint32_t p_col0 = p_sep ? 256u : 0u, p_colstep = p_sep ? 64u : 128u;
  const uint32_t o_col0 = p_sep ? 384u : 256u, o_colstep = p_sep ? 64u : 128u;

  const int nkv = (p.Skv + ATT_BKV - 1) / ATT_BKV;
  const int trc = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? g_gemm_trace_on : 0;
#define ATT_STAMP(j, slot) do { if (trc && (j) < GEMM_TRACE_TILES) g_gemm_trace[(j) * 8 + (slot)] = clock64(); } while (0)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_q);
    tma_prefetch_desc(&p.map_k);
    tma_prefetch_desc(&p.map_vt);
    mbar_init(q_full, 1);
    for (int i = 0; i < nstages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], p.two_issuers ? 2 : 1);      // a K/V stage is free once BOTH tiles' P.V over it are done
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_done[i], 1);
      mbar_init(&s_free[i], 128);
      mbar_init(&q_ready[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == ATT2_MMA_WARP) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();          // nothing above touched global data (PDL: the predecessor grid may still be running)

  if (warp == ATT2_TMA_WARP) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues)
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, (uint32_t)(2 * q_bytes));
      for (int t = 0; t < 2; ++t)
        for (int c = 0; c < p.dchunks; ++c)
          tma_load_4d(&p.map_q, q_full, q_smem + t * q_bytes + c * ATT_CHUNK_BYTES, c * 64, h,
                      q0 + t * ATT_BQ, n);
    }
    __syncwarp();
    int st = 0;
    uint32_t ph = 1;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&kv_empty[st], ph, 21);
      uint8_t* kd = kv_smem + st * stage_bytes;
      uint8_t* vd = kd + k_bytes;
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[st], (uint32_t)stage_bytes);
        for (int c = 0; c < p.dchunks; ++c)
          tma_load_4d(&p.map_k, &kv_full[st], kd + c * ATT_CHUNK_BYTES, c * 64, h, j * ATT_BKV, n);
        tma_load_3d(&p.map_vt, &kv_full[st], vd, j * ATT_BKV, n, h * p.vt_rows);
        tma_load_3d(&p.map_vt, &kv_full[st], vd + v_chunk_bytes, j * ATT_BKV + 64, n, h * p.vt_rows);
      }
      __syncwarp();
      if (++st == nstages) { st = 0; ph ^= 1u; }
    }
  } else if (warp == ATT2_MMA_WARP || (warp == ATT2_MMA_WARP2 && p.two_issuers)) {
    // ===================== MMA issuer(s)
    const uint32_t idesc_qk = make_idesc_bf16(ATT_BQ, ATT_BKV);
    // p_f16 variant: P (A operand, from TMEM) and V^T (B operand) are both IEEE half (formats 0)
    const uint32_t idesc_pv = make_idesc_bf16(ATT_BQ, (uint32_t)p.dv_pad) & (p.p_f16 ? ~((7u << 7) | (7u << 10)) : ~0u);
    const uint32_t q_addr = smem_u32(q_smem);
    const uint32_t kv_addr = smem_u32(kv_smem);
    mbar_wait(q_full, 0, 22);
    // S_t = Q_t . K(stage)^T
    auto issue_qk = [&](int t, int st) {
      const uint32_t qa = q_addr + (uint32_t)(t * q_bytes);
      const uint32_t ka = kv_addr + (uint32_t)(st * stage_bytes);
      const uint32_t d_tmem = tmem_base + (uint32_t)(t * 128);
      if (elect_one()) {
        for (int ks = 0; ks < p.dk_steps; ++ks) {
          const uint32_t off = (uint32_t)((ks >> 2) * ATT_CHUNK_BYTES + (ks & 3) * 32);
          mma_ss(d_tmem, make_kmajor_sw128_desc(qa + off), make_kmajor_sw128_desc(ka + off), idesc_qk,
                 ks > 0 ? 1u : 0u);
        }
        tc_commit(&s_full[t]);
      }
      __syncwarp();
    };
    // O_t += P_t . V(stage)
    auto issue_pv = [&](int t, int st, bool first, bool release) {
      const uint32_t va = kv_addr + (uint32_t)(st * stage_bytes + k_bytes);
      const uint32_t d_tmem = tmem_base + o_col0 + (uint32_t)t * o_colstep;
      const uint32_t p_tmem = tmem_base + p_col0 + (uint32_t)t * p_colstep;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < ATT_BKV / 16; ++ks) {
          const uint32_t off = (uint32_t)((ks >> 2) * v_chunk_bytes + (ks & 3) * 32);
          mma_ts(d_tmem, p_tmem + (uint32_t)(ks * 8), make_kmajor_sw128_desc(va + off), idesc_pv,
                 (!first || ks > 0) ? 1u : 0u);
        }
        tc_commit(&o_done[t]);
        if (release) tc_commit(&kv_empty[st]);
      }
      __syncwarp();
    };
    mbar_wait(&kv_full[0], 0, 23);
    tc_fence_after();
    if (p.two_issuers) {
      // one issuer per query tile: this warp owns tile t's Q.K^T / P.V chain; the tiles only meet at the K/V ring
      // (both wait for kv_full, both commit their P.V to kv_empty)
      const int t = warp - ATT2_MMA_WARP;
      if (t == 1 && p.stagger < 0) {
        // tile 1 starts about half a period after tile 0 so that the exponential phases of the two softmax
        // warpgroups interleave on the sub-partitions (see the single-issuer form below)
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)(-p.stagger)) { }
      }
      issue_qk(t, 0);
      int st = 0;
      uint32_t ph_kv = 0;
      for (int j = 0; j < nkv; ++j) {
        int st_next = st + 1;
        uint32_t ph_next = ph_kv;
        if (st_next == nstages) { st_next = 0; ph_next ^= 1u; }
        const bool more = (j + 1 < nkv);
        const uint32_t par = (uint32_t)(j & 1);
        if (p_sep) {
          // P apart from S: the next key block's scores only need S_t to have been read
          if (more) {
            mbar_wait(&kv_full[st_next], ph_next, 25);
            mbar_wait(&s_free[t], par, 30);
            if (p.qk_fold && j == 0) mbar_wait(&q_ready[t], 0, 34);       // the rows' offsets are in the Q tile
            tc_fence_after();
            issue_qk(t, st_next);
          }
          if (t == 0 && lane == 0) ATT_STAMP(j, 7);
          mbar_wait(&p_full[t], par, 24);
          tc_fence_after();
          issue_pv(t, st, j == 0, true);
          if (t == 0 && lane == 0) ATT_STAMP(j, 6);
        } else {
          mbar_wait(&p_full[t], par, 24);
          tc_fence_after();
          issue_pv(t, st, j == 0, true);
          if (more) {
            mbar_wait(&kv_full[st_next], ph_next, 25);
            tc_fence_after();
            issue_qk(t, st_next);
          }
        }
        st = st_next;
        ph_kv = ph_next;
      }
    } else if (p_sep && p.stagger > 0) {
      // EXPERIMENT, off by default (SDB_ATTN_STAGGER=<clocks> enables it): event-driven issue with tile 1 started
      // <clocks> after tile 0. In lock step both tiles' softmax warps sit in the exponential loop at the same time
      // (the sub-partitions are saturated) and then in the TMEM load / row maximum / P store phases at the same
      // time (they idle); staggered, the exponential loop of a tile alone on its sub-partitions does run faster
      // (1180 instead of 1900 clk per key block) - but tcgen05.mma issue blocks the issuing thread while the
      // tensor pipe is busy, each tile then waits ~1500 clk for its next S, and the kernel is 10 % SLOWER
      // (S = 4096, d = 40: 952 us against 844 us; deeper K/V ring and issuer warp id made no difference).
      auto ready = [&](uint64_t* bar, uint32_t parity) {
        return __shfl_sync(0xffffffffu, mbar_test_wait(bar, parity) ? 1 : 0, 0) != 0;
      };
      int jq[2] = {1, 0};              // next key block whose scores are to be issued, per tile
      int jp[2] = {0, 0};              // next key block whose P.V is to be issued
      issue_qk(0, 0);
      const long long t_start = clock64();
      unsigned spins = 0;
      while (jp[0] < nkv || jp[1] < nkv) {
        bool progress = false;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (jq[t] < nkv) {
            const int stg = jq[t] % nstages;
            const uint32_t ph = (uint32_t)((jq[t] / nstages) & 1);
            bool ok = (jq[t] == 0) ? (clock64() - t_start >= (long long)p.stagger)
                                   : ready(&s_free[t], (uint32_t)((jq[t] - 1) & 1));
            if (ok && ready(&kv_full[stg], ph)) {
              tc_fence_after();
              issue_qk(t, stg);
              if (t == 0 && lane == 0) ATT_STAMP(jq[0] - 1, 7);      // scores of the NEXT block issued (row of block j)
              ++jq[t];
              progress = true;
            }
          }
          if (jp[t] < nkv && jp[t] < jq[t] && ready(&p_full[t], (uint32_t)(jp[t] & 1))) {
            tc_fence_after();
            // the K/V stage is released by whichever tile's product over it is issued second
            issue_pv(t, jp[t] % nstages, jp[t] == 0, jp[1 - t] > jp[t]);
            if (t == 0 && lane == 0) ATT_STAMP(jp[0], 6);
            ++jp[t];
            progress = true;
          }
        }
        if (progress) spins = 0;
        else if (++spins > (1u << 24)) { if (lane == 0) atomicCAS(&g_sdb_fault, 0u, (33u << 8) | 1u); break; }
      }
    } else {
    issue_qk(0, 0);
    if (p.stagger < 0) {
      // Fixed issue order, tile 1 started -stagger clocks after tile 0. In lock step both tiles' softmax warps sit
      // in the exponential loop at the same time (the sub-partitions saturated) and then in the TMEM load / row
      // maximum / P store phases at the same time (idle); nothing in the protocol re-synchronises the tiles, so an
      // initial offset of about half a period persists and the phases of one tile fill the gaps of the other
      // (period 3 060 -> 2 800 clk).
      const long long t0 = clock64();
      while (clock64() - t0 < (long long)(-p.stagger)) { }
    }
    issue_qk(1, 0);
    int st = 0;
    uint32_t ph_kv = 0;                 // phase of kv_full[next stage]
    for (int j = 0; j < nkv; ++j) {
      int st_next = st + 1;
      uint32_t ph_next = ph_kv;
      if (st_next == nstages) { st_next = 0; ph_next ^= 1u; }
      const bool more = (j + 1 < nkv);
      const uint32_t par = (uint32_t)(j & 1);
      if (p_sep) {
        // next key block's scores first: they only need S_t to have been read, not P_t.V to be done
        if (more) {
          mbar_wait(&kv_full[st_next], ph_next, 25);
          mbar_wait(&s_free[0], par, 30);
          if (p.qk_fold && j == 0) mbar_wait(&q_ready[0], 0, 34);     // tile 0's row offsets are in its Q tile
          tc_fence_after();
          issue_qk(0, st_next);
          mbar_wait(&s_free[1], par, 31);
          if (p.qk_fold && j == 0) mbar_wait(&q_ready[1], 0, 35);
          tc_fence_after();
          issue_qk(1, st_next);
        }
        if (lane == 0) ATT_STAMP(j, 7);
        mbar_wait(&p_full[0], par, 24);
        tc_fence_after();
        issue_pv(0, st, j == 0, false);
        if (lane == 0) ATT_STAMP(j, 6);
        mbar_wait(&p_full[1], par, 26);
        tc_fence_after();
        issue_pv(1, st, j == 0, true);
      } else {
        mbar_wait(&p_full[0], par, 24);
        tc_fence_after();
        issue_pv(0, st, j == 0, false);
        if (more) {
          mbar_wait(&kv_full[st_next], ph_next, 25);
          tc_fence_after();
          issue_qk(0, st_next);
        }
        mbar_wait(&p_full[1], par, 26);
        tc_fence_after();
        issue_pv(1, st, j == 0, true);
        if (more) issue_qk(1, st_next);
      }
      st = st_next;
      ph_kv = ph_next;
    }
    }
  } else if (warp < 8) {
    // ===================== softmax / correction / output: warps 0-3 tile 0, warps 4-7 tile 1
    // (warp 10 falls through idle when the kernel runs with a single issuer)
    const int t = warp >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int qrow = q0 + t * ATT_BQ + row;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + (uint32_t)(t * 128);
    const uint32_t p_addr = tmem_base + lane_addr + p_col0 + (uint32_t)t * p_colstep;
    const uint32_t o_addr = tmem_base + lane_addr + o_col0 + (uint32_t)t * o_colstep;
    float m_ref = -INFINITY;  // maximum the exponentials are currently taken against
    float l_run = 0.f;
    const float sl2 = p.scale_log2;
    const bool tr0 = (t == 0 && quad == 2 && lane == 0);
    const bool fold = p.qk_fold != 0;
    for (int j = 0; j < nkv; ++j) {
      if (tr0) ATT_STAMP(j, 0);
      mbar_wait(&s_full[t], (uint32_t)(j & 1), 27);
      if (tr0) ATT_STAMP(j, 1);
      tc_fence_after();
      uint32_t sv[128];
      tmem_ld32(s_addr + 0, sv + 0);
      tmem_ld32(s_addr + 32, sv + 32);
      tmem_ld32(s_addr + 64, sv + 64);
      tmem_ld32(s_addr + 96, sv + 96);
      tmem_ld_wait();
      if (p_sep) {                       // S_t is in registers: the next Q.K^T may overwrite it
        tc_fence_before();
        mbar_arrive(&s_free[t]);
      }
      if (tr0) ATT_STAMP(j, 2);
      const int visible = p.Skv - j * ATT_BKV;          // keys [0, visible) of this block are real
      if (fold && j > 0 && visible >= 128 && __all_sync(0xffffffffu, m_ref == 0.f)) {
        // qk_fold, no row-maximum pass: the scores arrive relative to (row maximum of block 0) + 7, i.e. every
        // probability carries 2^-7 (so does the denominator column of O: it cancels exactly). "This block raised
        // the row maximum by 2^8 or more" is then "some probability >= 2.0" = bit 14 of a bf16 pattern: OR the packed
        // words (32 LOP3 instead of 64 FMNMX3) and test that bit in both halves.
        uint32_t pk[64];
        uint32_t orv = 0;
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          const float p0 = ex2_approx(__uint_as_float(sv[i]));
          const float p1 = ex2_approx(__uint_as_float(sv[i + 1]));
          const float p2 = ex2_approx(__uint_as_float(sv[i + 2]));
          const float p3 = exp2_poly(__uint_as_float(sv[i + 3]));
          pk[i >> 1] = pack_bf16x2(p0, p1);
          pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
          orv |= pk[i >> 1] | pk[(i >> 1) + 1];
        }
        if (tr0) ATT_STAMP(j, 4);
        mbar_wait(&o_done[t], (uint32_t)((j - 1) & 1), 32);     // P_t of the previous block consumed by its P.V
        tc_fence_after();
        const bool grow = (orv & 0x40004000u) != 0u;
        if (__any_sync(0xffffffffu, grow)) {
          // rare: rescale O (denominator column included) and this block's P by an exact power of two and carry
          // the shift as the row's relative reference - those rows take the subtracting path from here on
          float pmax = 0.f;
#pragma unroll
          for (int i = 0; i < 64; ++i) pmax = fmaxf(pmax, fmaxf(bf16lo(pk[i]), bf16hi(pk[i])));
          const float e = grow ? floorf(log2f(pmax)) + 7.0f : 0.f;
          const float factor = grow ? ex2_approx(-e) : 1.0f;
          for (int c = 0; c < p.dv_pad; c += 16) {
            uint32_t ov[16];
            tmem_ld16(o_addr + (uint32_t)c, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * factor);
            tmem_st16(o_addr + (uint32_t)c, ov);
          }
#pragma unroll
          for (int i = 0; i < 64; ++i) pk[i] = pack_bf16x2(bf16lo(pk[i]) * factor, bf16hi(pk[i]) * factor);
          if (grow) m_ref = e;
        }
        tmem_st32(p_addr + 0, pk + 0);
        tmem_st32(p_addr + 32, pk + 32);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[t]);
        if (tr0) ATT_STAMP(j, 5);
        continue;
      }
      float mloc = -INFINITY;
      if (visible < 128) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= visible) sv[i] = 0xff800000u;           // -inf
      }
      {
        // four independent chains: a single running maximum is a 64-deep dependent chain
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 128; i += 8) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(sv[i + 0]), __uint_as_float(sv[i + 1])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])));
          m2 = fmaxf(m2, fmaxf(__uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5])));
          m3 = fmaxf(m3, fmaxf(__uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7])));
        }
        mloc = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      const float m_new = fmaxf(m_ref, mloc);
      // lazy rescale: keep the old reference maximum while the new one is within 2^8 of it
      const bool need = (m_new > m_ref) && ((m_new - m_ref) * sl2 > 8.0f);
      const bool any_need = __any_sync(0xffffffffu, need);
      if (any_need && j > 0) {
        mbar_wait(&o_done[t], (uint32_t)((j - 1) & 1), 28);
        tc_fence_after();
        const float factor = need ? ex2_approx((m_ref - m_new) * sl2) : 1.0f;
        for (int c = 0; c < p.dv_pad; c += 16) {
          uint32_t ov[16];
          tmem_ld16(o_addr + (uint32_t)c, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * factor);
          tmem_st16(o_addr + (uint32_t)c, ov);
        }
        tmem_st_wait();
        l_run *= factor;
      }
      if (need) m_ref = m_new;
      if (fold && j == 0) {
        // qk_fold: from the second key block on, the scores leave the tensor core already relative to this row's
        // reference. Column d of K is all ones; column d of this row of the Q tile (shared memory, 128B-swizzled:
        // 16-byte chunk (d/8) ^ (row & 7)) gets -m, m = the row maximum of block 0 rounded to an integer (exact in
        // bf16 up to 256; scores are in log2 units, so 2^-m is an exact factor and the ones-row denominator sees the
        // same P). The issuer waits for q_ready before the second block's Q.K^T.
        float mi = (m_ref == -INFINITY) ? 0.f : rintf(m_ref) + 7.0f;   // + 7: see the no-max-pass path below
        mi = fminf(fmaxf(mi, -256.f), 256.f);
        m_ref = mi;
        const uint32_t qrow_addr = smem_u32(q_smem) + (uint32_t)(t * q_bytes + ((p.d >> 6) * ATT_CHUNK_BYTES)) +
                                   (uint32_t)(row * 128) + (uint32_t)(((((p.d & 63) >> 3) ^ (row & 7)) << 4) + (p.d & 7) * 2);
        const __nv_bfloat16 hv = __float2bfloat16_rn(-mi);
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(qrow_addr), "h"(*reinterpret_cast<const uint16_t*>(&hv)) : "memory");
        fence_proxy_async();
        mbar_arrive(&q_ready[t]);
      }
      const float mb = (m_ref == -INFINITY) ? 0.f : m_ref * sl2;
      if (tr0) ATT_STAMP(j, 3);
      uint32_t pk[64];
      if (p.p_f16) {
        // denominator comes from the V^T ones row (sum_col >= 0 is required with p_f16)
#pragma unroll
        for (int i = 0; i < 128; i += 2)
          pk[i >> 1] = ex2_f16x2(fmaf(__uint_as_float(sv[i]), sl2, -mb), fmaf(__uint_as_float(sv[i + 1]), sl2, -mb));
      } else if (p.sum_col >= 0 && p.exp_poly && p.prescaled) {
        // scores already in log2 units (q carries log2(e)/sqrt(d) from its projection): the argument of the
        // exponential is a 2-register FADD instead of a 3-register FFMA. Measured (tools/micro/softmax_issue2.cu):
        // the loop is bound by the ~1.7 clk a 3-register fp32 instruction costs per sub-partition as much as by
        // the MUFU - 1025 -> 894 clk per 128-column row and warp.
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          const float p0 = ex2_approx(__uint_as_float(sv[i]) - mb);
          const float p1 = ex2_approx(__uint_as_float(sv[i + 1]) - mb);
          const float p2 = ex2_approx(__uint_as_float(sv[i + 2]) - mb);
          const float p3 = exp2_poly(__uint_as_float(sv[i + 3]) - mb);
          pk[i >> 1] = pack_bf16x2(p0, p1);
          pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
        }
      } else if (p.sum_col >= 0 && p.exp_poly) {
        // no row sums here (the ones row of V^T makes the P.V product accumulate them) and a quarter of the
        // exponentials on the FMA pipe
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(sv[i]), sl2, -mb));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sv[i + 1]), sl2, -mb));
          const float p2 = ex2_approx(fmaf(__uint_as_float(sv[i + 2]), sl2, -mb));
          const float p3 = exp2_poly(fmaf(__uint_as_float(sv[i + 3]), sl2, -mb));
          pk[i >> 1] = pack_bf16x2(p0, p1);
          pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
        }
      } else if (p.sum_col >= 0) {
#pragma unroll
        for (int i = 0; i < 128; i += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(sv[i]), sl2, -mb));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sv[i + 1]), sl2, -mb));
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
      } else if (p.exp_poly) {
        // the exponentials bound this kernel (16 MUFU ops / clk / SM): every fourth one goes to the FMA pipe
        float lsum0 = 0.f, lsum1 = 0.f;
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(sv[i]), sl2, -mb));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sv[i + 1]), sl2, -mb));
          const float p2 = ex2_approx(fmaf(__uint_as_float(sv[i + 2]), sl2, -mb));
          const float p3 = exp2_poly(fmaf(__uint_as_float(sv[i + 3]), sl2, -mb));
          lsum0 += p0 + p2;
          lsum1 += p1 + p3;
          pk[i >> 1] = pack_bf16x2(p0, p1);
          pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
        }
        l_run += lsum0 + lsum1;
      } else {
        float lsum0 = 0.f, lsum1 = 0.f;
#pragma unroll
        for (int i = 0; i < 128; i += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(sv[i]), sl2, -mb));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sv[i + 1]), sl2, -mb));
          lsum0 += p0;
          lsum1 += p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        l_run += lsum0 + lsum1;
      }
      if (tr0) ATT_STAMP(j, 4);
      if (p_sep && j > 0) {              // P_t of the previous block must have been consumed by its P.V
        mbar_wait(&o_done[t], (uint32_t)((j - 1) & 1), 32);
        tc_fence_after();
      }
      tmem_st32(p_addr + 0, pk + 0);
      tmem_st32(p_addr + 32, pk + 32);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
      if (fold && j == 0) m_ref = 0.f;       // from here on the scores arrive relative to the baked reference
      if (tr0) ATT_STAMP(j, 5);
    }
    // ---- epilogue: O / l -> bf16
    mbar_wait(&o_done[t], (uint32_t)((nkv - 1) & 1), 29);
    tc_fence_after();
    if (p.sum_col >= 0) {               // l = sum_k P[k] * 1, accumulated by the MMA in column sum_col of O
      uint32_t lv[16];
      tmem_ld16(o_addr + (uint32_t)(p.sum_col & ~15), lv);
      tmem_ld_wait();
      float l = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i == (p.sum_col & 15)) l = __uint_as_float(lv[i]);
      l_run = l;
    }
    const float inv = (l_run > 0.f) ? 1.0f / l_run : 0.f;
    __nv_bfloat16* orow = p.out + ((long long)n * p.S + qrow) * p.ldo + (long long)h * p.d;
    for (int c = 0; c < p.dv_pad; c += 16) {
      uint32_t ov[16];
      tmem_ld16(o_addr + (uint32_t)c, ov);
      tmem_ld_wait();
      if (qrow < p.S) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int col = c + g * 8;
          if (col + 8 <= p.d) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(ov[g * 8 + 0]) * inv, __uint_as_float(ov[g * 8 + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(ov[g * 8 + 2]) * inv, __uint_as_float(ov[g * 8 + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(ov[g * 8 + 4]) * inv, __uint_as_float(ov[g * 8 + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(ov[g * 8 + 6]) * inv, __uint_as_float(ov[g * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + col) = u;
          }
        }
      }
      __syncwarp();
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == ATT2_MMA_WARP) tmem_dealloc(tmem_base, 512);
}

}  // namespace sdb

extern "C" int sdb_attention(const sdb_attn_args* a, void* stream) {
  using namespace sdb;
  if (!a || !a->q || !a->k || !a->vt || !a->out) { set_error("sdb_attention: null pointer"); return SDB_ERR_ARG; }
  if (a->NB <= 0 || a->heads <= 0 || a->S <= 0 || a->Skv <= 0 || a->Skv_pad < a->Skv) {
    set_error("sdb_attention: bad sizes"); return SDB_ERR_ARG;
  }
  if (a->d % 8 != 0 || a->d < 8 || a->d > 160) {
    set_error("sdb_attention: head dim %d unsupported (multiple of 8, <= 160)", a->d);
    return SDB_ERR_UNSUPPORTED;
  }
  const int vt_ld = a->vt_ld ? a->vt_ld : a->Skv_pad;
  if (vt_ld % 8 != 0 || vt_ld < a->Skv) {
    set_error("sdb_attention: vt_ld (%d) must be a multiple of 8 and >= Skv", vt_ld);
    return SDB_ERR_ARG;
  }
  const long long ldq = a->ldq ? a->ldq : (long long)a->heads * a->d;
  const long long ldk = a->ldk ? a->ldk : (long long)a->heads * a->d;
  const long long ldo = a->ldo ? a->ldo : (long long)a->heads * a->d;
  if (((ldo * 2) % 16) != 0 || (reinterpret_cast<uintptr_t>(a->out) & 15u)) {
    set_error("sdb_attention: output must be 16-byte aligned with ldo %% 8 == 0"); return SDB_ERR_ARG;
  }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  // columns stored per head in q and k (>= d; the extra ones are part of the reduction: zeros, or the ones / offset
  // column of qk_fold)
  const int dqk = a->qk_cols > 0 ? a->qk_cols : a->d;
  if (dqk < a->d || dqk % 8 != 0 || dqk > 160) { set_error("sdb_attention: qk_cols %d invalid", dqk); return SDB_ERR_ARG; }
  {
    uint64_t dims[4] = {(uint64_t)dqk, (uint64_t)a->heads, (uint64_t)a->S, (uint64_t)a->NB};
    uint64_t str[3] = {(uint64_t)dqk * 2, (uint64_t)ldq * 2, (uint64_t)ldq * 2 * a->S};
    uint32_t box[4] = {64, 1, 128, 1};
    if ((rc = make_tmap_bf16(&p.map_q, a->q, 4, dims, str, box, "attention Q"))) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)dqk, (uint64_t)a->heads, (uint64_t)a->Skv_pad, (uint64_t)a->NB};
    uint64_t str[3] = {(uint64_t)dqk * 2, (uint64_t)ldk * 2, (uint64_t)ldk * 2 * a->Skv_pad};
    uint32_t box[4] = {64, 1, 128, 1};
    if ((rc = make_tmap_bf16(&p.map_k, a->k, 4, dims, str, box, "attention K"))) return rc;
  }
  // V^T rows per head: d, or (sum_row) round16(d + 1) with row d all ones - the P.V MMA then also
  // accumulates the softmax denominator in column d of O
  const int dv_pad = ((a->d + (a->sum_row ? 1 : 0) + 15) / 16) * 16;
  const int vt_rows = a->sum_row ? dv_pad : a->d;
  if (a->p_f16 && !a->sum_row) { set_error("sdb_attention: p_f16 needs sum_row"); return SDB_ERR_ARG; }
  if (a->sum_row && (a->causal || a->d > 112)) {
    set_error("sdb_attention: sum_row is implemented by the two-tile kernel only (no causal mask, d <= 112)");
    return SDB_ERR_UNSUPPORTED;
  }
  {
    uint64_t dims[3] = {(uint64_t)vt_ld, (uint64_t)a->NB, (uint64_t)a->heads * vt_rows};
    uint64_t str[2] = {(uint64_t)vt_ld * 2, (uint64_t)vt_ld * 2 * a->NB};
    uint32_t box[3] = {64, 1, (uint32_t)dv_pad};
    if ((rc = make_tmap_bf16(&p.map_vt, a->vt, 3, dims, str, box, "attention V^T"))) return rc;
  }
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.ldo = ldo;
  p.S = a->S; p.Skv = a->Skv; p.d = a->d; p.heads = a->heads; p.NB = a->NB;
  p.causal = a->causal;
  p.scale_log2 = a->q_prescaled ? 1.0f : a->scale * 1.4426950408889634f;
  p.prescaled = a->q_prescaled ? 1 : 0;
  {
    static int stagger = 0;
    static bool stagger_set = false;
    if (!stagger_set) { const char* ev = getenv("SDB_ATTN_STAGGER"); stagger = ev ? atoi(ev) : -1300; stagger_set = true; }
    // default: fixed issue order, tile 1 started 1 300 clocks after tile 0 when there are enough key blocks for
    // the offset to pay (S = 4096, d = 40: 843 -> 783 us; 1 100 / 1 500 / 1 900 clocks: 800 / 791 / 797 us)
    p.stagger = (stagger < 0 && (a->Skv + ATT_BKV - 1) / ATT_BKV < 8) ? 0 : stagger;
    static int one_issuer = -1;
    if (one_issuer < 0) { const char* ev = getenv("SDB_ATTN_ONE_ISSUER"); one_issuer = (ev && ev[0] == '1') ? 1 : 0; }
    // one MMA issuer warp per query tile (the event-driven experiment, stagger > 0, keeps its single issuer)
    p.two_issuers = (!one_issuer && p.stagger <= 0) ? 1 : 0;
  }
  p.dchunks = (dqk + 63) / 64;
  p.dk_steps = (dqk + 15) / 16;
  p.dv_pad = dv_pad;
  p.vt_rows = vt_rows;
  p.sum_col = a->sum_row ? a->d : -1;
  p.p_f16 = a->p_f16;
  const int q_bytes = p.dchunks * ATT_CHUNK_BYTES;
  const int stage_bytes = p.dchunks * ATT_CHUNK_BYTES + 2 * dv_pad * 128;
  static bool configured[64] = {false};
  {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           227 * 1024);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attn2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SDB_ERR_CUDA; }
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
  }
  // two 128-query tiles per CTA when the accumulators fit (d <= 128) and there is more than one tile
  int stages2 = (227 * 1024 - 2 * q_bytes - 1024 - 256) / stage_bytes;
  if (stages2 > ATT2_MAX_STAGES) stages2 = ATT2_MAX_STAGES;
  p.stages = stages2;
  p.exp_poly = (a->exp_poly == 1) ? 0 : 1;
  const int smem2 = stages2 >= 2 ? 2 * q_bytes + stages2 * stage_bytes + 1024 + 256 : (1 << 30);
  // a single key block (cross-attention): two light one-tile CTAs per SM beat one two-tile CTA
  const bool single_block = (a->Skv <= ATT_BKV) && dv_pad <= 128 && !a->sum_row;
  p.tmem_cols = single_block ? 256 : 512;
  p.o_col = single_block ? 128 : 256;
  const bool two_tile = !a->causal && a->d <= 128 && (a->S > ATT_BQ || a->sum_row) && smem2 <= 227 * 1024 &&
                        (a->variant != 1 || a->sum_row) && !(single_block && a->variant != 2);
  if (a->qk_fold) {
    if (!two_tile || !a->sum_row || !a->q_prescaled || !p.exp_poly || a->p_f16 || dv_pad > 64 || dqk < a->d + 1 ||
        (a->d >> 6) != ((dqk - 1) >> 6)) {
      set_error("sdb_attention: qk_fold needs the two-tile kernel with sum_row, q_prescaled, the polynomial quarter, "
                "V^T rows per head <= 64 and qk_cols > d in the same 64-column chunk");
      return SDB_ERR_UNSUPPORTED;
    }
    p.qk_fold = 1;
  }
  if (two_tile) {
    dim3 grid((unsigned)((a->S + 2 * ATT_BQ - 1) / (2 * ATT_BQ)), (unsigned)a->heads, (unsigned)a->NB);
    (void)launch_k(attn2_tc_kernel, grid, dim3(ATT2_THREADS), (size_t)smem2, (cudaStream_t)stream, 1, p);
    return check_launch("attn2_tc_kernel");
  }
  if (a->sum_row) { set_error("sdb_attention: sum_row does not fit the two-tile kernel for d = %d", a->d); return SDB_ERR_UNSUPPORTED; }
  // single key block, no mask (cross-attention): a CTA walks several query tiles of its (sample, head)
  const int q_tiles = (a->S + ATT_BQ - 1) / ATT_BQ;
  p.tiles_per_cta = 1;
  {
    static int multi = -1;
    if (multi < 0) { const char* ev = getenv("SDB_ATTN_MULTI_TILE"); multi = (ev && ev[0] == '0') ? 0 : 1; }
    if (multi && single_block && !a->causal && q_tiles >= 2)
      p.tiles_per_cta = q_tiles >= 16 ? 8 : (q_tiles >= 8 ? 4 : 2);
  }
  const int smem_bytes = q_bytes + (single_block ? 1 : 2) * stage_bytes + 1024 + 256;
  if (smem_bytes > 227 * 1024) { set_error("sdb_attention: shared memory %d too large", smem_bytes); return SDB_ERR_UNSUPPORTED; }
  dim3 grid((unsigned)((q_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta), (unsigned)a->heads, (unsigned)a->NB);
  (void)launch_k(attn_tc_kernel, grid, dim3(ATT_THREADS), (size_t)smem_bytes, (cudaStream_t)stream, 1, p);
  return check_launch("attn_tc_kernel");
}
