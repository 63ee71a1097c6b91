// EXPERIMENT, NOT BUILT INTO THE LIBRARY (round 2). Measured on a B200 (profiles/r02h_attn_half_row_trace.log):
// S = 4096, d = 40, batch 8: 741 us against 703 us for attn2_tc_kernel - SLOWER. With four softmax warps per
// sub-partition the exponential phase of a warp takes 1 254 clocks for 64 columns (attn2: 1 428 for 128): the
// sub-partition's issue bandwidth was already saturated by two warps inside the exponential loop (a 3-register
// fp32 instruction costs ~1.7 issue clocks there), the period of a key block stays ~2 450 clocks, and the extra
// named barrier per key block adds skew. One of its unit-test shapes also hung (bar.sync has no watchdog).
// Kept for the record of what was tried; to build it, include it from attn_tc.cu before sdb_attention.
// Two-tile flash attention with HALF-ROW softmax threads: the kernel for the UNet's full-resolution self-attention
// (narrow heads, d <= 47, thousands of keys: S = 4096 / 9216 at d = 40), where the exponentials bound everything.
//
// attn2_tc_kernel gives a query row to ONE thread: 8 softmax warps, two per sub-partition (one per query tile). Its
// per-key-block timeline (tools/attn_trace.py, round 2) reads: tcgen05.ld 330 clk, exponentials 1 430 clk, wait for
// the previous P.V + tcgen05.st + hand-over 440 clk, loop overhead 270 clk - a period of 2 470 clk of which the MUFU
// pipe of a sub-partition works 1 536 (2 warps x 96 ex2 x 8 clk): while one warp loads or stores, only ONE other warp
// is left to issue exponentials, and alone it keeps the pipe 65 % busy. Two MMA issuer warps did not change the period
// (the tensor pipe was never the limit). Here a row belongs to TWO threads (64 key columns each): 16 softmax warps,
// four per sub-partition, so that three warps can issue while one waits for TMEM. Everything else is attn2's
// protocol (TMA ring, S/P/O columns, ones-row denominator, pre-scaled queries, row offset folded into Q.K^T,
// polynomial quarter, one issuer warp per tile). What the split adds:
//   * key block 0: the row maximum needs both halves -> partial maxima through shared memory, one named barrier
//     per tile (once per query tile);
//   * every later block: "did any row of this tile grow by >= 2^8" must reach all 256 threads of the tile before P is
//     stored -> a flag in shared memory (three-deep rotation) and ONE named barrier per key block, which costs
//     nothing new: P.V cannot start before all 256 threads have stored P anyway. The rare growth case exchanges
//     the halves' maxima through shared memory and rescales P (both halves) and O (half 0) by the same power of two.
// Replaces sd/attention.py:55-76 for the shapes above; bit-level behaviour (reference offsets, rescale rule, bf16 P,
// denominator from the ones row) is attn2's qk_fold path.
#include "common.cuh"
#include "host.h"

namespace sdb {

constexpr int ATTH_SOFTMAX_WARPS = 16;
constexpr int ATTH_TMA_WARP = 16;
constexpr int ATTH_MMA_WARP = 17;                 // tile 0; tile 1 = warp 18
constexpr int ATTH_THREADS = 19 * 32;
constexpr int ATTH_XCH_BYTES = 2 * 2 * 128 * 4 + 64;   // partial row maxima [tile][half][row] + growth flags

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(ATTH_THREADS, 1)
attn2h_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();
  const int q0 = blockIdx.x * (2 * ATT_BQ);
  const int h = blockIdx.y;
  const int n = blockIdx.z;

  const int q_bytes = p.dchunks * ATT_CHUNK_BYTES;
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
  uint64_t* p_full = s_full + 2;                     // [2]
  uint64_t* o_done = p_full + 2;                     // [2]
  uint64_t* s_free = o_done + 2;                     // [2]
  uint64_t* q_ready = s_free + 2;                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 2);
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);      // [2][2][128]
  volatile uint32_t* gflag = reinterpret_cast<volatile uint32_t*>(xch + 2 * 2 * 128); // [2][3]
  // TMEM: S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O0 [384,448) O1 [448,512)
  const int nkv = p.Skv / ATT_BKV;
  const int trc = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? g_gemm_trace_on : 0;
#define ATTH_STAMP(j, slot) do { if (trc && (j) < GEMM_TRACE_TILES) g_gemm_trace[(j) * 8 + (slot)] = clock64(); } while (0)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_q);
    tma_prefetch_desc(&p.map_k);
    tma_prefetch_desc(&p.map_vt);
    mbar_init(q_full, 1);
    for (int i = 0; i < nstages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);          // both tiles' P.V over the stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 256);
      mbar_init(&o_done[i], 1);
      mbar_init(&s_free[i], 256);
      mbar_init(&q_ready[i], 128);
    }
    for (int i = 0; i < 6; ++i) gflag[i] = 0u;
    fence_mbar_init();
  }
  if (warp == ATTH_MMA_WARP) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == ATTH_TMA_WARP) {
    // ===================== TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, (uint32_t)(2 * q_bytes));
      for (int t = 0; t < 2; ++t)
        for (int c = 0; c < p.dchunks; ++c)
          tma_load_4d(&p.map_q, q_full, q_smem + t * q_bytes + c * ATT_CHUNK_BYTES, c * 64, h, q0 + t * ATT_BQ, n);
    }
    __syncwarp();
    int st = 0;
    uint32_t ph = 1;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&kv_empty[st], ph, 41);
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
  } else if (warp == ATTH_MMA_WARP || warp == ATTH_MMA_WARP + 1) {
    // ===================== MMA issuers: one warp per query tile
    const int t = warp - ATTH_MMA_WARP;
    const uint32_t idesc_qk = make_idesc_bf16(ATT_BQ, ATT_BKV);
    const uint32_t idesc_pv = make_idesc_bf16(ATT_BQ, (uint32_t)p.dv_pad);
    const uint32_t qa = smem_u32(q_smem) + (uint32_t)(t * q_bytes);
    const uint32_t kv_addr = smem_u32(kv_smem);
    const uint32_t s_tmem = tmem_base + (uint32_t)(t * 128);
    const uint32_t p_tmem = tmem_base + 256u + (uint32_t)(t * 64);
    const uint32_t o_tmem = tmem_base + 384u + (uint32_t)(t * 64);
    auto issue_qk = [&](int st) {
      const uint32_t ka = kv_addr + (uint32_t)(st * stage_bytes);
      if (elect_one()) {
        for (int ks = 0; ks < p.dk_steps; ++ks) {
          const uint32_t off = (uint32_t)((ks >> 2) * ATT_CHUNK_BYTES + (ks & 3) * 32);
          mma_ss(s_tmem, make_kmajor_sw128_desc(qa + off), make_kmajor_sw128_desc(ka + off), idesc_qk, ks > 0 ? 1u : 0u);
        }
        tc_commit(&s_full[t]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int st, bool first) {
      const uint32_t va = kv_addr + (uint32_t)(st * stage_bytes + k_bytes);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < ATT_BKV / 16; ++ks) {
          const uint32_t off = (uint32_t)((ks >> 2) * v_chunk_bytes + (ks & 3) * 32);
          mma_ts(o_tmem, p_tmem + (uint32_t)(ks * 8), make_kmajor_sw128_desc(va + off), idesc_pv,
                 (!first || ks > 0) ? 1u : 0u);
        }
        tc_commit(&o_done[t]);
        tc_commit(&kv_empty[st]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0, 42);
    mbar_wait(&kv_full[0], 0, 43);
    tc_fence_after();
    if (t == 1 && p.stagger < 0) {
      const long long t0 = clock64();
      while (clock64() - t0 < (long long)(-p.stagger)) { }
    }
    issue_qk(0);
    int st = 0;
    uint32_t ph_kv = 0;
    for (int j = 0; j < nkv; ++j) {
      int st_next = st + 1;
      uint32_t ph_next = ph_kv;
      if (st_next == nstages) { st_next = 0; ph_next ^= 1u; }
      const uint32_t par = (uint32_t)(j & 1);
      if (j + 1 < nkv) {
        // the next key block's scores only need S_t to have been read (P lives apart from S)
        mbar_wait(&kv_full[st_next], ph_next, 44);
        mbar_wait(&s_free[t], par, 45);
        if (j == 0) mbar_wait(&q_ready[t], 0, 46);           // the rows' offsets are in the Q tile
        tc_fence_after();
        issue_qk(st_next);
      }
      if (t == 0 && lane == 0) ATTH_STAMP(j, 7);
      mbar_wait(&p_full[t], par, 47);
      tc_fence_after();
      issue_pv(st, j == 0);
      if (t == 0 && lane == 0) ATTH_STAMP(j, 6);
      st = st_next;
      ph_kv = ph_next;
    }
  } else if (warp < ATTH_SOFTMAX_WARPS) {
    // ===================== softmax: warps 0-7 = columns [0, 64) of tiles 0 / 1, warps 8-15 = columns [64, 128)
    const int t = (warp >> 2) & 1;
    const int quad = warp & 3;
    const int hf = warp >> 3;
    const int row = quad * 32 + lane;
    const int qrow = q0 + t * ATT_BQ + row;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + (uint32_t)(t * 128 + hf * 64);
    const uint32_t p_addr = tmem_base + lane_addr + 256u + (uint32_t)(t * 64 + hf * 32);
    const uint32_t o_addr = tmem_base + lane_addr + 384u + (uint32_t)(t * 64);
    float* xrow = xch + (t * 2 + hf) * 128 + row;               // this thread's slot
    const float* xother = xch + (t * 2 + (hf ^ 1)) * 128 + row;   // the other half of the row
    const bool leader = (hf == 0 && quad == 0 && lane == 0);      // one thread per tile looks after the flags
    const bool tr0 = (t == 0 && hf == 0 && quad == 2 && lane == 0);
    float m_ref = 0.f;        // shift of this row relative to the reference baked into its Q tile (0 until it grows)
    for (int j = 0; j < nkv; ++j) {
      if (tr0) ATTH_STAMP(j, 0);
      mbar_wait(&s_full[t], (uint32_t)(j & 1), 48);
      if (tr0) ATTH_STAMP(j, 1);
      tc_fence_after();
      uint32_t sv[64];
      tmem_ld32(s_addr + 0, sv + 0);
      tmem_ld32(s_addr + 32, sv + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[t]);
      if (tr0) ATTH_STAMP(j, 2);
      uint32_t pk[32];
      if (j == 0) {
        // ---- key block 0: absolute scores (log2 units). Row maximum over both halves, reference = round(max) + 7
        // baked into column d of this row of the Q tile (see attn2_tc_kernel, qk_fold)
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(sv[i + 0]), __uint_as_float(sv[i + 1])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])));
          m2 = fmaxf(m2, fmaxf(__uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5])));
          m3 = fmaxf(m3, fmaxf(__uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7])));
        }
        const float mloc = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        *xrow = mloc;
        named_bar_sync(1 + t, 256);
        const float mrow = fmaxf(mloc, *xother);
        float mi = (mrow == -INFINITY) ? 0.f : rintf(mrow) + 7.0f;
        mi = fminf(fmaxf(mi, -256.f), 256.f);
        if (hf == 0) {
          const uint32_t qrow_addr = smem_u32(q_smem) + (uint32_t)(t * q_bytes + ((p.d >> 6) * ATT_CHUNK_BYTES)) +
                                     (uint32_t)(row * 128) + (uint32_t)(((((p.d & 63) >> 3) ^ (row & 7)) << 4) + (p.d & 7) * 2);
          const __nv_bfloat16 hv = __float2bfloat16_rn(-mi);
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(qrow_addr), "h"(*reinterpret_cast<const uint16_t*>(&hv)) : "memory");
          fence_proxy_async();
          mbar_arrive(&q_ready[t]);
        }
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          const float p0 = ex2_approx(__uint_as_float(sv[i]) - mi);
          const float p1 = ex2_approx(__uint_as_float(sv[i + 1]) - mi);
          const float p2 = ex2_approx(__uint_as_float(sv[i + 2]) - mi);
          const float p3 = exp2_poly(__uint_as_float(sv[i + 3]) - mi);
          pk[i >> 1] = pack_bf16x2(p0, p1);
          pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
        }
        if (tr0) ATTH_STAMP(j, 4);
        named_bar_sync(1 + t, 256);      // the exchange slots may be reused from here on
      } else {
        // ---- later blocks: scores arrive relative to the row's reference (+ 7): no maximum pass, no subtraction
        // unless the row has grown since (m_ref != 0)
        uint32_t orv = 0;
        if (__all_sync(0xffffffffu, m_ref == 0.f)) {
#pragma unroll
          for (int i = 0; i < 64; i += 4) {
            const float p0 = ex2_approx(__uint_as_float(sv[i]));
            const float p1 = ex2_approx(__uint_as_float(sv[i + 1]));
            const float p2 = ex2_approx(__uint_as_float(sv[i + 2]));
            const float p3 = exp2_poly(__uint_as_float(sv[i + 3]));
            pk[i >> 1] = pack_bf16x2(p0, p1);
            pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
            orv |= pk[i >> 1] | pk[(i >> 1) + 1];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 64; i += 4) {
            const float p0 = ex2_approx(__uint_as_float(sv[i]) - m_ref);
            const float p1 = ex2_approx(__uint_as_float(sv[i + 1]) - m_ref);
            const float p2 = ex2_approx(__uint_as_float(sv[i + 2]) - m_ref);
            const float p3 = exp2_poly(__uint_as_float(sv[i + 3]) - m_ref);
            pk[i >> 1] = pack_bf16x2(p0, p1);
            pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
            orv |= pk[i >> 1] | pk[(i >> 1) + 1];
          }
        }
        if (tr0) ATTH_STAMP(j, 4);
        // "some probability >= 2.0" (bit 14 of a bf16 pattern) = this block raised the row maximum by >= 2^8
        const bool grow = (orv & 0x40004000u) != 0u;
        if (leader) gflag[t * 3 + (j + 1) % 3] = 0u;             // recycled for block j + 1 (last read in block j - 2)
        if (__any_sync(0xffffffffu, grow) && lane == 0) gflag[t * 3 + j % 3] = 1u;
        mbar_wait(&o_done[t], (uint32_t)((j - 1) & 1), 49);      // P_t (and O_t) of the previous block are at rest
        tc_fence_after();
        named_bar_sync(1 + t, 256);
        if (gflag[t * 3 + j % 3] != 0u) {
          // rare: some row of the tile grew. Row maximum of P over both halves -> the same exact power of two for the
          // two halves of P and (half 0) for O, denominator column included; the row carries the shift from now on
          float pmax = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) pmax = fmaxf(pmax, fmaxf(bf16lo(pk[i]), bf16hi(pk[i])));
          *xrow = pmax;
          named_bar_sync(1 + t, 256);
          const float rmax = fmaxf(pmax, *xother);
          if (rmax >= 2.0f) {
            const float e = floorf(log2f(rmax)) + 7.0f;
            const float factor = ex2_approx(-e);
            if (hf == 0) {
              for (int c = 0; c < p.dv_pad; c += 16) {
                uint32_t ov[16];
                tmem_ld16(o_addr + (uint32_t)c, ov);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * factor);
                tmem_st16(o_addr + (uint32_t)c, ov);
              }
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) pk[i] = pack_bf16x2(bf16lo(pk[i]) * factor, bf16hi(pk[i]) * factor);
            m_ref += e;
          }
          named_bar_sync(1 + t, 256);     // exchange slots free again
        }
      }
      tmem_st32(p_addr, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
      if (tr0) ATTH_STAMP(j, 5);
    }
    // ---- epilogue (half 0): O / l -> bf16, l = column d of O (ones row of V^T)
    mbar_wait(&o_done[t], (uint32_t)((nkv - 1) & 1), 50);
    tc_fence_after();
    if (hf == 0) {
      uint32_t lv[16];
      tmem_ld16(o_addr + (uint32_t)(p.sum_col & ~15), lv);
      tmem_ld_wait();
      float l = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i == (p.sum_col & 15)) l = __uint_as_float(lv[i]);
      const float inv = (l > 0.f) ? 1.0f / l : 0.f;
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == ATTH_MMA_WARP) tmem_dealloc(tmem_base, 512);
}

}  // namespace sdb
