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
  int dv_pad;    // d rounded up to 16
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
  const int h = blockIdx.y;
  const int n = blockIdx.z;

  const int q_bytes = p.dchunks * ATT_CHUNK_BYTES;
  const int k_bytes = p.dchunks * ATT_CHUNK_BYTES;
  const int v_chunk_bytes = p.dv_pad * 128;
  const int stage_bytes = k_bytes + 2 * v_chunk_bytes;
  uint8_t* q_smem = smem;
  uint8_t* kv_smem = smem + q_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(kv_smem + 2 * stage_bytes);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;    // [2]
  uint64_t* p_full = bars + 7;    // [2]
  uint64_t* o_done = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

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
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 256;

  if (warp == 0) {
    // ===================== TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, (uint32_t)q_bytes);
      for (int c = 0; c < p.dchunks; ++c)
        tma_load_4d(&p.map_q, q_full, q_smem + c * ATT_CHUNK_BYTES, c * 64, h, q0, n);
    }
    for (int j = 0; j < nkv; ++j) {
      const int s = j & 1;
      if (j >= 2) mbar_wait(&kv_empty[s], ((j >> 1) - 1) & 1, 11);
      if (lane == 0) {
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
      if (lane == 0) {
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
      if (lane == 0) {
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
  if (warp == 1) tmem_dealloc(tmem_base, 512);
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
  {
    uint64_t dims[4] = {(uint64_t)a->d, (uint64_t)a->heads, (uint64_t)a->S, (uint64_t)a->NB};
    uint64_t str[3] = {(uint64_t)a->d * 2, (uint64_t)ldq * 2, (uint64_t)ldq * 2 * a->S};
    uint32_t box[4] = {64, 1, 128, 1};
    if ((rc = make_tmap_bf16(&p.map_q, a->q, 4, dims, str, box, "attention Q"))) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)a->d, (uint64_t)a->heads, (uint64_t)a->Skv_pad, (uint64_t)a->NB};
    uint64_t str[3] = {(uint64_t)a->d * 2, (uint64_t)ldk * 2, (uint64_t)ldk * 2 * a->Skv_pad};
    uint32_t box[4] = {64, 1, 128, 1};
    if ((rc = make_tmap_bf16(&p.map_k, a->k, 4, dims, str, box, "attention K"))) return rc;
  }
  const int dv_pad = ((a->d + 15) / 16) * 16;
  {
    uint64_t dims[3] = {(uint64_t)vt_ld, (uint64_t)a->NB, (uint64_t)a->heads * a->d};
    uint64_t str[2] = {(uint64_t)vt_ld * 2, (uint64_t)vt_ld * 2 * a->NB};
    uint32_t box[3] = {64, 1, (uint32_t)dv_pad};
    if ((rc = make_tmap_bf16(&p.map_vt, a->vt, 3, dims, str, box, "attention V^T"))) return rc;
  }
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.ldo = ldo;
  p.S = a->S; p.Skv = a->Skv; p.d = a->d; p.heads = a->heads; p.NB = a->NB;
  p.causal = a->causal;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.dchunks = (a->d + 63) / 64;
  p.dk_steps = (a->d + 15) / 16;
  p.dv_pad = dv_pad;
  const int q_bytes = p.dchunks * ATT_CHUNK_BYTES;
  const int stage_bytes = p.dchunks * ATT_CHUNK_BYTES + 2 * dv_pad * 128;
  const int smem_bytes = q_bytes + 2 * stage_bytes + 1024 + 256;
  if (smem_bytes > 227 * 1024) { set_error("sdb_attention: shared memory %d too large", smem_bytes); return SDB_ERR_UNSUPPORTED; }
  {
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           227 * 1024);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SDB_ERR_CUDA; }
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
  }
  dim3 grid((unsigned)((a->S + ATT_BQ - 1) / ATT_BQ), (unsigned)a->heads, (unsigned)a->NB);
  attn_tc_kernel<<<grid, ATT_THREADS, smem_bytes, (cudaStream_t)stream>>>(p);
  return check_launch("attn_tc_kernel");
}
