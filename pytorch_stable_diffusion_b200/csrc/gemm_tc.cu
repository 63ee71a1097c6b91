// tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   out[m, n] = act( sum_k A[m, k] * W[n, k] + bias ) + residual[m, n]
//
// One kernel serves nn.Linear, 1x1 conv and 3x3 conv (stride 1, stride 2 with symmetric or
// right/bottom padding), reading NHWC bf16 activations:
//   * A tiles (128 output pixels x 64 input channels) are fetched by TMA straight from the NHWC
//     tensor, one box per filter tap; the zero padding of the convolution is the TMA
//     out-of-bounds fill, so no im2col buffer ever exists.
//   * A second source tensor extends the channel axis (UNet skip concatenation without a copy).
//   * W tiles (BLOCK_N output channels x 64) come from the packed [Cout][tap][Cin] bf16 matrix.
//   * tcgen05.mma (cta_group::1, M=128, N=BLOCK_N, K=16) accumulates fp32 in TMEM.
//   * Epilogue: tcgen05.ld -> +bias -> activation -> +residual -> bf16/fp32 store, or raw fp32
//     partial sums into a split-K workspace.
//
// Replaces the reference's nn.Conv2d / nn.Linear call sites (sd/diffusion.py:125,135,143,256,
// 266-269,410,545-569,712; sd/attention.py:12,16,143-152; sd/decoder.py:112-129,235-339;
// sd/encoder.py:56-92; sd/clip.py:117,121).
#include "common.cuh"
#include "host.h"
#include "../../include/sdb200.h"

namespace sdb {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_MAX_STAGES = 8;
constexpr int GEMM_A_STAGE_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KiB

struct GemmTcParams {
  CUtensorMap map_a0;
  CUtensorMap map_a1;
  CUtensorMap map_w;
  // output-pixel space and tiling
  int NB, HO, WO;          // output rows are (n, h, w), row index m = (n*HO + h)*WO + w
  int bw, bh, bn;          // tile box in (w, h, n); bw*bh*bn <= 128
  int tiles_w, tiles_h;    // tiles along w and h (tiles along n = gridDim.x / (tiles_w*tiles_h))
  int a_rank;              // 2: plain [M, K] matrix; 5: NHWC conv addressing
  // reduction
  int C0, C1;              // channels from source 0 / source 1 (multiples of 64 when C1 > 0)
  int cblocks0, cblocks;   // 64-channel blocks in source 0 / in both sources
  int ntaps;
  int ktot;                // row length of W = ntaps * (C0 + C1)
  int8_t tap_dc_sel[9];    // 0/1: add tap_dc_unit to the channel coordinate (stride-2 fold)
  int8_t tap_dw[9];
  int8_t tap_d2[9];
  int8_t tap_dh[9];
  int tap_dc_unit;
  // epilogue
  int N;                   // valid output columns (Cout)
  int block_n;             // UMMA N (multiple of 16, <= 256)
  int tmem_cols;           // power of two >= max(32, round_up(block_n, 32))
  int stages;
  int a_tx_bytes;           // bytes one A box delivers (bw*bh*bn rows of 128 B)
  int nsplit;              // split-K factor (gridDim.z)
  void* out;
  long long ldo;
  int out_fp32;
  const float* bias;
  int bias_mode;           // 0 none, 1 per column, 2 per row
  const void* residual;    // bf16, or fp32 when res_fp32
  int res_fp32;
  long long ldr;
  __nv_bfloat16* out2;     // optional bf16 copy of the output (same row stride), or nullptr
  int act;                 // 0 none, 1 quick-GELU, 2 SiLU
  float* workspace;        // [nsplit][M_total][N] fp32 when nsplit > 1
  long long m_total;
};

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == 1) return quick_gelu_f(x);
  if (act == 2) return silu_f(x);
  return x;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b_stage_bytes = p.block_n * GEMM_BK * 2;
  const int stage_bytes = GEMM_A_STAGE_BYTES + b_stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
  uint64_t* accum_bar = empty_bar + GEMM_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  // ---- tile coordinates
  const int n_tile = blockIdx.y;
  int mt = blockIdx.x;
  const int tw_i = mt % p.tiles_w;
  mt /= p.tiles_w;
  const int th_i = mt % p.tiles_h;
  const int tn_i = mt / p.tiles_h;
  const int w0 = tw_i * p.bw, h0 = th_i * p.bh, nb0 = tn_i * p.bn;
  const int n0 = n_tile * p.block_n;

  // ---- this CTA's slice of the reduction
  const int nkb_total = p.ntaps * p.cblocks;
  const int per_split = (nkb_total + p.nsplit - 1) / p.nsplit;
  const int kb_begin = blockIdx.z * per_split;
  const int kb_end = min(nkb_total, kb_begin + per_split);
  const int nkb = max(0, kb_end - kb_begin);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_a0);
    if (p.C1 > 0) tma_prefetch_desc(&p.map_a1);
    tma_prefetch_desc(&p.map_w);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
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

  if (warp == 0) {
    // ===== TMA producer (one lane issues; the warp stays converged on the waits)
    for (int i = 0; i < nkb; ++i) {
      const int s = i % p.stages;
      if (i >= p.stages) mbar_wait(&empty_bar[s], ((i / p.stages) - 1) & 1, 1);
      if (lane == 0) {
        const int kb = kb_begin + i;
        const int tap = kb / p.cblocks;
        const int cb = kb - tap * p.cblocks;
        uint8_t* a_dst = smem + s * stage_bytes;
        uint8_t* b_dst = a_dst + GEMM_A_STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(p.a_tx_bytes + b_stage_bytes));
        const bool second = cb >= p.cblocks0;
        const CUtensorMap* ma = second ? &p.map_a1 : &p.map_a0;
        const int c = (second ? (cb - p.cblocks0) : cb) * GEMM_BK;
        if (p.a_rank == 2) {
          tma_load_2d(ma, &full_bar[s], a_dst, c, w0);
        } else {
          tma_load_5d(ma, &full_bar[s], a_dst, c + p.tap_dc_sel[tap] * p.tap_dc_unit,
                      w0 + p.tap_dw[tap], p.tap_d2[tap], h0 + p.tap_dh[tap], nb0);
        }
        tma_load_2d(&p.map_w, &full_bar[s], b_dst, tap * (p.C0 + p.C1) + cb * GEMM_BK, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    const uint32_t idesc = make_idesc_bf16(GEMM_BM, (uint32_t)p.block_n);
    for (int i = 0; i < nkb; ++i) {
      const int s = i % p.stages;
      mbar_wait(&full_bar[s], (i / p.stages) & 1, 2);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
        const uint32_t b_addr = a_addr + GEMM_A_STAGE_BYTES;
        const uint64_t a_desc = make_kmajor_sw128_desc(a_addr);
        const uint64_t b_desc = make_kmajor_sw128_desc(b_addr);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          // advancing 16 bf16 (32 bytes) along K inside the 128B swizzle atom: +2 in (addr>>4)
          mma_ss(tmem_base, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                 (i > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(&empty_bar[s]);
        if (i == nkb - 1) tc_commit(accum_bar);
      }
      __syncwarp();
    }
  }

  // ===== Epilogue: all 8 warps. Warp w reads TMEM lanes [32*(w%4), +32), column chunks w/4, w/4+2, ...
  __syncwarp();
  if (nkb > 0) mbar_wait(accum_bar, 0, 3);
  tc_fence_after();

  const int q = warp & 3;
  const int r = q * 32 + lane;  // row of the tile owned by this thread
  const int dw = r % p.bw;
  const int dh = (r / p.bw) % p.bh;
  const int dn = r / (p.bw * p.bh);
  const int ww = w0 + dw, hh = h0 + dh, nn = nb0 + dn;
  const bool row_ok = (dn < p.bn) && (ww < p.WO) && (hh < p.HO) && (nn < p.NB);
  const long long m = ((long long)nn * p.HO + hh) * p.WO + ww;
  const int nchunks = (p.block_n + 31) / 32;
  const float row_bias = (p.bias_mode == 2 && row_ok) ? __ldg(p.bias + m) : 0.0f;

  for (int ch = (warp >> 2); ch < nchunks; ch += 2) {
    uint32_t v[32];
    if (nkb > 0) {
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0u;
    }
    const int col0 = n0 + ch * 32;
    const int ncol = min(32, min(p.block_n - ch * 32, p.N - col0));
    if (row_ok && ncol > 0) {
      if (p.nsplit > 1) {
        float* wsp = p.workspace + ((long long)blockIdx.z * p.m_total + m) * p.N + col0;
        if (ncol == 32 && ((reinterpret_cast<uintptr_t>(wsp) & 15u) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(wsp + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncol) wsp[j] = __uint_as_float(v[j]);
        }
      } else {
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
        if (p.bias_mode == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncol) x[j] += __ldg(p.bias + col0 + j);
        } else if (p.bias_mode == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] += row_bias;
        }
        if (p.act != 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = apply_act(x[j], p.act);
        }
        if (p.residual != nullptr && p.res_fp32) {
          const float* rp = reinterpret_cast<const float*>(p.residual) + m * p.ldr + col0;
          if (ncol == 32 && ((reinterpret_cast<uintptr_t>(rp) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 u = __ldg(reinterpret_cast<const float4*>(rp + j));
              x[j + 0] += u.x; x[j + 1] += u.y; x[j + 2] += u.z; x[j + 3] += u.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol) x[j] += rp[j];
          }
        } else if (p.residual != nullptr) {
          const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + m * p.ldr + col0;
          if (ncol == 32 && ((reinterpret_cast<uintptr_t>(rp) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint4 u = __ldg(reinterpret_cast<const uint4*>(rp + j));
              x[j + 0] += bf16lo(u.x); x[j + 1] += bf16hi(u.x);
              x[j + 2] += bf16lo(u.y); x[j + 3] += bf16hi(u.y);
              x[j + 4] += bf16lo(u.z); x[j + 5] += bf16hi(u.z);
              x[j + 6] += bf16lo(u.w); x[j + 7] += bf16hi(u.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol) x[j] += __bfloat162float(rp[j]);
          }
        }
        if (p.out_fp32) {
          float* op = reinterpret_cast<float*>(p.out) + m * p.ldo + col0;
          if (ncol == 32 && ((reinterpret_cast<uintptr_t>(op) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(op + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol) op[j] = x[j];
          }
        }
        __nv_bfloat16* bp = p.out_fp32 ? p.out2 : reinterpret_cast<__nv_bfloat16*>(p.out);
        if (bp != nullptr) {
          __nv_bfloat16* op = bp + m * p.ldo + col0;
          if (ncol == 32 && ((reinterpret_cast<uintptr_t>(op) & 15u) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 u;
              u.x = pack_bf16x2(x[j + 0], x[j + 1]);
              u.y = pack_bf16x2(x[j + 2], x[j + 3]);
              u.z = pack_bf16x2(x[j + 4], x[j + 5]);
              u.w = pack_bf16x2(x[j + 6], x[j + 7]);
              *reinterpret_cast<uint4*>(op + j) = u;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol) op[j] = __float2bfloat16_rn(x[j]);
          }
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// Split-K finalize: out = act(sum_z ws[z] + bias) + residual.
__global__ void gemm_splitk_finalize_kernel(const float* __restrict__ ws, int nsplit,
                                            long long m_total, int N, void* out, long long ldo,
                                            int out_fp32, const float* __restrict__ bias,
                                            int bias_mode, const void* __restrict__ residual,
                                            int res_fp32, long long ldr, int act,
                                            __nv_bfloat16* __restrict__ out2) {
  const long long total = m_total * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / N;
    const int n = (int)(i - m * N);
    float acc = 0.f;
    for (int z = 0; z < nsplit; ++z) acc += ws[(long long)z * total + i];
    if (bias_mode == 1) acc += bias[n];
    else if (bias_mode == 2) acc += bias[m];
    acc = apply_act(acc, act);
    if (residual) {
      acc += res_fp32 ? reinterpret_cast<const float*>(residual)[m * ldr + n]
                      : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(residual)[m * ldr + n]);
    }
    if (out_fp32) {
      reinterpret_cast<float*>(out)[m * ldo + n] = acc;
      if (out2) out2[m * ldo + n] = __float2bfloat16_rn(acc);
    } else {
      reinterpret_cast<__nv_bfloat16*>(out)[m * ldo + n] = __float2bfloat16_rn(acc);
    }
  }
}

static int pick_tile_box(int NB, int HO, int WO, int* bw, int* bh, int* bn) {
  // (bw, bh, bn) with bw*bh*bn <= 128 covering the (n, h, w) space in the fewest tiles; ties go to
  // the squarest patch (smallest 3x3 halo), then to the widest one.
  long long best_tiles = -1;
  int best_perim = 0;
  const int max_w = WO < 128 ? WO : 128;
  for (int cw = 1; cw <= max_w; ++cw) {
    for (int chh = 1; cw * chh <= 128 && chh <= HO; ++chh) {
      int cn = 1;
      if (cw == WO && chh == HO) {  // whole images: several of them may share one tile
        cn = 128 / (cw * chh);
        if (cn > NB) cn = NB;
        if (cn < 1) cn = 1;
      }
      const long long tiles = (long long)((WO + cw - 1) / cw) * ((HO + chh - 1) / chh) *
                              ((NB + cn - 1) / cn);
      const int perim = cw + chh;
      if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && perim < best_perim) ||
          (tiles == best_tiles && perim == best_perim && cw > *bw)) {
        best_tiles = tiles;
        best_perim = perim;
        *bw = cw; *bh = chh; *bn = cn;
      }
    }
  }
  return best_tiles > 0 ? 0 : -1;
}

static int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

}  // namespace sdb

extern "C" int sdb_gemm_tc(const sdb_gemm_args* a, void* stream_) {
  using namespace sdb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->a0 || !a->w || !a->out) { set_error("sdb_gemm_tc: null pointer"); return SDB_ERR_ARG; }
  if (a->Cout <= 0 || a->C0 <= 0) { set_error("sdb_gemm_tc: bad channel counts"); return SDB_ERR_ARG; }
  const int kind = a->kind;
  if (kind < SDB_GEMM_LINEAR || kind > SDB_GEMM_CONV3X3_S2_PAD_RB) {
    set_error("sdb_gemm_tc: unknown kind %d", kind);
    return SDB_ERR_ARG;
  }
  if (a->C1 > 0 && (a->C0 % 64 != 0 || a->C1 % 64 != 0 || !a->a1)) {
    set_error("sdb_gemm_tc: dual-source needs C0 %% 64 == 0 and C1 %% 64 == 0");
    return SDB_ERR_UNSUPPORTED;
  }
  if (a->C0 % 8 != 0) { set_error("sdb_gemm_tc: C0 must be a multiple of 8"); return SDB_ERR_UNSUPPORTED; }

  GemmTcParams p;
  memset(&p, 0, sizeof(p));
  const int ctot = a->C0 + a->C1;
  p.C0 = a->C0; p.C1 = a->C1;
  p.cblocks0 = (a->C0 + 63) / 64;
  p.cblocks = p.cblocks0 + (a->C1 + 63) / 64;
  p.N = a->Cout;

  int rc;
  if (kind == SDB_GEMM_LINEAR) {
    p.a_rank = 2;
    p.NB = 1; p.HO = 1; p.WO = a->M;
    p.bw = 128; p.bh = 1; p.bn = 1;
    p.ntaps = 1;
    if (a->M <= 0) { set_error("sdb_gemm_tc: M <= 0"); return SDB_ERR_ARG; }
    const long long lda0 = a->lda0 ? a->lda0 : a->C0;
    const long long lda1 = a->lda1 ? a->lda1 : a->C1;
    uint64_t dims[2] = {(uint64_t)a->C0, (uint64_t)a->M};
    uint64_t str[1] = {(uint64_t)lda0 * 2};
    uint32_t box[2] = {64, 128};
    if ((rc = make_tmap_bf16(&p.map_a0, a->a0, 2, dims, str, box, "gemm A0"))) return rc;
    if (a->C1 > 0) {
      uint64_t dims1[2] = {(uint64_t)a->C1, (uint64_t)a->M};
      uint64_t str1[1] = {(uint64_t)lda1 * 2};
      if ((rc = make_tmap_bf16(&p.map_a1, a->a1, 2, dims1, str1, box, "gemm A1"))) return rc;
    }
  } else {
    p.a_rank = 5;
    const int NB = a->NB, HI = a->HI, WI = a->WI;
    if (NB <= 0 || HI <= 0 || WI <= 0) { set_error("sdb_gemm_tc: bad conv dims"); return SDB_ERR_ARG; }
    const bool s2 = (kind != SDB_GEMM_CONV3X3_S1);
    if (s2 && (a->C1 > 0 || (HI & 1) || (WI & 1))) {
      set_error("sdb_gemm_tc: stride-2 conv needs even H, W and a single source");
      return SDB_ERR_UNSUPPORTED;
    }
    if (a->C0 % 64 != 0) { set_error("sdb_gemm_tc: conv needs C0 %% 64 == 0"); return SDB_ERR_UNSUPPORTED; }
    p.NB = NB; p.HO = s2 ? HI / 2 : HI; p.WO = s2 ? WI / 2 : WI;
    p.ntaps = 9;
    if (pick_tile_box(p.NB, p.HO, p.WO, &p.bw, &p.bh, &p.bn)) { set_error("tile box"); return SDB_ERR_ARG; }
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int t = ky * 3 + kx;
        if (!s2) {
          p.tap_dc_sel[t] = 0; p.tap_dw[t] = (int8_t)(kx - 1); p.tap_d2[t] = 0; p.tap_dh[t] = (int8_t)(ky - 1);
        } else if (kind == SDB_GEMM_CONV3X3_S2) {  // input index = 2*o + k - 1
          p.tap_dc_sel[t] = (kx != 1); p.tap_dw[t] = (int8_t)(kx == 0 ? -1 : 0);
          p.tap_d2[t] = (ky != 1); p.tap_dh[t] = (int8_t)(ky == 0 ? -1 : 0);
        } else {  // right/bottom padded: input index = 2*o + k
          p.tap_dc_sel[t] = (kx == 1); p.tap_dw[t] = (int8_t)(kx == 2 ? 1 : 0);
          p.tap_d2[t] = (ky == 1); p.tap_dh[t] = (int8_t)(ky == 2 ? 1 : 0);
        }
      }
    p.tap_dc_unit = a->C0;
    uint32_t box[5] = {64, (uint32_t)p.bw, 1, (uint32_t)p.bh, (uint32_t)p.bn};
    if (!s2) {
      const uint64_t c0b = (uint64_t)a->C0 * 2;
      uint64_t dims[5] = {(uint64_t)a->C0, (uint64_t)WI, 1, (uint64_t)HI, (uint64_t)NB};
      uint64_t str[4] = {c0b, c0b * WI, c0b * WI, c0b * WI * HI};
      if ((rc = make_tmap_bf16(&p.map_a0, a->a0, 5, dims, str, box, "conv A0"))) return rc;
      if (a->C1 > 0) {
        const uint64_t c1b = (uint64_t)a->C1 * 2;
        uint64_t dims1[5] = {(uint64_t)a->C1, (uint64_t)WI, 1, (uint64_t)HI, (uint64_t)NB};
        uint64_t str1[4] = {c1b, c1b * WI, c1b * WI, c1b * WI * HI};
        if ((rc = make_tmap_bf16(&p.map_a1, a->a1, 5, dims1, str1, box, "conv A1"))) return rc;
      }
    } else {
      // [N, H/2, 2, W/2, 2*C]: w parity folded into the channel axis, h parity its own axis.
      const uint64_t cb = (uint64_t)a->C0 * 2;
      uint64_t dims[5] = {(uint64_t)a->C0 * 2, (uint64_t)WI / 2, 2, (uint64_t)HI / 2, (uint64_t)NB};
      uint64_t str[4] = {cb * 2, cb * WI, cb * WI * 2, cb * WI * HI};
      if ((rc = make_tmap_bf16(&p.map_a0, a->a0, 5, dims, str, box, "conv-s2 A0"))) return rc;
    }
  }
  p.a_tx_bytes = (p.a_rank == 2) ? GEMM_A_STAGE_BYTES : p.bw * p.bh * p.bn * 128;
  p.tiles_w = (p.WO + p.bw - 1) / p.bw;
  p.tiles_h = (p.HO + p.bh - 1) / p.bh;
  const int tiles_n = (p.NB + p.bn - 1) / p.bn;
  const long long m_tiles = (long long)p.tiles_w * p.tiles_h * tiles_n;
  p.m_total = (long long)p.NB * p.HO * p.WO;
  p.ktot = p.ntaps * ctot;

  // ---- N tiling
  int block_n = a->block_n;
  if (block_n <= 0) {
    if (a->Cout % 256 == 0) block_n = 256;
    else if (a->Cout % 160 == 0) block_n = 160;
    else if (a->Cout % 128 == 0) block_n = 128;
    else if (a->Cout >= 256) block_n = 256;
    else block_n = ((a->Cout + 15) / 16) * 16;
  }
  if (block_n % 16 != 0 || block_n < 16 || block_n > 256) {
    set_error("sdb_gemm_tc: block_n %d invalid", block_n);
    return SDB_ERR_ARG;
  }
  p.block_n = block_n;
  p.tmem_cols = pow2_cols(((block_n + 31) / 32) * 32);
  const int n_tiles = (a->Cout + block_n - 1) / block_n;
  {
    uint64_t dims[2] = {(uint64_t)p.ktot, (uint64_t)a->Cout};
    const long long ldw = a->ldw ? a->ldw : p.ktot;
    uint64_t str[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {64, (uint32_t)block_n};
    if ((rc = make_tmap_bf16(&p.map_w, a->w, 2, dims, str, box, "gemm W"))) return rc;
  }

  // ---- pipeline depth from the shared-memory budget
  const int stage_bytes = GEMM_A_STAGE_BYTES + block_n * GEMM_BK * 2;
  const int smem_budget = a->smem_budget > 0 ? a->smem_budget : (block_n > 160 ? 220 * 1024 : 110 * 1024);
  int stages = (smem_budget - 2048) / stage_bytes;
  if (stages > GEMM_MAX_STAGES) stages = GEMM_MAX_STAGES;
  if (stages < 2) stages = 2;
  const int nkb_total = p.ntaps * p.cblocks;
  int nsplit = a->nsplit > 0 ? a->nsplit : 1;
  if (nsplit > nkb_total) nsplit = nkb_total;
  if (nsplit > 1 && !a->workspace) { set_error("sdb_gemm_tc: split-K needs a workspace"); return SDB_ERR_ARG; }
  p.nsplit = nsplit;
  p.stages = stages;
  const int smem_bytes = stages * stage_bytes + 1024 + 256;

  p.out = a->out;
  p.ldo = a->ldo ? a->ldo : a->Cout;
  p.out_fp32 = a->out_fp32;
  p.bias = a->bias;
  p.bias_mode = a->bias ? (a->bias_per_row ? 2 : 1) : 0;
  p.residual = a->residual;
  p.res_fp32 = a->res_fp32;
  p.out2 = reinterpret_cast<__nv_bfloat16*>(a->out2);
  if (a->out2 && !a->out_fp32) { set_error("sdb_gemm_tc: out2 (bf16 copy) needs out_fp32"); return SDB_ERR_ARG; }
  p.ldr = a->ldr ? a->ldr : a->Cout;
  p.act = a->act;
  p.workspace = a->workspace;

  {
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           227 * 1024);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SDB_ERR_CUDA; }
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
  }
  if (m_tiles > 2147483647LL || n_tiles > 65535) { set_error("sdb_gemm_tc: grid too large"); return SDB_ERR_UNSUPPORTED; }
  dim3 grid((unsigned)m_tiles, (unsigned)n_tiles, (unsigned)nsplit);
  gemm_tc_kernel<<<grid, GEMM_THREADS, smem_bytes, stream>>>(p);
  if ((rc = check_launch("gemm_tc_kernel"))) return rc;
  if (nsplit > 1) {
    const long long total = p.m_total * p.N;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    gemm_splitk_finalize_kernel<<<blocks, 256, 0, stream>>>(
        p.workspace, nsplit, p.m_total, p.N, p.out, p.ldo, p.out_fp32, p.bias, p.bias_mode,
        p.residual, p.res_fp32, p.ldr, p.act, p.out2);
    if ((rc = check_launch("gemm_splitk_finalize_kernel"))) return rc;
  }
  return SDB_OK;
}
