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
#include <cuda_fp16.h>
#include <algorithm>
#include "common.cuh"
#include "host.h"
#include "../../include/sdb200.h"

namespace sdb {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 8;                          // two per TMEM lane quadrant
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;     // warp 0 TMA, warp 1 MMA, warps 2.. epilogue
constexpr int GEMM_MAX_STAGES = 8;
constexpr int GEMM_A_STAGE_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KiB
constexpr int GEMM_EPI_STAGE_BYTES = 32 * 32 * 4;          // one 32x32 fp32 chunk per epilogue warp
constexpr int GEMM_RES_RING = 3;                           // residual chunks in flight per epilogue warp (+1)
constexpr int GEMM_BAR_BYTES = 512;                        // mbarriers + TMEM slot at the end of shared memory
constexpr int GEMM_EPI16_BYTES = 32 * 32 * 2;              // one 32x32 16-bit chunk
constexpr int GEMM_BIAS_LINES = 6;                         // TMA epilogue: 128-byte bias lines per warp (its <= 6 32-column chunks of a tile)

struct GemmTcParams {
  CUtensorMap map_a0;
  CUtensorMap map_a1;
  CUtensorMap map_w;
  // output-pixel space and tiling
  int NB, HO, WO;          // output rows are (n, h, w), row index m = (n*HO + h)*WO + w
  int bw, bh, bn;          // tile box in (w, h, n); bw*bh*bn <= 128
  int tiles_w, tiles_h;    // tiles along w and h
  int m_tiles, n_tiles;    // tiles along the row space / the output channels
  int total_tiles;         // m_tiles * n_tiles * nsplit
  int a_rank;              // 2: plain [M, K] matrix; 5: NHWC conv addressing
  // reduction
  int C0, C1;              // channels from source 0 / source 1 (multiples of 64 when C1 > 0)
  int cblocks0, cblocks;   // 64-channel blocks in source 0 / in both sources
  int ntaps;
  int ktot;                // row length of W = ntaps * (C0 + C1)
  int8_t tap_dc_sel[10];   // 0/1: add tap_dc_unit to the channel coordinate (stride-2 fold); entry 9 = centre tap
  int8_t tap_dw[10];
  int8_t tap_d2[10];
  int8_t tap_dh[10];
  // extra 1x1 source behind the nine taps of a stride-1 3x3 conv (the resblock's skip convolution accumulated into
  // the same TMEM tile: sd/diffusion.py:138-143,208): channels Cx0 (+ Cx1 from a second tensor), W columns
  // [9 * (C0 + C1), +Cx0 + Cx1)
  CUtensorMap map_x0;
  CUtensorMap map_x1;
  int cblocks_x0, cblocks_x;
  int nkb_total;           // ntaps * cblocks + cblocks_x
  int tap_dc_unit;
  // "Filter-column" staging (a3) of a stride-1 3x3 conv whose tile is a bw x bh patch of ONE sample: a pipeline
  // stage holds the A box of one filter column kx WITH its two halo rows - (bh + 2) x bw pixels x 64 channels,
  // fetched once at (w0 + kx - 1, h0 - 1), zero padding = TMA out-of-bounds fill - plus the three W sub-tiles
  // (ky = 0, 1, 2) of that column. Tap (ky, kx) is then the SAME shared-memory box viewed ky * bw rows further
  // down (ky * bw * 128 bytes: a multiple of the 1024-byte swizzle atom for bw % 8 == 0, so the UMMA descriptor
  // only moves its start address). 3 boxes instead of 9 per channel block: (bh + 2) / (3 bh) of the A bytes leave
  // L2 (0.42 for 16 x 8 patches) - the narrow pair tiles were bound by exactly that L2 -> SM traffic - and the
  // producer / issuer hand-shake runs once per three k-blocks.
  int a3;                  // 0 = one tap per stage (classic), 1 = filter-column staging
  int a3_nrow, a3_ncol;    // filter rows per column / filter columns: 3 x 3, or 2 x 2 for an up-sampling phase
  int a3_h0, a3_w0;        // offset of the first filter row / column from the output pixel: -1 / -1, or a - 1 / b - 1
  int a3_box_bytes;        // (bh + a3_nrow - 1) * bw * 128
  int a3_iters;            // pipeline stages consumed per tile: 3 * cblocks + ceil(cblocks_x / 2)
  int a3_stage_bytes;      // max(box + 3 W sub-tiles, 2 x (classic A tile + W sub-tile)): the extra 1x1 source's k-blocks
                           // travel TWO per stage ([A0][A1][W0][W1]) - one per stage left only `stages` k-blocks in
                           // flight and the extra phase ran at half the speed of the taps
  // epilogue
  int N;                   // valid output columns (Cout)
  int block_n;             // UMMA N (multiple of 16, <= 256)
  int acc_stride;          // TMEM columns between the two accumulator buffers (power of two >= block_n)
  int tmem_cols;           // 2 * acc_stride (512 in wide mode)
  int n_acc;               // 1: one UMMA of N = block_n per k-step, two TMEM buffers (tile i+1 accumulates while
                           //    tile i drains). 2 ("wide"): block_n = 2 * acc_n, two UMMAs per k-step sharing the
                           //    A tile, accumulators at TMEM columns 0 and 256, ONE buffer: fewer TMA bytes per
                           //    MMA cycle (the narrow 256 x 160 pair tile is TMA-bound), epilogue not overlapped
  int acc_n;               // UMMA N (= block_n / n_acc)
  int stages;
  int a_tx_bytes;          // bytes one A box delivers (bw*bh*bn rows of 128 B)
  int nsplit;              // split-K factor
  int per_split;           // k-blocks per split
  void* out;
  long long ldo;
  int out_fp32;
  int out_f16;             // the 16-bit tensor written (out, or out2 next to an fp32 out) is IEEE half instead of bf16
  int ab_f16;              // A and W are IEEE half (kind::f16 with f16 operand formats); accumulation stays fp32
  const float* bias;
  int bias_mode;           // 0 none, 1 per column, 2 per row
  const void* residual;    // bf16 (res_fp32 = 0), fp32 (1) or IEEE half (2)
  int res_fp32;
  int res_async;           // fp32 residual streamed through a per-warp cp.async ring in shared memory
  int res_direct;          // fp32 residual, 16-byte aligned, fetched with vector loads per chunk
  long long ldr;
  __nv_bfloat16* out2;     // optional bf16 copy of the output (same row stride), or nullptr
  int act;                 // 0 none, 1 quick-GELU, 2 SiLU
  float* workspace;        // [nsplit][M_total][N] fp32 when nsplit > 1
  long long m_total;
  uint32_t fd_mul[4], fd_shr[4];   // fast division by n_tiles, m_tiles, tiles_w, tiles_h (decode_tile)
  // GroupNorm statistics of the OUTPUT, produced by the epilogue: per (sample, 32-row slab, channel) {sum, sum of
  // squares} of the final fp32 values -> gn_part[n][k][N][2]; a small kernel reduces them per group, so the
  // consumer's statistics pass (one more read of the tensor) disappears. nullptr = off.
  float* gn_part;
  int gn_K;                // slabs per sample
  int gn_hw;               // rank-2 outputs: rows per sample (multiple of 32)
  int gn_spq;              // conv outputs: slabs of one sample inside a tile = min(4, bw*bh/32)
  int up_all;              // CONV2X2_UP: all four parity phases in ONE launch - the tile's z index is the phase (up_phase = 4)
  int gn_stride, gn_off;   // slab index = sample * gn_stride + gn_off + ...: gn_K and 0, or 4 * gn_K and phase * gn_K for
                           // the four phase launches of an up-sampling conv that share one partial-sum tensor
  // Nearest-neighbour x2 up-sampling folded into the following 3x3 conv (SDB_GEMM_CONV2X2_UP): the output pixels of
  // one parity (2y + a, 2x + b) are a 2x2 convolution of the LOW-resolution input (taps (a - 1 + u, b - 1 + v)) with
  // the 3x3 taps that fall on the same input pixel summed at pack time. The tile space is the low-resolution grid; the
  // epilogue scatters row (n, y, x) to row ((n*HO + y)*2 + a) * 2*WO + 2*x + b of the high-resolution tensor.
  int up, up_a, up_b;
  // TMA epilogue (short reductions, where the epilogue bounds the tile time): thread = row, result staged in
  // swizzled shared memory and written with cp.async.bulk.tensor stores, fp32 residual fetched by TMA loads
  int epi_tma;
  int epi_bytes;           // shared memory of all epilogue warps (either path)
  int epi_warp_bytes;      // per warp: epi_nslot fp32 slots of 4 KiB, then two 2 KiB 16-bit buffers (if any)
  int epi_nslot;           // fp32 slots per warp (residual in, result out, in place): 0, 2, 3 or 4
  int epi_w64;             // 16-bit-only outputs: a warp owns PAIRS of 32-column chunks and stores 64-column boxes (128-byte rows)
  int8_t slab_w0[4], slab_h0[4], slab_n0[4];   // origin of TMEM lane quadrant q's 32 rows inside the tile box
  int8_t slab_ok[4];       // quadrant holds rows of the tile at all
  CUtensorMap map_out;     // boxes of 32 columns x 32 rows (rank 2: [M][N]; rank 4: [NB][HO][WO][N])
  CUtensorMap map_out2;
  CUtensorMap map_res;
};

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == 1) return quick_gelu_f(x);
  if (act == 2) return silu_f(x);
  return x;
}

struct TileCoord {
  int n0, w0, h0, nb0, z, kb_begin, nkb;
  int sp;                  // index of the tile's (w, h) patch inside a sample: th * tiles_w + tw
};

// Debug timeline (sdb_debug_gemm_trace): CTA 0 records SM-clock stamps per tile, 8 slots each:
// 0 producer tile start, 1 producer all loads issued, 2 issuer before accumulator wait, 3 issuer first
// operands landed, 4 issuer last MMA committed, 5 epilogue waiting, 6 accumulator ready, 7 tile stored.
constexpr int GEMM_TRACE_TILES = 64;
__device__ long long g_gemm_trace[GEMM_TRACE_TILES * 8];
__device__ int g_gemm_trace_on = 0;
// `trc` is the trace mode read ONCE per kernel (0 for every CTA but the first): the stamps sit next to
// the single-thread issue loops, where a global load per call would itself distort the timeline.
__device__ __forceinline__ void trace_stamp(int trc, int tile_local, int slot) {
  if (trc && !(trc & 24) && tile_local < GEMM_TRACE_TILES) g_gemm_trace[tile_local * 8 + slot] = clock64();
}
// mode 16: warp 2 / lane 0, the SECOND chunk it processes in every tile (steady state of the TMA epilogue's chain):
// 0 previous chunk's store issued, 1 buffers free + residual requested, 2 tcgen05.ld landed, 3 accumulator released (last chunk),
// 4 rows done, 5 16-bit rows in shared memory, 6 proxy fence + warp sync, 7 store issued
__device__ __forceinline__ void trace_m16(int trc, int tile_local, int slot) {
  if ((trc & 16) && tile_local < GEMM_TRACE_TILES) g_gemm_trace[tile_local * 8 + slot] = clock64();
}
// mode 8: epilogue-internal stamps of warp 2 / lane 0 for the first chunk of every tile
__device__ __forceinline__ void trace_epi(int trc, int tile_local, int slot) {
  if ((trc & 8) && tile_local < GEMM_TRACE_TILES) g_gemm_trace[tile_local * 8 + slot] = clock64();
}

// Tile order: output-channel tile fastest, so CTAs that run at the same time share the activation
// tile through L2 and the (smaller) weight matrix stays L2-resident. In pair mode (CG = 2) a "tile"
// is two consecutive row tiles; CTA rank r of the pair owns row tile 2*t + r.
// x / d for 0 <= x < 2^31 with a host-computed multiplier (d = 1: mul = 0): one IMAD.HI and a shift instead of
// the ~40-instruction software division - decode_tile sits on the critical path of every role once per tile.
__device__ __forceinline__ int fast_div(int x, uint32_t mul, uint32_t shr) {
  return mul ? (int)(__umulhi((uint32_t)x, mul) >> shr) : x;
}
template <int CG>
__device__ __forceinline__ TileCoord decode_tile(const GemmTcParams& p, int tile, int cta_rank) {
  TileCoord t;
  int r = fast_div(tile, p.fd_mul[0], p.fd_shr[0]);
  const int n_tile = tile - r * p.n_tiles;
  t.z = fast_div(r, p.fd_mul[1], p.fd_shr[1]);
  int mt = (r - t.z * p.m_tiles) * CG + cta_rank;
  int q = fast_div(mt, p.fd_mul[2], p.fd_shr[2]);
  const int tw_i = mt - q * p.tiles_w;
  const int tn_i = fast_div(q, p.fd_mul[3], p.fd_shr[3]);
  const int th_i = q - tn_i * p.tiles_h;
  t.sp = th_i * p.tiles_w + tw_i;
  t.w0 = tw_i * p.bw;
  t.h0 = th_i * p.bh;
  t.nb0 = tn_i * p.bn;
  t.n0 = n_tile * p.block_n;
  t.kb_begin = p.up_all ? 0 : t.z * p.per_split;       // up_all: z is the up-sampling phase, every tile reduces over all of k
  t.nkb = min(p.nkb_total, t.kb_begin + p.per_split) - t.kb_begin;
  return t;
}

// One accumulator row of a 32-column chunk in the TMA epilogue (thread = row): + bias (a broadcast 128-byte
// line in shared memory, plus the per-row value), activation, + fp32 residual read from the row's swizzled
// 16-byte pieces, fp32 result written back in place, 16-bit copy packed into pk. Compile-time variants keep
// the 32-element body free of branches.
template <bool ACT, bool RES, bool F32, bool F16>
__device__ __forceinline__ void epi_rows(const uint32_t (&v)[32], uint32_t (&pk)[16], uint32_t srow, uint32_t sx,
                                         uint32_t bias_line, float rowb, int act) {   // fp32 slots: 32-column chunks only
  // All shared-memory loads of a phase are issued back to back BEFORE their first use (the asm statements are
  // volatile, i.e. kept in program order: a load -> add -> load -> add sequence paid one shared-memory latency per
  // 4 columns - 477 clocks for this function on the epilogue's critical path, tools/gemm_trace.py --mode 8).
  float x[32];
  {
    float4 b4[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(b4[j].x), "=f"(b4[j].y), "=f"(b4[j].z), "=f"(b4[j].w)
                   : "r"(bias_line + (uint32_t)(j * 16)));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x[4 * j] = __uint_as_float(v[4 * j]) + (b4[j].x + rowb);
      x[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + (b4[j].y + rowb);
      x[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + (b4[j].z + rowb);
      x[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + (b4[j].w + rowb);
    }
  }
  if (ACT) {
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = apply_act(x[i], act);
  }
  if (RES) {
    float4 rr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(rr[j].x), "=f"(rr[j].y), "=f"(rr[j].z), "=f"(rr[j].w)
                   : "r"(srow + ((((uint32_t)j) ^ sx) << 4)));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x[4 * j] += rr[j].x; x[4 * j + 1] += rr[j].y; x[4 * j + 2] += rr[j].z; x[4 * j + 3] += rr[j].w;
    }
  }
  if (F32) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((((uint32_t)j) ^ sx) << 4)),
                   "f"(x[4 * j]), "f"(x[4 * j + 1]), "f"(x[4 * j + 2]), "f"(x[4 * j + 3])
                   : "memory");
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (F16) { pk[2 * j] = pack_f16x2_sat(x[4 * j], x[4 * j + 1]); pk[2 * j + 1] = pack_f16x2_sat(x[4 * j + 2], x[4 * j + 3]); }
    else { pk[2 * j] = pack_bf16x2(x[4 * j], x[4 * j + 1]); pk[2 * j + 1] = pack_bf16x2(x[4 * j + 2], x[4 * j + 3]); }
  }
}

// The same row with an IEEE-half residual and a 16-bit result only (the token stream inside an attention block): the
// residual chunk is a 32 x 32 x 2-byte box in a 64B-swizzled slot (16-byte piece j of row r at (j ^ ((r >> 1) & 3))),
// or half of a 64-column row of a 128B-swizzled slot (piece (sub4 + j) ^ (r & 7)); nothing is written back.
template <bool ACT, bool F16>
__device__ __forceinline__ void epi_rows16(const uint32_t (&v)[32], uint32_t (&pk)[16], uint32_t rrow, uint32_t hx,
                                           uint32_t sub4, uint32_t bias_line, float rowb, int act) {
  uint32_t r[16];
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4)
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[4 * j4]), "=r"(r[4 * j4 + 1]), "=r"(r[4 * j4 + 2]), "=r"(r[4 * j4 + 3])
                 : "r"(rrow + (((sub4 + (uint32_t)j4) ^ hx) << 4)));
  float x[32];
  {
    float4 b4[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(b4[j].x), "=f"(b4[j].y), "=f"(b4[j].z), "=f"(b4[j].w)
                   : "r"(bias_line + (uint32_t)(j * 16)));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x[4 * j] = __uint_as_float(v[4 * j]) + (b4[j].x + rowb);
      x[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + (b4[j].y + rowb);
      x[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + (b4[j].z + rowb);
      x[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + (b4[j].w + rowb);
    }
  }
  if (ACT) {
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = apply_act(x[i], act);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&r[i]));
    x[2 * i] += h.x; x[2 * i + 1] += h.y;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (F16) { pk[2 * j] = pack_f16x2_sat(x[4 * j], x[4 * j + 1]); pk[2 * j + 1] = pack_f16x2_sat(x[4 * j + 2], x[4 * j + 3]); }
    else { pk[2 * j] = pack_bf16x2(x[4 * j], x[4 * j + 1]); pk[2 * j + 1] = pack_bf16x2(x[4 * j + 2], x[4 * j + 3]); }
  }
}

// Persistent, warp-specialised kernel. CG = 1: one CTA per SM computes 128 x block_n tiles.
// CG = 2: a cluster of two CTAs (one SM pair) computes 256 x block_n tiles with cta_group::2 MMAs -
// each CTA stages its own 128 rows of A and HALF of the W tile, so a pipeline stage holds fewer bytes
// per FLOP and the ring gets deeper for the same shared memory.
//   warp 0      TMA producer (A boxes + W tiles into a `stages`-deep ring)
//   warp 1      tcgen05.mma issuer (leader CTA only); two TMEM accumulators, so tile i+1 accumulates
//               while tile i drains
//   warps 2-9   epilogue: tcgen05.ld -> swizzled shared staging -> row-coalesced global I/O, with the
//               residual of the next chunk prefetched into registers
template <int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();
  const int trc = (blockIdx.x == 0) ? g_gemm_trace_on : 0;
  const int cta_rank = (CG == 2) ? (int)cluster_ctarank() : 0;
  const bool leader = (cta_rank == 0);
  const int first_tile = blockIdx.x / CG;
  const int tile_step = gridDim.x / CG;
  const int b_rows = p.acc_n / CG;                       // W rows this CTA stages per k-block and accumulator
  const int b_sub_bytes = b_rows * GEMM_BK * 2;
  const int b_stage_bytes = p.n_acc * b_sub_bytes;
  const int stage_bytes = p.a3 ? p.a3_stage_bytes : (GEMM_A_STAGE_BYTES + b_stage_bytes);
  const int nbuf = (p.n_acc == 2) ? 1 : 2;               // TMEM accumulator buffers
  uint8_t* epi_smem = smem + p.stages * stage_bytes;
  // [staging: one chunk per epilogue warp][residual ring: GEMM_RES_RING chunks per warp, if any][barriers]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + p.epi_bytes);
  uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + GEMM_MAX_STAGES;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;                // [2] accumulator drained (leader's copy is used)
  uint64_t* res_bar = tempty_bar + 2;                  // [GEMM_EPI_WARPS][4] residual slot landed (TMA epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + GEMM_EPI_WARPS * 4);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_a0);
    if (p.C1 > 0) tma_prefetch_desc(&p.map_a1);
    tma_prefetch_desc(&p.map_w);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], GEMM_EPI_WARPS * CG);
    }
    if (p.epi_tma) {
      for (int i = 0; i < GEMM_EPI_WARPS * 4; ++i) mbar_init(&res_bar[i], 1);
      tma_prefetch_desc(&p.map_out);
      if (p.out2 != nullptr) tma_prefetch_desc(&p.map_out2);
      if (p.residual != nullptr) tma_prefetch_desc(&p.map_res);
    }
    fence_mbar_init();
  }
  if constexpr (CG == 2) {
    cluster_sync_all();          // the peer's barriers exist before anything remote can reach them
    if (warp == 1) {
      tmem_alloc_pair(tmem_slot, (uint32_t)p.tmem_cols);
      tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
  } else {
    if (warp == 1) {
      tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
      tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const uint32_t tmem_base = *tmem_slot;
  // the prologue above touched no global data: only now wait for the predecessor grid (PDL)
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer. The whole warp walks the loop (so every value stays warp-uniform and lives
    // in uniform registers); one elected lane issues. Stage / phase / tap / channel block advance
    // incrementally - no divisions on this critical instruction chain.
    int s = 0;
    uint32_t ph = 1;                      // parity of the "slot is free" phase (fresh barriers pass)
    const uint32_t tx_bytes = (uint32_t)((p.a_tx_bytes + b_stage_bytes) * CG);
    const int ctot = p.C0 + p.C1;
    int ltp = 0;
    if (p.a3) {
      // ----- filter-column staging: per channel block three stages (kx = 0, 1, 2), each ONE halo box + three W
      // sub-tiles; the extra 1x1 source (if any) follows as classic one-tap stages
      const uint32_t tx3 = (uint32_t)((p.a3_box_bytes + p.a3_nrow * b_stage_bytes) * CG);
      const int nrow = p.a3_nrow, ncol = p.a3_ncol;
      for (int tile = first_tile; tile < p.total_tiles; tile += tile_step, ++ltp) {
        const TileCoord t = decode_tile<CG>(p, tile, cta_rank);
        trace_stamp(trc, ltp, 0);
        // up_all: phase z = 2a + b reads the input at rows a - 1 + u, columns b - 1 + v and its own Cout weight rows
        const int wn = t.n0 + cta_rank * b_rows + (p.up_all ? t.z * p.N : 0);
        const int a3h = p.up_all ? (t.z >> 1) - 1 : p.a3_h0, a3w = p.up_all ? (t.z & 1) - 1 : p.a3_w0;
        int cb = 0, kx = 0;
        bool extra = false;
        for (int i = 0; i < p.a3_iters; ++i) {
          mbar_wait(&empty_bar[s], ph, 1);
          uint8_t* a_dst = smem + s * stage_bytes;
          uint8_t* b_dst = a_dst + p.a3_box_bytes;
          const int cbl0 = extra ? p.cblocks_x0 : p.cblocks0;
          const bool second = cb >= cbl0;
          const CUtensorMap* ma = extra ? (second ? &p.map_x1 : &p.map_x0) : (second ? &p.map_a1 : &p.map_a0);
          const int c = (second ? (cb - cbl0) : cb) * GEMM_BK;
          if (elect_one()) {
            if (!extra) {
              if (leader) mbar_arrive_expect_tx(&full_bar[s], tx3);
              const int wk = kx * ctot + cb * GEMM_BK;                    // tap (ky, kx) starts at (ncol ky + kx) * ctot
              if constexpr (CG == 2) {
                tma_load_5d_pair(ma, &full_bar[s], a_dst, c, t.w0 + kx + a3w, 0, t.h0 + a3h, t.nb0);
                for (int ky = 0; ky < nrow; ++ky) {
                  tma_load_2d_pair(&p.map_w, &full_bar[s], b_dst + ky * b_stage_bytes, wk + ncol * ky * ctot, wn);
                  if (p.n_acc == 2)
                    tma_load_2d_pair(&p.map_w, &full_bar[s], b_dst + ky * b_stage_bytes + b_sub_bytes,
                                     wk + ncol * ky * ctot, wn + p.acc_n);
                }
              } else {
                tma_load_5d(ma, &full_bar[s], a_dst, c, t.w0 + kx + a3w, 0, t.h0 + a3h, t.nb0);
                for (int ky = 0; ky < nrow; ++ky) {
                  tma_load_2d(&p.map_w, &full_bar[s], b_dst + ky * b_stage_bytes, wk + ncol * ky * ctot, wn);
                  if (p.n_acc == 2)
                    tma_load_2d(&p.map_w, &full_bar[s], b_dst + ky * b_stage_bytes + b_sub_bytes, wk + ncol * ky * ctot,
                                wn + p.acc_n);
                }
              }
            } else {
              // extra 1x1 source: two k-blocks per stage, [A0][A1][W0][W1]
              const int nx = min(2, p.cblocks_x - cb);
              if (leader) mbar_arrive_expect_tx(&full_bar[s], tx_bytes * (uint32_t)nx);
              uint8_t* bx_dst = a_dst + 2 * GEMM_A_STAGE_BYTES;
              for (int e = 0; e < nx; ++e) {
                const int cbe = cb + e;
                const bool sec = cbe >= p.cblocks_x0;
                const CUtensorMap* mx = sec ? &p.map_x1 : &p.map_x0;
                const int cx = (sec ? (cbe - p.cblocks_x0) : cbe) * GEMM_BK;
                const int wk = p.ntaps * ctot + cbe * GEMM_BK;
                uint8_t* ad = a_dst + e * GEMM_A_STAGE_BYTES;
                uint8_t* bd = bx_dst + e * b_stage_bytes;
                if constexpr (CG == 2) {
                  tma_load_5d_pair(mx, &full_bar[s], ad, cx, t.w0, 0, t.h0, t.nb0);
                  tma_load_2d_pair(&p.map_w, &full_bar[s], bd, wk, wn);
                  if (p.n_acc == 2) tma_load_2d_pair(&p.map_w, &full_bar[s], bd + b_sub_bytes, wk, wn + p.acc_n);
                } else {
                  tma_load_5d(mx, &full_bar[s], ad, cx, t.w0, 0, t.h0, t.nb0);
                  tma_load_2d(&p.map_w, &full_bar[s], bd, wk, wn);
                  if (p.n_acc == 2) tma_load_2d(&p.map_w, &full_bar[s], bd + b_sub_bytes, wk, wn + p.acc_n);
                }
              }
            }
          }
          __syncwarp();
          if (!extra) {
            if (++kx == ncol) { kx = 0; if (++cb == p.cblocks) { cb = 0; extra = true; } }
          } else {
            cb += 2;
          }
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        trace_stamp(trc, ltp, 1);
      }
    } else
    for (int tile = first_tile; tile < p.total_tiles; tile += tile_step, ++ltp) {
      const TileCoord t = decode_tile<CG>(p, tile, cta_rank);
      trace_stamp(trc, ltp, 0);
      int tap = t.kb_begin / p.cblocks;
      int cb = t.kb_begin - tap * p.cblocks;
      if (tap >= p.ntaps) { tap = p.ntaps; cb = t.kb_begin - p.ntaps * p.cblocks; }    // inside the extra 1x1 source
      const int wn = t.n0 + cta_rank * b_rows + (p.up_all ? t.z * p.N : 0);
      const int up_dh = p.up_all ? (t.z >> 1) : 0, up_dw = p.up_all ? (t.z & 1) : 0;   // tap_dh / tap_dw hold phase 0's offsets
      for (int i = 0; i < t.nkb; ++i) {
        mbar_wait(&empty_bar[s], ph, 1);
        uint8_t* a_dst = smem + s * stage_bytes;
        uint8_t* b_dst = a_dst + GEMM_A_STAGE_BYTES;
        const bool extra = (tap == p.ntaps);                   // k-blocks of the extra 1x1 source (centre tap)
        const int cbl0 = extra ? p.cblocks_x0 : p.cblocks0;
        const bool second = cb >= cbl0;
        const CUtensorMap* ma = extra ? (second ? &p.map_x1 : &p.map_x0) : (second ? &p.map_a1 : &p.map_a0);
        const int c = (second ? (cb - cbl0) : cb) * GEMM_BK;
        const int wk = tap * ctot + cb * GEMM_BK;
        const int ca = c + p.tap_dc_sel[tap] * p.tap_dc_unit;
        const int cw = t.w0 + p.tap_dw[tap] + up_dw;
        const int c2 = p.tap_d2[tap];
        const int chh = t.h0 + p.tap_dh[tap] + up_dh;
        if (elect_one()) {
          // the leader's barrier collects the bytes of both CTAs
          if (leader) mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
          if constexpr (CG == 2) {
            if (p.a_rank == 2) tma_load_2d_pair(ma, &full_bar[s], a_dst, c, t.w0);
            else tma_load_5d_pair(ma, &full_bar[s], a_dst, ca, cw, c2, chh, t.nb0);
            tma_load_2d_pair(&p.map_w, &full_bar[s], b_dst, wk, wn);
            if (p.n_acc == 2) tma_load_2d_pair(&p.map_w, &full_bar[s], b_dst + b_sub_bytes, wk, wn + p.acc_n);
          } else {
            if (p.a_rank == 2) tma_load_2d(ma, &full_bar[s], a_dst, c, t.w0);
            else tma_load_5d(ma, &full_bar[s], a_dst, ca, cw, c2, chh, t.nb0);
            tma_load_2d(&p.map_w, &full_bar[s], b_dst, wk, wn);
            if (p.n_acc == 2) tma_load_2d(&p.map_w, &full_bar[s], b_dst + b_sub_bytes, wk, wn + p.acc_n);
          }
        }
        __syncwarp();
        if (++cb == (extra ? p.cblocks_x : p.cblocks)) { cb = 0; ++tap; }
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      trace_stamp(trc, ltp, 1);
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp 1 of the leader CTA, one elected lane issues
    if (leader) {
      const uint32_t idesc = make_idesc_16(GEMM_BM * CG, (uint32_t)p.acc_n, p.ab_f16);
      const uint64_t b_half = (uint64_t)(b_sub_bytes >> 4);
      const bool wide = (p.n_acc == 2);
      const uint64_t a_desc0 = make_kmajor_sw128_desc(smem_u32(smem));
      const uint64_t b_desc0 = make_kmajor_sw128_desc(smem_u32(smem) + GEMM_A_STAGE_BYTES);
      const uint32_t stage_step = (uint32_t)(stage_bytes >> 4);
      int s = 0;
      uint32_t ph = 0;
      uint32_t soff = 0;                     // s * stage_bytes >> 4
      int lt = 0;
      for (int tile = first_tile; tile < p.total_tiles; tile += tile_step, ++lt) {
        const TileCoord t = decode_tile<CG>(p, tile, cta_rank);
        const int acc = wide ? 0 : (lt & 1);
        trace_stamp(trc, lt, 2);
        if (lt >= nbuf) mbar_wait(&tempty_bar[acc], (uint32_t)((lt / nbuf) - 1) & 1u, 4);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
        uint32_t accum = 0;
        if (p.a3) {
          // filter-column staging: a stage holds ONE halo box and the W sub-tiles of ky = 0, 1, 2; tap (ky, kx)
          // reads the box ky * bw rows further down. The extra 1x1 source follows as one-tap stages.
          const uint64_t b3_desc0 = make_kmajor_sw128_desc(smem_u32(smem) + (uint32_t)p.a3_box_bytes);
          const uint64_t a_row_step = (uint64_t)((p.bw * 128) >> 4);
          const uint64_t b_tap_step = (uint64_t)(b_stage_bytes >> 4);
          const int n3 = p.a3_ncol * p.cblocks;
          const uint64_t bx_desc0 = make_kmajor_sw128_desc(smem_u32(smem) + 2u * GEMM_A_STAGE_BYTES);
          const uint64_t ax_step = (uint64_t)(GEMM_A_STAGE_BYTES >> 4);
          for (int i = 0; i < p.a3_iters; ++i) {
            mbar_wait(&full_bar[s], ph, 2);
            if (i == 0) trace_stamp(trc, lt, 3);
            tc_fence_after();
            const bool taps = i < n3;
            const int nsub = taps ? p.a3_nrow : min(2, p.cblocks_x - 2 * (i - n3));
            if (elect_one()) {
              for (int j = 0; j < nsub; ++j) {
                const uint64_t a_desc = a_desc0 + soff + (uint64_t)j * (taps ? a_row_step : ax_step);
                const uint64_t b_desc = (taps ? b3_desc0 : bx_desc0) + soff + (uint64_t)j * b_tap_step;
                if constexpr (CG == 2) {
                  mma_ss_pair(d_tmem, a_desc, b_desc, idesc, accum);
                  mma_ss_pair(d_tmem, a_desc + 2, b_desc + 2, idesc, 1u);
                  mma_ss_pair(d_tmem, a_desc + 4, b_desc + 4, idesc, 1u);
                  mma_ss_pair(d_tmem, a_desc + 6, b_desc + 6, idesc, 1u);
                  if (wide) {
                    mma_ss_pair(d_tmem + 256, a_desc, b_desc + b_half, idesc, accum);
                    mma_ss_pair(d_tmem + 256, a_desc + 2, b_desc + b_half + 2, idesc, 1u);
                    mma_ss_pair(d_tmem + 256, a_desc + 4, b_desc + b_half + 4, idesc, 1u);
                    mma_ss_pair(d_tmem + 256, a_desc + 6, b_desc + b_half + 6, idesc, 1u);
                  }
                } else {
                  mma_ss(d_tmem, a_desc, b_desc, idesc, accum);
                  mma_ss(d_tmem, a_desc + 2, b_desc + 2, idesc, 1u);
                  mma_ss(d_tmem, a_desc + 4, b_desc + 4, idesc, 1u);
                  mma_ss(d_tmem, a_desc + 6, b_desc + 6, idesc, 1u);
                  if (wide) {
                    mma_ss(d_tmem + 256, a_desc, b_desc + b_half, idesc, accum);
                    mma_ss(d_tmem + 256, a_desc + 2, b_desc + b_half + 2, idesc, 1u);
                    mma_ss(d_tmem + 256, a_desc + 4, b_desc + b_half + 4, idesc, 1u);
                    mma_ss(d_tmem + 256, a_desc + 6, b_desc + b_half + 6, idesc, 1u);
                  }
                }
                accum = 1;
              }
              if constexpr (CG == 2) tc_commit_pair(&empty_bar[s], 3);
              else tc_commit(&empty_bar[s]);
            }
            __syncwarp();
            accum = 1;
            soff += stage_step;
            if (++s == p.stages) { s = 0; ph ^= 1u; soff = 0; }
          }
        } else
        for (int i = 0; i < t.nkb; ++i) {
          mbar_wait(&full_bar[s], ph, 2);
          if (i == 0) trace_stamp(trc, lt, 3);
          tc_fence_after();
          const uint64_t a_desc = a_desc0 + soff;
          const uint64_t b_desc = b_desc0 + soff;
          if (elect_one()) {
            // advancing 16 bf16 (32 bytes) along K inside the 128B swizzle atom: +2 in (addr>>4)
            if constexpr (CG == 2) {
              mma_ss_pair(d_tmem, a_desc, b_desc, idesc, accum);
              mma_ss_pair(d_tmem, a_desc + 2, b_desc + 2, idesc, 1u);
              mma_ss_pair(d_tmem, a_desc + 4, b_desc + 4, idesc, 1u);
              mma_ss_pair(d_tmem, a_desc + 6, b_desc + 6, idesc, 1u);
              if (wide) {            // second accumulator (TMEM column 256): same A tile, next acc_n W rows
                mma_ss_pair(d_tmem + 256, a_desc, b_desc + b_half, idesc, accum);
                mma_ss_pair(d_tmem + 256, a_desc + 2, b_desc + b_half + 2, idesc, 1u);
                mma_ss_pair(d_tmem + 256, a_desc + 4, b_desc + b_half + 4, idesc, 1u);
                mma_ss_pair(d_tmem + 256, a_desc + 6, b_desc + b_half + 6, idesc, 1u);
              }
              tc_commit_pair(&empty_bar[s], 3);
            } else {
              mma_ss(d_tmem, a_desc, b_desc, idesc, accum);
              mma_ss(d_tmem, a_desc + 2, b_desc + 2, idesc, 1u);
              mma_ss(d_tmem, a_desc + 4, b_desc + 4, idesc, 1u);
              mma_ss(d_tmem, a_desc + 6, b_desc + 6, idesc, 1u);
              if (wide) {
                mma_ss(d_tmem + 256, a_desc, b_desc + b_half, idesc, accum);
                mma_ss(d_tmem + 256, a_desc + 2, b_desc + b_half + 2, idesc, 1u);
                mma_ss(d_tmem + 256, a_desc + 4, b_desc + b_half + 4, idesc, 1u);
                mma_ss(d_tmem + 256, a_desc + 6, b_desc + b_half + 6, idesc, 1u);
              }
              tc_commit(&empty_bar[s]);
            }
          }
          __syncwarp();
          accum = 1;
          soff += stage_step;
          if (++s == p.stages) { s = 0; ph ^= 1u; soff = 0; }
        }
        if (elect_one()) {
          if constexpr (CG == 2) tc_commit_pair(&tfull_bar[acc], 3);
          else tc_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        trace_stamp(trc, lt, 4);
      }
    }
  } else if (p.epi_tma) {
    // ===== Epilogue, TMA form. Warp e owns TMEM lanes [32*(warp%4), +32) - 32 rows of the tile that form a
    // rectangular box of the output tensor - and the 32-column chunks e/4, e/4 + 2, ... Thread = row: the
    // accumulator row arrives from tcgen05.ld, the fp32 residual row (TMA-loaded, 128B-swizzled slot) is added
    // and the result written back IN PLACE, a 16-bit copy goes to a 64B-swizzled buffer, and one lane issues the
    // bulk tensor stores. No per-lane global address arithmetic, no second pass through shared memory.
    const int e = warp - 2;
    const int q = warp & 3;
    const int half = e >> 2;
    const int nchunks = (p.block_n + 31) / 32;
    const int nslot = p.epi_nslot;
    const bool has_res = (p.residual != nullptr);
    const bool f32out = (p.out_fp32 != 0);
    const bool b16out = (!f32out) || (p.out2 != nullptr);
    const CUtensorMap* map16 = f32out ? &p.map_out2 : &p.map_out;
    const bool res16 = has_res && p.res_fp32 == 2;      // IEEE-half residual: 2 KB slots (64B swizzle), read only
    // w64 (16-bit result only): the unit of ownership, residual prefetch and store is a PAIR of 32-column chunks - one
    // 64-column box with 128-byte rows (128B swizzle) - so the per-store costs (proxy fence, warp sync, bulk-store
    // issue, read-completion wait) are paid once per 64 columns and every store writes whole 128-byte lines
    const bool w64 = p.epi_w64 != 0;
    const int cw = w64 ? 64 : 32;                        // columns per owned unit
    const int nunits = (p.block_n + cw - 1) / cw;
    const uint32_t b16_bytes = (uint32_t)GEMM_EPI16_BYTES * (w64 ? 2u : 1u);
    const uint32_t slot_bytes = res16 ? b16_bytes : (uint32_t)GEMM_EPI_STAGE_BYTES;
    const uint32_t row16 = w64 ? 128u : 64u;             // bytes per row of a 16-bit buffer / half residual slot
    const uint32_t key16 = w64 ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);   // its swizzle key for this lane's row
    const uint32_t wblk = smem_u32(epi_smem) + (uint32_t)(e * p.epi_warp_bytes);
    const uint32_t b16_base = wblk + (uint32_t)nslot * slot_bytes;
    const uint32_t rbar0 = smem_u32(res_bar + e * 4);
    const uint32_t bias_line = smem_u32(epi_smem) + (uint32_t)(GEMM_EPI_WARPS * p.epi_warp_bytes + e * (GEMM_BIAS_LINES * 128));
    const bool slab_ok = p.slab_ok[q] != 0;
    const int sw0 = p.slab_w0[q], sh0 = p.slab_h0[q], sn0 = p.slab_n0[q];
    // residual prefetch distance in units (fp32 slots: ahead | current | draining - the result is stored from the slot;
    // half slots are only read: ahead | current)
    const int pfd = (has_res && p.res_fp32 == 2) ? nslot - 1 : nslot - 2;
    // prefetch cursor: the same (tile, chunk) sequence as the main loop, pfd chunks ahead
    int pf_tile = first_tile, pf_lt = 0, pf_ch = half, pf_slot = 0;
    int pf_n0 = 0, pf_c1 = 0, pf_c2 = 0, pf_c3 = 0;
    auto pf_decode = [&]() {
      if (pf_tile < p.total_tiles) {
        const TileCoord t = decode_tile<CG>(p, pf_tile, cta_rank);
        pf_n0 = t.n0; pf_c1 = t.w0 + sw0; pf_c2 = t.h0 + sh0; pf_c3 = t.nb0 + sn0;
      }
    };
    auto pf_settle = [&]() {       // move to the next existing chunk at or after (pf_tile, pf_ch)
      while (pf_tile < p.total_tiles && !(pf_ch < nunits && pf_n0 + pf_ch * cw < p.N)) {
        pf_tile += tile_step; ++pf_lt; pf_ch = (half + pf_lt) & 1;
        pf_decode();
      }
    };
    auto issue_prefetch = [&]() {
      if (pf_tile < p.total_tiles) {
        if (lane == 0) {
          const uint32_t bar = rbar0 + (uint32_t)(pf_slot * 8);
          const uint32_t dst = wblk + (uint32_t)pf_slot * slot_bytes;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(slot_bytes) : "memory");
          if (p.a_rank == 2) tma_load_2d_u32(&p.map_res, bar, dst, pf_n0 + pf_ch * cw, pf_c1);
          else tma_load_4d_u32(&p.map_res, bar, dst, pf_n0 + pf_ch * cw, pf_c1, pf_c2, pf_c3);
        }
        if (++pf_slot == nslot) pf_slot = 0;
        pf_ch += 2;
        pf_settle();
      }
    };
    if (has_res && slab_ok) {
      pf_decode();
      pf_settle();
      for (int d = 0; d < pfd; ++d) issue_prefetch();
    }
    int slot = 0;          // fp32 slot of the chunk being processed
    uint32_t rphase = 0;   // parity of the residual barrier of `slot`
    int b16 = 0;           // 16-bit buffer of the chunk being processed
    int lt = 0;
    for (int tile = first_tile; tile < p.total_tiles; tile += tile_step, ++lt) {
      const TileCoord t = decode_tile<CG>(p, tile, cta_rank);
      const int acc = (p.n_acc == 2) ? 0 : (lt & 1);
      const int ch_first = (half + lt) & 1;
      const int c1 = t.w0 + sw0, c2 = t.h0 + sh0, c3 = t.nb0 + sn0;
      // Bias of every chunk this warp owns in the tile (chunks ch_first, ch_first + 2, ...: at most GEMM_BIAS_LINES),
      // requested BEFORE the accumulator wait: five independent loads whose latency hides behind the main loop - a
      // load per chunk sat on the epilogue's per-chunk latency chain (the bound of every short-K GEMM). The 32 column
      // values of a chunk reach the row-threads through a 128-byte shared-memory line (broadcast reads).
      float rowb = 0.f;
      if (p.bias_mode == 1) {
        float cb[GEMM_BIAS_LINES];
#pragma unroll
        for (int k = 0; k < GEMM_BIAS_LINES; ++k) {
          // k-th 32-column chunk of this warp: ch_first + 2k, or (w64) half k & 1 of unit ch_first + 2 (k >> 1)
          const int ck = w64 ? 2 * (ch_first + 2 * (k >> 1)) + (k & 1) : ch_first + 2 * k;
          const int col = t.n0 + ck * 32 + lane;
          cb[k] = (ck < nchunks && col < p.N) ? __ldg(p.bias + col) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < GEMM_BIAS_LINES; ++k)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_line + (uint32_t)(k * 128 + lane * 4)), "f"(cb[k]) : "memory");
      } else {
        if (lt == 0) {
#pragma unroll
          for (int k = 0; k < GEMM_BIAS_LINES; ++k)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_line + (uint32_t)(k * 128 + lane * 4)), "f"(0.f) : "memory");
        }
        if (p.bias_mode == 2) {
          const long long m = (long long)c1 + lane;          // rank-2 outputs only (host-checked)
          if (m < p.m_total) rowb = __ldg(p.bias + m);
        }
      }
      __syncwarp();
      if (e == 0 && lane == 0) { trace_stamp(trc, lt, 5); trace_epi(trc, lt, 0); }
      mbar_wait(&tfull_bar[acc], (uint32_t)(lt / nbuf) & 1u, 3);
      if (e == 0 && lane == 0) { trace_stamp(trc, lt, 6); trace_epi(trc, lt, 1); }
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)(acc * p.acc_stride) + ((uint32_t)(q * 32) << 16);
      const int cpa = p.acc_n >> 5;
      // all tcgen05.ld of this accumulator have completed: hand it back to the MMA issuer (the leader's barrier;
      // a remote arrive from the peer CTA) - before the last chunk's arithmetic and stores
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster_relaxed(smem_u32(&tempty_bar[acc]) & PEER_BIT_MASK);
          else mbar_arrive_relaxed(&tempty_bar[acc]);
        }
      };
      auto valid = [&](int c) { return c < nchunks && t.n0 + c * 32 < p.N; };
      auto taddr = [&](int c) {
        return t_addr + (uint32_t)((p.n_acc == 2 && c >= cpa) ? 256 + (c - cpa) * 32 : c * 32);
      };
      // chunk prologue: free the buffers of the chunk before the previous one and request the next residual -
      // everything that does not need the accumulator
      bool tr16 = false;                           // trace mode 16: the chunk being processed is the stamped one
      const bool tr16_lane = (trc & 16) && e == 0 && lane == 0;
      // in w64 mode `ch` still counts 32-column chunks; unit = ch >> 1, half = ch & 1
      auto first_of_unit = [&](int ch) { return !w64 || (ch & 1) == 0; };
      auto last_of_unit = [&](int ch) { return !w64 || (ch & 1) == 1 || !valid(ch + 1); };
      auto next_chunk = [&](int ch) { return w64 ? ((ch & 1) ? ch + 3 : ch + 1) : ch + 2; };
      auto pre = [&](int ch) {
        // stores of the unit before the previous one have left shared memory: its slot / 16-bit buffer
        // are free again (and the slot the prefetch below targets is the one that unit used)
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        if (has_res) issue_prefetch();
      };
      // chunk body: accumulator row in v (loaded and waited for)
      auto post = [&](uint32_t (&v)[32], int ch, int ch_next, bool more) {
        const uint32_t sub4 = w64 ? (uint32_t)((ch & 1) * 4) : 0u;           // first 16-byte piece of this chunk in its row
        const int col0 = t.n0 + (w64 ? (ch & ~1) : ch) * 32;               // first column of the unit's store
        const int ord = w64 ? 2 * (((ch >> 1) - ch_first) >> 1) + (ch & 1) : (ch - ch_first) >> 1;
        const uint32_t bline = bias_line + (uint32_t)(ord * 128);
        const bool t0 = e == 0 && lane == 0 && ord == 0;
        if (t0) trace_epi(trc, lt, 2);
        const uint32_t srow = wblk + (uint32_t)slot * slot_bytes + (uint32_t)lane * (res16 ? row16 : 128u);
        if (has_res && first_of_unit(ch)) mbar_wait(res_bar + e * 4 + slot, rphase, 7);
        if (t0) trace_epi(trc, lt, 3);
        uint32_t pk[16];
        if (res16) {
          switch ((p.act != 0 ? 1 : 0) | (p.out_f16 ? 2 : 0)) {
            case 0: epi_rows16<false, false>(v, pk, srow, key16, sub4, bline, rowb, p.act); break;
            case 1: epi_rows16<true, false>(v, pk, srow, key16, sub4, bline, rowb, p.act); break;
            case 2: epi_rows16<false, true>(v, pk, srow, key16, sub4, bline, rowb, p.act); break;
            default: epi_rows16<true, true>(v, pk, srow, key16, sub4, bline, rowb, p.act); break;
          }
        } else {
          const uint32_t sx = (uint32_t)(lane & 7);
          switch ((p.act != 0 ? 1 : 0) | (has_res ? 2 : 0) | (f32out ? 4 : 0) | (p.out_f16 ? 8 : 0)) {
            case 0: epi_rows<false, false, false, false>(v, pk, srow, sx, bline, rowb, p.act); break;
            case 1: epi_rows<true, false, false, false>(v, pk, srow, sx, bline, rowb, p.act); break;
            case 2: epi_rows<false, true, false, false>(v, pk, srow, sx, bline, rowb, p.act); break;
            case 3: epi_rows<true, true, false, false>(v, pk, srow, sx, bline, rowb, p.act); break;
            case 4: epi_rows<false, false, true, false>(v, pk, srow, sx, bline, rowb, p.act); break;
            case 5: epi_rows<true, false, true, false>(v, pk, srow, sx, bline, rowb, p.act); break;
            case 6: epi_rows<false, true, true, false>(v, pk, srow, sx, bline, rowb, p.act); break;
            case 7: epi_rows<true, true, true, false>(v, pk, srow, sx, bline, rowb, p.act); break;
            case 8: epi_rows<false, false, false, true>(v, pk, srow, sx, bline, rowb, p.act); break;
            case 9: epi_rows<true, false, false, true>(v, pk, srow, sx, bline, rowb, p.act); break;
            case 10: epi_rows<false, true, false, true>(v, pk, srow, sx, bline, rowb, p.act); break;
            case 11: epi_rows<true, true, false, true>(v, pk, srow, sx, bline, rowb, p.act); break;
            case 12: epi_rows<false, false, true, true>(v, pk, srow, sx, bline, rowb, p.act); break;
            case 13: epi_rows<true, false, true, true>(v, pk, srow, sx, bline, rowb, p.act); break;
            case 14: epi_rows<false, true, true, true>(v, pk, srow, sx, bline, rowb, p.act); break;
            default: epi_rows<true, true, true, true>(v, pk, srow, sx, bline, rowb, p.act); break;
          }
        }
        // the row is in x / pk now: the next chunk's tcgen05.ld goes into the SAME registers and lands while this chunk
        // is staged, fenced and stored
        if (more) tmem_ld32(taddr(ch_next), v);
        if (t0) trace_epi(trc, lt, 6);
        if (tr16) trace_m16(trc, lt, 4);
        const uint32_t hbuf = b16_base + (uint32_t)b16 * b16_bytes;
        if (b16out) {
          const uint32_t hrow = hbuf + (uint32_t)lane * row16;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hrow + (((sub4 + (uint32_t)j) ^ key16) << 4)),
                         "r"(pk[4 * j]), "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                         : "memory");
        }
        if (tr16) trace_m16(trc, lt, 5);
        if (!last_of_unit(ch)) return;            // w64: the second half of the row follows, one store for both
        fence_proxy_async();
        __syncwarp();
        if (t0) trace_epi(trc, lt, 7);
        if (tr16) trace_m16(trc, lt, 6);
        if (p.gn_part != nullptr) {
          // GroupNorm statistics of the chunk (rank-2 outputs, fp32 result resident in the slot): lane = column,
          // 32 conflict-free reads down the swizzled rows
          const int row0 = c1;                                    // first row of the slab
          const int nrow = min(32, (int)p.m_total - row0);
          const uint32_t sbase = wblk + (uint32_t)(slot * GEMM_EPI_STAGE_BYTES) + (uint32_t)((lane & 3) * 4);
          const uint32_t cpiece = (uint32_t)(lane >> 2);
          float cs = 0.f, cq = 0.f;
          for (int r = 0; r < nrow; ++r) {
            float vv;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(vv)
                         : "r"(sbase + (uint32_t)(r * 128) + ((cpiece ^ (uint32_t)(r & 7)) << 4)));
            cs += vv;
            cq = fmaf(vv, vv, cq);
          }
          const int sn = row0 / p.gn_hw;
          const long long slab = (long long)sn * p.gn_K + (row0 - sn * p.gn_hw) / 32;
          if (col0 + lane < p.N)
            *reinterpret_cast<float2*>(p.gn_part + (slab * p.N + col0 + lane) * 2) = make_float2(cs, cq);
        }
        if (lane == 0) {
          const uint32_t fsrc = wblk + (uint32_t)(slot * GEMM_EPI_STAGE_BYTES);
          if (p.a_rank == 2) {
            if (f32out) tma_store_2d(&p.map_out, fsrc, col0, c1);
            if (b16out) tma_store_2d(map16, hbuf, col0, c1);
          } else {
            if (f32out) tma_store_4d(&p.map_out, fsrc, col0, c1, c2, c3);
            if (b16out) tma_store_4d(map16, hbuf, col0, c1, c2, c3);
          }
          bulk_commit();
        }
        if (t0) trace_epi(trc, lt, 4);
        if (tr16) trace_m16(trc, lt, 7);
        if (nslot > 0 && ++slot == nslot) { slot = 0; rphase ^= 1u; }
        b16 ^= 1;
      };
      const int ch0 = w64 ? 2 * ch_first : ch_first;
      if (slab_ok && valid(ch0)) {
        // ONE register buffer and one copy of the chunk body: the next chunk's tcgen05.ld is issued as soon as the
        // current rows have been consumed. (The first form kept two buffers and two unrolled copies of the body; ncu
        // on the 65536 x 320 x 320 projection: 27 % of the stall samples "no instruction", instruction-cache hit rate
        // 85 % - two warps per sub-partition walking ~25 KB of straight-line code. Half the code, 32 registers less.)
        uint32_t v[32];
        int ch = ch0;
        int nth = 0;
        tmem_ld32(taddr(ch), v);
        for (;;) {
          const bool stamp = tr16_lane && nth == 1;          // trace mode 16: this warp's second chunk of the tile
          if (first_of_unit(ch)) pre(ch);
          if (stamp) trace_m16(trc, lt, 1);
          tmem_ld_wait();
          if (stamp) trace_m16(trc, lt, 2);
          const int ch_next = next_chunk(ch);
          const bool more = valid(ch_next);
          if (!more) release_acc();                          // all tcgen05.ld of this accumulator have completed
          if (stamp) trace_m16(trc, lt, 3);
          tr16 = stamp;
          post(v, ch, ch_next, more);
          tr16 = false;
          if (!more) break;
          if (tr16_lane && nth == 0) trace_m16(trc, lt, 0);
          ch = ch_next;
          ++nth;
        }
      } else {
        release_acc();
      }
      if (e == 0 && lane == 0) { trace_epi(trc, lt, 5); trace_stamp(trc, lt, 7); }
    }
    if (lane == 0) bulk_wait<0>();       // every store of this warp has been performed before the CTA retires
    __syncwarp();
  } else {
    // ===== Epilogue warps. Warp e = warp - 2 owns TMEM lanes [32*(warp%4), +32) and the 32-column
    // chunks e/4, e/4 + 2, ...  Phase 1: thread = row, TMEM -> 128B-swizzled shared chunk.
    // Phase 2: 8 lanes per row, 16 bytes each, so every global access is a full 128-byte row segment.
    const int e = warp - 2;
    const int q = warp & 3;
    const int half = e >> 2;
    uint8_t* stg = epi_smem + e * GEMM_EPI_STAGE_BYTES;
    const uint32_t stg_addr = smem_u32(stg);
    const int nchunks = (p.block_n + 31) / 32;
    const int sub = lane >> 3;   // row within a group of 4
    const int c4 = lane & 7;     // 16-byte column group
    const int cl = c4 * 4;       // first of this lane's 4 columns in a chunk
    const bool has_res = (p.residual != nullptr) && (p.nsplit == 1);
    const bool res_async = has_res && p.res_async;
    const bool res_direct = has_res && p.res_direct;
    // warp-uniform: the vectorised epilogue applies (no split-K partials, residual via the ring or
    // absent, 16-byte aligned fp32 rows / 8-byte aligned bf16 rows)
    const bool fast_ok =
        (p.nsplit == 1) && (!has_res || res_async || res_direct) && (p.ldo % 4 == 0) &&
        (p.out_fp32 ? ((reinterpret_cast<uintptr_t>(p.out) & 15u) == 0 &&
                       (p.out2 == nullptr || (reinterpret_cast<uintptr_t>(p.out2) & 7u) == 0))
                    : ((reinterpret_cast<uintptr_t>(p.out) & 7u) == 0)) &&
        (p.bias_mode != 1 || (reinterpret_cast<uintptr_t>(p.bias) & 15u) == 0);
    const bool split_fast = (p.nsplit > 1) && (p.N % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.workspace) & 15u) == 0);
    const uint32_t ring_addr = smem_u32(epi_smem + GEMM_EPI_WARPS * GEMM_EPI_STAGE_BYTES +
                                        e * (GEMM_RES_RING * GEMM_EPI_STAGE_BYTES));
    // (dw, dh, dn) of the 8 tile rows this lane touches in phase 2 (local rows sub + 4*i): fixed for the
    // whole kernel, so the divisions happen once and not per tile
    int rdec[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = q * 32 + i * 4 + sub;
      const int dw = r % p.bw;
      const int dh = (r / p.bw) % p.bh;
      const int dn = r / (p.bw * p.bh);
      rdec[i] = dw | (dh << 8) | (dn << 16);
    }
    auto rows_for = [&](const TileCoord& t, int (&mr)[8]) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int dn = rdec[i] >> 16;
        const int ww = t.w0 + (rdec[i] & 0xff), hh = t.h0 + ((rdec[i] >> 8) & 0xff), nn = t.nb0 + dn;
        const bool ok = (dn < p.bn) && (ww < p.WO) && (hh < p.HO) && (nn < p.NB);
        // output row index, -1 = outside the tensor (up-sampling phases scatter into the 2x larger output)
        const int ua = p.up_all ? (t.z >> 1) : p.up_a, ub = p.up_all ? (t.z & 1) : p.up_b;
        mr[i] = !ok ? -1 : (p.up ? ((nn * p.HO + hh) * 2 + ua) * (2 * p.WO) + 2 * ww + ub
                                 : (nn * p.HO + hh) * p.WO + ww);
      }
    };
    // ---- residual prefetch cursor: walks the same (tile, chunk) sequence as the main loop,
    // GEMM_RES_RING - 1 chunks ahead, each chunk one cp.async group into ring slot seq % GEMM_RES_RING.
    // A lane later reads back exactly the 16-byte pieces it requested itself, so cp.async.wait_group is
    // the only synchronisation needed.
    int pf_tile = first_tile, pf_lt = 0, pf_ch = half, pf_slot = 0;
    int pf_mrow[8];
    TileCoord pf_t;
    if (res_async) {
      while (pf_tile < p.total_tiles && pf_ch >= nchunks) { pf_tile += tile_step; ++pf_lt; pf_ch = (half + pf_lt) & 1; }
      if (pf_tile < p.total_tiles) { pf_t = decode_tile<CG>(p, pf_tile, cta_rank); rows_for(pf_t, pf_mrow); }
    }
    auto issue_prefetch = [&]() {
      if (pf_tile < p.total_tiles) {
        const int c0n = pf_t.n0 + pf_ch * 32;
        const int ncn = min(32, min(p.block_n - pf_ch * 32, p.N - c0n));
        if (cl + 4 <= ncn) {
          const uint32_t dst0 = ring_addr + (uint32_t)(pf_slot * GEMM_EPI_STAGE_BYTES);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = pf_mrow[i];
            if (m < 0) continue;
            const int lr = i * 4 + sub;
            const float* src = reinterpret_cast<const float*>(p.residual) + (long long)m * p.ldr + c0n + cl;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (uint32_t)(lr * 128 + ((c4 ^ (lr & 7)) << 4))),
                         "l"(src)
                         : "memory");
          }
        }
        pf_ch += 2;
        if (pf_ch >= nchunks) {
          do { pf_tile += tile_step; ++pf_lt; pf_ch = (half + pf_lt) & 1; } while (pf_tile < p.total_tiles && pf_ch >= nchunks);
          if (pf_tile < p.total_tiles) { pf_t = decode_tile<CG>(p, pf_tile, cta_rank); rows_for(pf_t, pf_mrow); }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (++pf_slot == GEMM_RES_RING) pf_slot = 0;
    };
    if (res_async) {
#pragma unroll
      for (int dpf = 0; dpf < GEMM_RES_RING - 1; ++dpf) issue_prefetch();
    }
    int rd_slot = 0;       // ring slot of the chunk being consumed
    int lt = 0;
    for (int tile = first_tile; tile < p.total_tiles; tile += tile_step, ++lt) {
      const TileCoord t = decode_tile<CG>(p, tile, cta_rank);
      const int acc = (p.n_acc == 2) ? 0 : (lt & 1);
      const int ch_first = (half + lt) & 1;     // alternate so both halves get the odd chunk in turn
      int mrow[8];
      rows_for(t, mrow);
      // (sample, slab) of this warp's 32 rows for the GroupNorm partial sums
      bool gn_ok = false;
      long long gn_slab = 0;
      if (p.gn_part != nullptr) {
        const int r0 = q * 32;
        if (p.a_rank == 2) {
          const int row0 = t.w0 + r0;
          const int sn = row0 / p.gn_hw;
          gn_ok = row0 < p.WO;
          gn_slab = (long long)sn * p.gn_K + (row0 - sn * p.gn_hw) / 32;
        } else {
          const int plane = p.bw * p.bh;
          const int dn = r0 / plane;
          const int sn = t.nb0 + dn;
          gn_ok = (dn < p.bn) && (sn < p.NB);
          gn_slab = (long long)sn * p.gn_stride + (p.up_all ? t.z * p.gn_K : p.gn_off) + t.sp * p.gn_spq + (r0 - dn * plane) / 32;
        }
      }
      if (e == 0 && lane == 0) { trace_stamp(trc, lt, 5); trace_epi(trc, lt, 0); }
      mbar_wait(&tfull_bar[acc], (uint32_t)(lt / nbuf) & 1u, 3);
      if (e == 0 && lane == 0) { trace_stamp(trc, lt, 6); trace_epi(trc, lt, 1); }
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)(acc * p.acc_stride) + ((uint32_t)(q * 32) << 16);
      for (int ch = ch_first; ch < nchunks; ch += 2) {
        const int col0 = t.n0 + ch * 32;
        const int ncol = min(32, min(p.block_n - ch * 32, p.N - col0));
        const bool vec_ok = (cl + 4 <= ncol);
        // bias of this lane's 4 columns: requested before the accumulator load so that its latency
        // hides behind tcgen05.ld and phase 1
        float4 cbias = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias_mode == 1 && cl < ncol) {
          if (vec_ok && ((reinterpret_cast<uintptr_t>(p.bias + col0 + cl) & 15u) == 0)) {
            cbias = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + cl));
          } else {
            cbias.x = __ldg(p.bias + col0 + cl);
            if (cl + 1 < ncol) cbias.y = __ldg(p.bias + col0 + cl + 1);
            if (cl + 2 < ncol) cbias.z = __ldg(p.bias + col0 + cl + 2);
            if (cl + 3 < ncol) cbias.w = __ldg(p.bias + col0 + cl + 3);
          }
        }
        uint32_t v[32];
        // wide mode: chunks [0, acc_n/32) live at TMEM column 0, the rest at column 256
        const int cpa = p.acc_n >> 5;
        tmem_ld32(t_addr + (uint32_t)((p.n_acc == 2 && ch >= cpa) ? 256 + (ch - cpa) * 32 : ch * 32), v);
        if (res_async) issue_prefetch();                       // chunk seq + GEMM_RES_RING - 1
        // long-K tiles keep their shared memory for pipeline stages: the residual of this chunk is
        // requested here and lands while the accumulator is loaded and staged
        float4 rdir[8];
        if (res_direct && fast_ok && ncol == 32) {
          const float* rb0 = reinterpret_cast<const float*>(p.residual) + col0 + cl;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            rdir[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (mrow[i] >= 0) {
              const float* rp = rb0 + (long long)mrow[i] * p.ldr;
              asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(rdir[i].x), "=f"(rdir[i].y), "=f"(rdir[i].z), "=f"(rdir[i].w)
                           : "l"(rp));
            }
          }
        }
        tmem_ld_wait();
        if (e == 0 && lane == 0 && ch == ch_first) trace_epi(trc, lt, 2);
        // phase 1: row `lane` of the chunk -> staging, 16-byte pieces XOR-swizzled by the row
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t dst = stg_addr + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[4 * j]),
                       "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                       : "memory");
        }
        if (res_async) asm volatile("cp.async.wait_group %0;" ::"n"(GEMM_RES_RING - 1) : "memory");
        __syncwarp();
        if (e == 0 && lane == 0 && ch == ch_first) trace_epi(trc, lt, 3);
        // phase 2
        const uint32_t rslot = ring_addr + (uint32_t)(rd_slot * GEMM_EPI_STAGE_BYTES);
        if (split_fast && ncol == 32) {
          // split-K partial sums: raw fp32 accumulator rows into this slice of the workspace, 16 bytes per lane
          float* wb = p.workspace + (long long)t.z * p.m_total * p.N + col0 + cl;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int lr = i * 4 + sub;
            const uint32_t swz = (uint32_t)(lr * 128 + ((c4 ^ (lr & 7)) << 4));
            float4 x;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                         : "r"(stg_addr + swz));
            if (mrow[i] >= 0) *reinterpret_cast<float4*>(wb + (long long)mrow[i] * p.N) = x;
          }
        } else if (fast_ok && ncol == 32) {
          // fast path (full chunk, aligned pointers): all shared-memory reads first, then the math,
          // then the stores - eight independent chains per lane, no per-element guards
          float4 x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int lr = i * 4 + sub;
            const uint32_t swz = (uint32_t)(lr * 128 + ((c4 ^ (lr & 7)) << 4));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(x[i].x), "=f"(x[i].y), "=f"(x[i].z), "=f"(x[i].w)
                         : "r"(stg_addr + swz));
          }
          if (p.bias_mode == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float rb = (mrow[i] >= 0) ? __ldg(p.bias + mrow[i]) : 0.f;
              x[i].x += rb; x[i].y += rb; x[i].z += rb; x[i].w += rb;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              x[i].x += cbias.x; x[i].y += cbias.y; x[i].z += cbias.z; x[i].w += cbias.w;
            }
          }
          if (p.act != 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              x[i].x = apply_act(x[i].x, p.act); x[i].y = apply_act(x[i].y, p.act);
              x[i].z = apply_act(x[i].z, p.act); x[i].w = apply_act(x[i].w, p.act);
            }
          }
          if (res_async) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int lr = i * 4 + sub;
              const uint32_t swz = (uint32_t)(lr * 128 + ((c4 ^ (lr & 7)) << 4));
              float4 rr;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(rr.x), "=f"(rr.y), "=f"(rr.z), "=f"(rr.w)
                           : "r"(rslot + swz));
              x[i].x += rr.x; x[i].y += rr.y; x[i].z += rr.z; x[i].w += rr.w;
            }
          } else if (res_direct) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              x[i].x += rdir[i].x; x[i].y += rdir[i].y; x[i].z += rdir[i].z; x[i].w += rdir[i].w;
            }
          }
          if (p.out_fp32) {
            float* ob = reinterpret_cast<float*>(p.out) + col0 + cl;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (mrow[i] >= 0) *reinterpret_cast<float4*>(ob + (long long)mrow[i] * p.ldo) = x[i];
          }
          __nv_bfloat16* bp = p.out_fp32 ? p.out2 : reinterpret_cast<__nv_bfloat16*>(p.out);
          if (bp != nullptr) {
            __nv_bfloat16* ob = bp + col0 + cl;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (mrow[i] >= 0)
                *reinterpret_cast<uint2*>(ob + (long long)mrow[i] * p.ldo) =
                    make_uint2(pack16x2(x[i].x, x[i].y, p.out_f16), pack16x2(x[i].z, x[i].w, p.out_f16));
          }
          if (p.gn_part != nullptr) {
            // GroupNorm statistics of the values just written: this lane's 4 columns over its 8 rows, then
            // over the 4 row groups of the warp (lane bits 3-4): column sums of the 32-row slab
            float4 cs = make_float4(0.f, 0.f, 0.f, 0.f), cq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (mrow[i] >= 0) {
                cs.x += x[i].x; cs.y += x[i].y; cs.z += x[i].z; cs.w += x[i].w;
                cq.x = fmaf(x[i].x, x[i].x, cq.x); cq.y = fmaf(x[i].y, x[i].y, cq.y);
                cq.z = fmaf(x[i].z, x[i].z, cq.z); cq.w = fmaf(x[i].w, x[i].w, cq.w);
              }
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
              cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
              cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
              cq.x += __shfl_xor_sync(0xffffffffu, cq.x, o); cq.y += __shfl_xor_sync(0xffffffffu, cq.y, o);
              cq.z += __shfl_xor_sync(0xffffffffu, cq.z, o); cq.w += __shfl_xor_sync(0xffffffffu, cq.w, o);
            }
            if (sub == 0 && gn_ok) {
              float4* dst = reinterpret_cast<float4*>(p.gn_part + (gn_slab * p.N + col0 + cl) * 2);
              dst[0] = make_float4(cs.x, cq.x, cs.y, cq.y);
              dst[1] = make_float4(cs.z, cq.z, cs.w, cq.w);
            }
          }
        } else {
  #pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int lr = i * 4 + sub;
            const int m = mrow[i];
            if (m < 0 || cl >= ncol) continue;
            const uint32_t swz = (uint32_t)(lr * 128 + ((c4 ^ (lr & 7)) << 4));
            float4 x;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                         : "r"(stg_addr + swz));
            if (p.nsplit > 1) {
              float* wsp = p.workspace + ((long long)t.z * p.m_total + m) * p.N + col0 + cl;
              if (vec_ok && ((reinterpret_cast<uintptr_t>(wsp) & 15u) == 0)) {
                *reinterpret_cast<float4*>(wsp) = x;
              } else {
                wsp[0] = x.x;
                if (cl + 1 < ncol) wsp[1] = x.y;
                if (cl + 2 < ncol) wsp[2] = x.z;
                if (cl + 3 < ncol) wsp[3] = x.w;
              }
              continue;
            }
            const float rb = (p.bias_mode == 2) ? __ldg(p.bias + m) : 0.0f;   // per-row bias (swapped GEMMs)
            x.x += cbias.x + rb; x.y += cbias.y + rb;
            x.z += cbias.z + rb; x.w += cbias.w + rb;
            if (p.act != 0) {
              x.x = apply_act(x.x, p.act); x.y = apply_act(x.y, p.act);
              x.z = apply_act(x.z, p.act); x.w = apply_act(x.w, p.act);
            }
            if (res_async) {
              float4 rr;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(rr.x), "=f"(rr.y), "=f"(rr.z), "=f"(rr.w)
                           : "r"(rslot + swz));
              x.x += rr.x; x.y += rr.y; x.z += rr.z; x.w += rr.w;
            } else if (has_res) {
              // bf16 or unaligned fp32 residual: direct loads (not on the hot path of the engines)
              const long long roff = (long long)m * p.ldr + col0 + cl;
              if (p.res_fp32 == 1) {
                const float* rp = reinterpret_cast<const float*>(p.residual) + roff;
                x.x += rp[0];
                if (cl + 1 < ncol) x.y += rp[1];
                if (cl + 2 < ncol) x.z += rp[2];
                if (cl + 3 < ncol) x.w += rp[3];
              } else if (p.res_fp32 == 2) {
                const __half* rp = reinterpret_cast<const __half*>(p.residual) + roff;
                x.x += __half2float(rp[0]);
                if (cl + 1 < ncol) x.y += __half2float(rp[1]);
                if (cl + 2 < ncol) x.z += __half2float(rp[2]);
                if (cl + 3 < ncol) x.w += __half2float(rp[3]);
              } else {
                const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + roff;
                if (vec_ok && ((reinterpret_cast<uintptr_t>(rp) & 7u) == 0)) {
                  const uint2 u = __ldg(reinterpret_cast<const uint2*>(rp));
                  x.x += bf16lo(u.x); x.y += bf16hi(u.x); x.z += bf16lo(u.y); x.w += bf16hi(u.y);
                } else {
                  x.x += __bfloat162float(rp[0]);
                  if (cl + 1 < ncol) x.y += __bfloat162float(rp[1]);
                  if (cl + 2 < ncol) x.z += __bfloat162float(rp[2]);
                  if (cl + 3 < ncol) x.w += __bfloat162float(rp[3]);
                }
              }
            }
            const long long ooff = (long long)m * p.ldo + col0 + cl;
            if (p.out_fp32) {
              float* op = reinterpret_cast<float*>(p.out) + ooff;
              if (vec_ok && ((reinterpret_cast<uintptr_t>(op) & 15u) == 0)) {
                *reinterpret_cast<float4*>(op) = x;
              } else {
                op[0] = x.x;
                if (cl + 1 < ncol) op[1] = x.y;
                if (cl + 2 < ncol) op[2] = x.z;
                if (cl + 3 < ncol) op[3] = x.w;
              }
            }
            __nv_bfloat16* bp = p.out_fp32 ? p.out2 : reinterpret_cast<__nv_bfloat16*>(p.out);
            if (bp != nullptr) {
              __nv_bfloat16* op = bp + ooff;
              if (vec_ok && ((reinterpret_cast<uintptr_t>(op) & 7u) == 0)) {
                *reinterpret_cast<uint2*>(op) = make_uint2(pack16x2(x.x, x.y, p.out_f16), pack16x2(x.z, x.w, p.out_f16));
              } else {
                unsigned short* os = reinterpret_cast<unsigned short*>(op);
                os[0] = cvt16(x.x, p.out_f16);
                if (cl + 1 < ncol) os[1] = cvt16(x.y, p.out_f16);
                if (cl + 2 < ncol) os[2] = cvt16(x.z, p.out_f16);
                if (cl + 3 < ncol) os[3] = cvt16(x.w, p.out_f16);
              }
            }
          }
        }
        if (++rd_slot == GEMM_RES_RING) rd_slot = 0;
        __syncwarp();
        if (e == 0 && lane == 0 && ch == ch_first) trace_epi(trc, lt, 4);
      }
      if (e == 0 && lane == 0) trace_epi(trc, lt, 5);
      // all tcgen05.ld of this accumulator have completed (tmem_ld_wait above): hand it back to the
      // MMA issuer (the leader's barrier; a remote arrive from the peer CTA)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster_relaxed(smem_u32(&tempty_bar[acc]) & PEER_BIT_MASK);
        else mbar_arrive_relaxed(&tempty_bar[acc]);
        if (e == 0) trace_stamp(trc, lt, 7);
      }
    }
  }

  __syncwarp();      // the single-thread producer / issuer loops rejoin their warps here
  tc_fence_before();
  if constexpr (CG == 2) {
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// Split-K finalize, four columns per thread (N % 4 == 0, 16-byte aligned fp32 operands, per-column bias or none):
// the scalar form below spent its time in one 64-bit division per element.
__global__ void gemm_splitk_finalize4_kernel(const float* __restrict__ ws, int nsplit, long long m_total, int N,
                                             void* out, long long ldo, int out_fp32, const float* __restrict__ bias,
                                             const float* __restrict__ residual, long long ldr, int act,
                                             __nv_bfloat16* __restrict__ out2, int f16) {
  pdl_trigger();
  pdl_wait();
  const int nv = N >> 2;
  const long long totalv = m_total * nv;
  const long long total = m_total * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < totalv;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / nv;
    const int n = (int)(i - m * nv) * 4;
    float4 acc = *reinterpret_cast<const float4*>(ws + m * N + n);
    for (int z = 1; z < nsplit; ++z) {
      const float4 v = *reinterpret_cast<const float4*>(ws + (long long)z * total + m * N + n);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (bias != nullptr) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n));
      acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    }
    if (act != 0) {
      acc.x = apply_act(acc.x, act); acc.y = apply_act(acc.y, act);
      acc.z = apply_act(acc.z, act); acc.w = apply_act(acc.w, act);
    }
    if (residual != nullptr) {
      const float4 r = *reinterpret_cast<const float4*>(residual + m * ldr + n);
      acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
    }
    const uint2 h = make_uint2(pack16x2(acc.x, acc.y, f16), pack16x2(acc.z, acc.w, f16));
    if (out_fp32) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + m * ldo + n) = acc;
      if (out2) *reinterpret_cast<uint2*>(out2 + m * ldo + n) = h;
    } else {
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + m * ldo + n) = h;
    }
  }
}

// Split-K finalize: out = act(sum_z ws[z] + bias) + residual.
__global__ void gemm_splitk_finalize_kernel(const float* __restrict__ ws, int nsplit,
                                            long long m_total, int N, void* out, long long ldo,
                                            int out_fp32, const float* __restrict__ bias,
                                            int bias_mode, const void* __restrict__ residual,
                                            int res_fp32, long long ldr, int act,
                                            __nv_bfloat16* __restrict__ out2, int f16) {
  pdl_trigger();
  pdl_wait();
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
      acc += res_fp32 == 1 ? reinterpret_cast<const float*>(residual)[m * ldr + n]
           : res_fp32 == 2 ? __half2float(reinterpret_cast<const __half*>(residual)[m * ldr + n])
                           : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(residual)[m * ldr + n]);
    }
    if (out_fp32) {
      reinterpret_cast<float*>(out)[m * ldo + n] = acc;
      if (out2) reinterpret_cast<unsigned short*>(out2)[m * ldo + n] = cvt16(acc, f16);
    } else {
      reinterpret_cast<unsigned short*>(out)[m * ldo + n] = cvt16(acc, f16);
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

static int device_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

static int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

}  // namespace sdb

// Debug: switch the per-tile timeline of CTA 0 on/off, or (on < 0) copy GEMM_TRACE_TILES*8 stamps out.
extern "C" int sdb_debug_gemm_trace(int on, long long* out_host) {
  if (on >= 0) {
    long long zero[sdb::GEMM_TRACE_TILES * 8];
    memset(zero, 0, sizeof(zero));
    cudaMemcpyToSymbol(sdb::g_gemm_trace, zero, sizeof(zero));
    cudaMemcpyToSymbol(sdb::g_gemm_trace_on, &on, sizeof(int));
    return SDB_OK;
  }
  if (!out_host) return SDB_ERR_ARG;
  cudaError_t e = cudaMemcpyFromSymbol(out_host, sdb::g_gemm_trace, sizeof(long long) * sdb::GEMM_TRACE_TILES * 8);
  return e == cudaSuccess ? SDB_OK : SDB_ERR_CUDA;
}

// Slabs (32 output rows of one TMEM lane quadrant) per sample for the epilogue's GroupNorm partial sums, or 0
// when a slab could straddle two samples / the geometry is unsupported.
extern "C" int sdb_gemm_gn_slabs(int kind, int NB, int HI, int WI, int M, int gn_hw) {
  using namespace sdb;
  if (kind == SDB_GEMM_LINEAR) {
    if (gn_hw <= 0 || gn_hw % 32 != 0 || M <= 0 || M % gn_hw != 0) return 0;
    return gn_hw / 32;
  }
  if (NB <= 0 || HI <= 0 || WI <= 0) return 0;
  const bool s2 = (kind == SDB_GEMM_CONV3X3_S2 || kind == SDB_GEMM_CONV3X3_S2_PAD_RB);
  const int HO = s2 ? HI / 2 : HI, WO = s2 ? WI / 2 : WI;     // CONV2X2_UP: slabs of ONE phase (low-resolution grid)
  int bw = 0, bh = 0, bn = 0;
  if (pick_tile_box(NB, HO, WO, &bw, &bh, &bn)) return 0;
  const int plane = bw * bh;
  if (plane % 32 != 0) return 0;
  const int spq = plane / 32 < 4 ? plane / 32 : 4;
  return ((WO + bw - 1) / bw) * ((HO + bh - 1) / bh) * spq;
}

// Bytes of A one pipeline stage fetches per THREE k-blocks when a stride-1 3x3 conv over [NB, HI, WI] takes the
// filter-column staging (one halo box per filter column), or 0 when its tile geometry rules that out (the classic
// form fetches 3 x 16 KiB). For the tile chooser's TMA model.
extern "C" int sdb_gemm_conv_a3_bytes(int NB, int HI, int WI) {
  using namespace sdb;
  int bw = 0, bh = 0, bn = 0;
  if (NB <= 0 || HI <= 0 || WI <= 0 || pick_tile_box(NB, HI, WI, &bw, &bh, &bn)) return 0;
  const char* ev = getenv("SDB_NO_A3");
  if (ev && ev[0] == '1') return 0;
  if (bn != 1 || bw * bh != GEMM_BM || bw % 8 != 0) return 0;
  return (bh + 2) * bw * 128;      // (an up-sampling phase fetches (bh + 1) * bw * 128 per two k-blocks: same per k-block)
}

extern "C" int sdb_gemm_tc(const sdb_gemm_args* a, void* stream_) {
  using namespace sdb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->a0 || !a->w || !a->out) { set_error("sdb_gemm_tc: null pointer"); return SDB_ERR_ARG; }
  if (a->Cout <= 0 || a->C0 <= 0) { set_error("sdb_gemm_tc: bad channel counts"); return SDB_ERR_ARG; }
  const int kind = a->kind;
  if (kind < SDB_GEMM_LINEAR || kind > SDB_GEMM_CONV2X2_UP) {
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
    const bool s2 = (kind == SDB_GEMM_CONV3X3_S2 || kind == SDB_GEMM_CONV3X3_S2_PAD_RB);
    const bool up = (kind == SDB_GEMM_CONV2X2_UP);
    if (s2 && (a->C1 > 0 || (HI & 1) || (WI & 1))) {
      set_error("sdb_gemm_tc: stride-2 conv needs even H, W and a single source");
      return SDB_ERR_UNSUPPORTED;
    }
    if (up && (a->up_phase < 0 || a->up_phase > 4 || a->nsplit > 1 || a->residual != nullptr)) {
      set_error("sdb_gemm_tc: CONV2X2_UP needs up_phase in 0..3 (or 4: all phases, w = [4][Cout][4 C]), no split-K and "
                "no residual");
      return SDB_ERR_UNSUPPORTED;
    }
    if (a->C0 % 64 != 0) { set_error("sdb_gemm_tc: conv needs C0 %% 64 == 0"); return SDB_ERR_UNSUPPORTED; }
    p.NB = NB; p.HO = s2 ? HI / 2 : HI; p.WO = s2 ? WI / 2 : WI;
    p.ntaps = up ? 4 : 9;
    if (pick_tile_box(p.NB, p.HO, p.WO, &p.bw, &p.bh, &p.bn)) { set_error("tile box"); return SDB_ERR_ARG; }
    if (up) {
      p.up = 1; p.up_all = (a->up_phase == 4) ? 1 : 0;
      p.up_a = p.up_all ? 0 : a->up_phase >> 1; p.up_b = p.up_all ? 0 : a->up_phase & 1;
      for (int u = 0; u < 2; ++u)
        for (int v = 0; v < 2; ++v) {
          const int t = u * 2 + v;
          p.tap_dc_sel[t] = 0; p.tap_d2[t] = 0;
          p.tap_dh[t] = (int8_t)(p.up_a - 1 + u); p.tap_dw[t] = (int8_t)(p.up_b - 1 + v);
        }
    } else
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
  if (a->ax0 != nullptr) {
    if (kind != SDB_GEMM_CONV3X3_S1 || a->Cx0 <= 0 || a->Cx0 % 64 != 0 || a->Cx1 % 64 != 0 || (a->Cx1 > 0 && !a->ax1) ||
        a->C0 % 64 != 0 || a->C1 % 64 != 0) {
      set_error("sdb_gemm_tc: the extra 1x1 source needs a stride-1 3x3 conv and channel counts that are multiples of 64");
      return SDB_ERR_UNSUPPORTED;
    }
    p.cblocks_x0 = a->Cx0 / 64;
    p.cblocks_x = p.cblocks_x0 + a->Cx1 / 64;
    p.ktot += a->Cx0 + a->Cx1;
    uint32_t box[5] = {64, (uint32_t)p.bw, 1, (uint32_t)p.bh, (uint32_t)p.bn};
    const uint64_t WI = (uint64_t)a->WI, HI = (uint64_t)a->HI;
    const uint64_t c0b = (uint64_t)a->Cx0 * 2;
    uint64_t dims[5] = {(uint64_t)a->Cx0, WI, 1, HI, (uint64_t)a->NB};
    uint64_t str[4] = {c0b, c0b * WI, c0b * WI, c0b * WI * HI};
    if ((rc = make_tmap_bf16(&p.map_x0, a->ax0, 5, dims, str, box, "conv extra source 0"))) return rc;
    if (a->Cx1 > 0) {
      const uint64_t c1b = (uint64_t)a->Cx1 * 2;
      uint64_t dims1[5] = {(uint64_t)a->Cx1, WI, 1, HI, (uint64_t)a->NB};
      uint64_t str1[4] = {c1b, c1b * WI, c1b * WI, c1b * WI * HI};
      if ((rc = make_tmap_bf16(&p.map_x1, a->ax1, 5, dims1, str1, box, "conv extra source 1"))) return rc;
    }
  }
  p.nkb_total = p.ntaps * p.cblocks + p.cblocks_x;

  // ---- N tiling
  int block_n = a->block_n;
  if (block_n <= 0) {
    if (a->Cout % 256 == 0) block_n = 256;
    else if (a->Cout % 160 == 0) block_n = 160;
    else if (a->Cout % 128 == 0) block_n = 128;
    else if (a->Cout >= 256) block_n = 256;
    else block_n = ((a->Cout + 15) / 16) * 16;
  }
  if ((block_n % 16 != 0 || block_n < 16 || block_n > 256) && block_n != 320) {
    set_error("sdb_gemm_tc: block_n %d invalid", block_n);
    return SDB_ERR_ARG;
  }
  p.block_n = block_n;
  p.n_acc = (block_n == 320) ? 2 : 1;                     // wide tiles: two accumulators of 160 columns
  p.acc_n = block_n / p.n_acc;
  p.acc_stride = p.n_acc == 2 ? 512 : pow2_cols(((block_n + 31) / 32) * 32);
  p.tmem_cols = p.n_acc == 2 ? 512 : 2 * p.acc_stride;
  const int n_tiles = (a->Cout + block_n - 1) / block_n;
  // CTA pairs (cta_group::2, 256-row tiles) whenever there are at least two row tiles
  int cg = (m_tiles >= 2) ? 2 : 1;
  if (a->cta_pair == 1) cg = 1;
  if (a->cta_pair == 2) cg = 2;
  {
    uint64_t dims[2] = {(uint64_t)p.ktot, (uint64_t)a->Cout * (p.up_all ? 4u : 1u)};     // up_all: the four phases' rows
    const long long ldw = a->ldw ? a->ldw : p.ktot;
    uint64_t str[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {64, (uint32_t)(p.acc_n / cg)};
    if ((rc = make_tmap_bf16(&p.map_w, a->w, 2, dims, str, box, "gemm W"))) return rc;
  }

  // ---- filter-column staging (a3): stride-1 3x3 conv, tile = bw x bh patch of one sample with bw % 8 == 0 (tap views
  // stay on 1024-byte swizzle atoms), no split-K (a split could cut a filter column), at least one whole channel block
  const int b_stage_host = p.n_acc * (p.acc_n / cg) * GEMM_BK * 2;
  {
    static int no_a3 = -1;
    if (no_a3 < 0) { const char* ev = getenv("SDB_NO_A3"); no_a3 = (ev && ev[0] == '1') ? 1 : 0; }
    const int fr = p.up ? 2 : 3;                 // filter rows = filter columns
    const bool ok = !no_a3 && (kind == SDB_GEMM_CONV3X3_S1 || p.up) && p.bn == 1 && p.bw * p.bh == GEMM_BM && p.bw % 8 == 0 &&
                    a->nsplit <= 1 && ((p.acc_n / cg) % 8 == 0) && a->C0 % 64 == 0 && a->C1 % 64 == 0 &&
                    // long reductions only (main-loop bound; short ones keep the shared memory for the TMA epilogue),
                    // and two stages must fit next to the per-lane epilogue's staging
                    p.nkb_total > 32 &&
                    2 * std::max((p.bh + fr - 1) * p.bw * 128 + fr * b_stage_host, 2 * (GEMM_A_STAGE_BYTES + b_stage_host)) +
                            GEMM_EPI_WARPS * GEMM_EPI_STAGE_BYTES + 1024 + GEMM_BAR_BYTES <= 227 * 1024 &&
                    (a->smem_budget <= 0 || a->smem_budget >= 227 * 1024);
    if (ok) {
      p.a3 = 1;
      p.a3_nrow = fr; p.a3_ncol = fr;
      p.a3_h0 = p.up ? p.up_a - 1 : -1;
      p.a3_w0 = p.up ? p.up_b - 1 : -1;
      p.a3_box_bytes = (p.bh + fr - 1) * p.bw * 128;
      p.a3_iters = fr * p.cblocks + (p.cblocks_x + 1) / 2;
      p.a3_stage_bytes = p.a3_box_bytes + fr * b_stage_host;
      if (p.cblocks_x > 0 && 2 * (GEMM_A_STAGE_BYTES + b_stage_host) > p.a3_stage_bytes)
        p.a3_stage_bytes = 2 * (GEMM_A_STAGE_BYTES + b_stage_host);
      // the main sources are fetched with the taller box; the extra 1x1 source keeps its one-tap boxes
      uint32_t box3[5] = {64, (uint32_t)p.bw, 1, (uint32_t)(p.bh + fr - 1), 1};
      const uint64_t WI = (uint64_t)a->WI, HI = (uint64_t)a->HI;
      const uint64_t c0b = (uint64_t)a->C0 * 2;
      uint64_t dims[5] = {(uint64_t)a->C0, WI, 1, HI, (uint64_t)a->NB};
      uint64_t str[4] = {c0b, c0b * WI, c0b * WI, c0b * WI * HI};
      if ((rc = make_tmap_bf16(&p.map_a0, a->a0, 5, dims, str, box3, "conv A0 (filter-column box)"))) return rc;
      if (a->C1 > 0) {
        const uint64_t c1b = (uint64_t)a->C1 * 2;
        uint64_t dims1[5] = {(uint64_t)a->C1, WI, 1, HI, (uint64_t)a->NB};
        uint64_t str1[4] = {c1b, c1b * WI, c1b * WI, c1b * WI * HI};
        if ((rc = make_tmap_bf16(&p.map_a1, a->a1, 5, dims1, str1, box3, "conv A1 (filter-column box)"))) return rc;
      }
    }
  }

  // ---- pipeline depth from the shared-memory budget
  const int stage_bytes = p.a3 ? p.a3_stage_bytes : (GEMM_A_STAGE_BYTES + b_stage_host);
  const long long ldr_eff = a->ldr ? a->ldr : a->Cout;
  const long long ldo_eff = a->ldo ? a->ldo : a->Cout;
  const int want_split = a->nsplit > 1;
  const int nkb_all = p.nkb_total;
  const int smem_budget = (a->smem_budget > 0 && a->smem_budget < 227 * 1024) ? a->smem_budget : 227 * 1024;
  int fixed_bytes = 0;
  int stages = 0;
  // Short reductions are epilogue-bound: TMA epilogue (bulk tensor stores, residual by TMA loads) when the
  // 32-row slabs of a tile are boxes of the output tensor and every pointer / stride meets the TMA alignment.
  {
    static int no_epi_tma = -1;
    if (no_epi_tma < 0) { const char* ev = getenv("SDB_NO_EPI_TMA"); no_epi_tma = (ev && ev[0] == '1') ? 1 : 0; }
    const bool f32o = a->out_fp32 != 0;
    const bool has16 = !f32o || a->out2 != nullptr;
    const int epi_mode = a->epi_mode;            // 0 auto, 1 never, 2 whenever eligible (also long reductions)
    bool ok = !no_epi_tma && epi_mode != 1 && !want_split && (nkb_all <= 32 || epi_mode == 2);
    if (a->gn_part != nullptr && p.a_rank == 5) ok = false;      // conv outputs: statistics in the per-lane epilogue
    if (p.up) ok = false;                                        // scattered output rows: per-lane epilogue
    ok = ok && (block_n % 32 == 0 || n_tiles == 1);
    ok = ok && (reinterpret_cast<uintptr_t>(a->out) & 15u) == 0 && (f32o ? (ldo_eff % 4 == 0) : (ldo_eff % 8 == 0));
    if (a->out2) ok = ok && (reinterpret_cast<uintptr_t>(a->out2) & 15u) == 0 && (ldo_eff % 8 == 0);
    const bool res16 = a->residual != nullptr && a->res_fp32 == 2;
    if (a->residual)
      ok = ok && (reinterpret_cast<uintptr_t>(a->residual) & 15u) == 0 &&
           (res16 ? (!f32o && ldr_eff % 8 == 0) : (a->res_fp32 == 1 && ldr_eff % 4 == 0));
    // 16-bit result only and 64-column tile boundaries: pairs of chunks per store (epi_w64)
    static int no_w64 = -1;
    if (no_w64 < 0) { const char* ev = getenv("SDB_NO_EPI_W64"); no_w64 = (ev && ev[0] == '1') ? 1 : 0; }
    const bool w64 = !no_w64 && !f32o && a->out2 == nullptr && (a->residual == nullptr || res16) && block_n % 64 == 0 &&
                     a->gn_part == nullptr;
    const int b16_bytes = GEMM_EPI16_BYTES * (w64 ? 2 : 1);
    const int slot_bytes = res16 ? b16_bytes : GEMM_EPI_STAGE_BYTES;
    if (a->bias)
      ok = ok && (a->bias_per_row ? (kind == SDB_GEMM_LINEAR) : ((reinterpret_cast<uintptr_t>(a->bias) & 15u) == 0));
    const int rows = p.bw * p.bh * p.bn, plane = p.bw * p.bh;
    if (p.a_rank == 5)
      ok = ok && ((p.bw % 32 == 0) || (32 % p.bw == 0)) && ((plane % 32 == 0) || (32 % plane == 0)) && (rows % 32 == 0);
    if (ok) {
      // residual slots per warp: fp32 (in place: ahead | current | draining) 4, else 3; half (read only: ahead |
      // current) 3, else 2 - the smaller count when the pipeline would otherwise be left with fewer than 3 stages
      int nslot = a->residual ? (res16 ? 3 : 4) : (f32o ? 2 : 0);
      const int nslot_min = a->residual ? (res16 ? 2 : 3) : nslot;
      for (;;) {
        p.epi_warp_bytes = nslot * slot_bytes + (has16 ? 2 * b16_bytes : 0);
        fixed_bytes = GEMM_EPI_WARPS * (p.epi_warp_bytes + GEMM_BIAS_LINES * 128) + 1024 + GEMM_BAR_BYTES;   // + the bias lines of a warp
        stages = (smem_budget - fixed_bytes) / stage_bytes;
        if (stages >= 3 || nslot == nslot_min) break;
        --nslot;
      }
      if (stages >= 2) {
        p.epi_tma = 1;
        p.epi_nslot = nslot;
        p.epi_w64 = w64 ? 1 : 0;
        uint32_t box[4] = {w64 ? 64u : 32u, 32, 1, 1};
        for (int q = 0; q < 4; ++q) {
          const int r0 = q * 32;
          if (p.a_rank == 2) { p.slab_w0[q] = (int8_t)r0; p.slab_ok[q] = 1; }
          else {
            p.slab_ok[q] = (int8_t)(r0 < rows);
            p.slab_w0[q] = (int8_t)(r0 % p.bw);
            p.slab_h0[q] = (int8_t)((r0 / p.bw) % p.bh);
            p.slab_n0[q] = (int8_t)(r0 / plane);
          }
        }
        if (p.a_rank == 5) {
          const int sbw = p.bw < 32 ? p.bw : 32;
          const int sbh = p.bw >= 32 ? 1 : (p.bh < 32 / p.bw ? p.bh : 32 / p.bw);
          box[1] = (uint32_t)sbw; box[2] = (uint32_t)sbh; box[3] = (uint32_t)(32 / (sbw * sbh));
        }
        auto mk = [&](CUtensorMap* m, const void* base, int esz, long long ld, const char* what) -> int {
          if (p.a_rank == 2) {
            uint64_t dims[2] = {(uint64_t)a->Cout, (uint64_t)a->M};
            uint64_t str[1] = {(uint64_t)ld * esz};
            return make_tmap(m, base, esz, (esz == 4 || w64) ? 128 : 64, 2, dims, str, box, what);
          }
          uint64_t dims[4] = {(uint64_t)a->Cout, (uint64_t)p.WO, (uint64_t)p.HO, (uint64_t)p.NB};
          uint64_t str[3] = {(uint64_t)ld * esz, (uint64_t)ld * esz * p.WO, (uint64_t)ld * esz * p.WO * p.HO};
          return make_tmap(m, base, esz, (esz == 4 || w64) ? 128 : 64, 4, dims, str, box, what);
        };
        if ((rc = mk(&p.map_out, a->out, f32o ? 4 : 2, ldo_eff, "gemm out"))) return rc;
        if (a->out2 && (rc = mk(&p.map_out2, a->out2, 2, ldo_eff, "gemm out2"))) return rc;
        if (a->residual && (rc = mk(&p.map_res, a->residual, res16 ? 2 : 4, ldr_eff, "gemm residual"))) return rc;
      }
    }
  }
  if (!p.epi_tma) {
    // fp32 residuals stream through a cp.async ring (3 chunks per epilogue warp) when 16-byte aligned
    const bool res_vec = (a->residual != nullptr && a->res_fp32 == 1 && !want_split && (ldr_eff % 4) == 0 &&
                          (a->Cout % 4) == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15u) == 0);
    // short reductions are epilogue-bound (ring: deep prefetch); long ones keep the shared memory for stages
    p.res_async = (res_vec && nkb_all <= 32) ? 1 : 0;
    p.res_direct = (res_vec && !p.res_async) ? 1 : 0;
    fixed_bytes = GEMM_EPI_WARPS * GEMM_EPI_STAGE_BYTES * (1 + (p.res_async ? GEMM_RES_RING : 0)) + 1024 + GEMM_BAR_BYTES;
    stages = (smem_budget - fixed_bytes) / stage_bytes;
    if (stages < 2 && p.res_async) {          // the residual ring does not fit next to two stages: direct loads
      p.res_async = 0;
      p.res_direct = 1;
      fixed_bytes = GEMM_EPI_WARPS * GEMM_EPI_STAGE_BYTES + 1024 + GEMM_BAR_BYTES;
      stages = (smem_budget - fixed_bytes) / stage_bytes;
    }
  }
  p.epi_bytes = fixed_bytes - 1024 - GEMM_BAR_BYTES;
  if (stages > GEMM_MAX_STAGES) stages = GEMM_MAX_STAGES;
  if (stages < 2) { set_error("sdb_gemm_tc: shared-memory budget too small"); return SDB_ERR_ARG; }
  const int nkb_total = p.nkb_total;
  int nsplit = a->nsplit > 0 ? a->nsplit : 1;
  if (nsplit > nkb_total) nsplit = nkb_total;
  p.per_split = (nkb_total + nsplit - 1) / nsplit;
  nsplit = (nkb_total + p.per_split - 1) / p.per_split;   // no empty slice
  if (nsplit > 1 && !a->workspace) { set_error("sdb_gemm_tc: split-K needs a workspace"); return SDB_ERR_ARG; }
  p.nsplit = nsplit;
  p.stages = stages;
  const int smem_bytes = stages * stage_bytes + fixed_bytes;

  p.out = a->out;
  p.ldo = a->ldo ? a->ldo : a->Cout;
  p.out_fp32 = a->out_fp32;
  p.out_f16 = a->out_f16;
  p.ab_f16 = a->ab_f16;
  // out_f16: every store path (TMA epilogue, both per-lane paths, split-K finalize) writes the 16-bit tensor as IEEE
  // half; a bf16 residual is the one combination without a path
  if (a->residual != nullptr && (a->res_fp32 < 0 || a->res_fp32 > 2)) {
    set_error("sdb_gemm_tc: res_fp32 must be 0 (bf16), 1 (fp32) or 2 (IEEE half)");
    return SDB_ERR_ARG;
  }
  if (a->out_f16 && a->residual != nullptr && !a->res_fp32) {
    set_error("sdb_gemm_tc: out_f16 with a bf16 residual is not supported");
    return SDB_ERR_UNSUPPORTED;
  }
  p.bias = a->bias;
  p.bias_mode = a->bias ? (a->bias_per_row ? 2 : 1) : 0;
  p.residual = a->residual;
  p.res_fp32 = a->res_fp32;
  p.out2 = reinterpret_cast<__nv_bfloat16*>(a->out2);
  if (a->out2 && !a->out_fp32) { set_error("sdb_gemm_tc: out2 (bf16 copy) needs out_fp32"); return SDB_ERR_ARG; }
  p.ldr = a->ldr ? a->ldr : a->Cout;
  p.act = a->act;
  p.workspace = a->workspace;
  if (a->gn_part != nullptr) {
    const int K = sdb_gemm_gn_slabs(kind, a->NB, a->HI, a->WI, a->M, a->gn_hw);
    const bool aligned = (ldo_eff % 4 == 0) && (reinterpret_cast<uintptr_t>(a->out) & 15u) == 0 &&
                         (!a->out2 || (reinterpret_cast<uintptr_t>(a->out2) & 7u) == 0) &&
                         (!a->bias || a->bias_per_row || (reinterpret_cast<uintptr_t>(a->bias) & 15u) == 0) &&
                         (!a->residual || (a->res_fp32 == 1 && (ldr_eff % 4) == 0 &&
                                           (reinterpret_cast<uintptr_t>(a->residual) & 15u) == 0)) &&
                         (reinterpret_cast<uintptr_t>(a->gn_part) & 15u) == 0;
    if (K <= 0 || (!a->out_fp32 && p.epi_tma) || nsplit != 1 || a->Cout % 32 != 0 || block_n % 32 != 0 || !aligned) {
      set_error("sdb_gemm_tc: gn_part needs an fp32 output, no split-K, Cout %% 32 == 0, 16-byte aligned operands and "
                "32-row slabs inside one sample (K=%d nsplit=%d block_n=%d)", K, nsplit, block_n);
      return SDB_ERR_UNSUPPORTED;
    }
    p.gn_part = a->gn_part;
    p.gn_K = K;
    p.gn_stride = p.up ? 4 * K : K;
    p.gn_off = (p.up && !p.up_all) ? a->up_phase * K : 0;
    p.gn_hw = a->gn_hw;
    const int plane = p.bw * p.bh;
    p.gn_spq = (p.a_rank == 5) ? (plane / 32 < 4 ? plane / 32 : 4) : 0;
  }

  {
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           227 * 1024);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SDB_ERR_CUDA; }
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
  }
  const long long row_tiles = (m_tiles + cg - 1) / cg;        // 256-row tiles in pair mode
  const long long total_tiles = row_tiles * n_tiles * (p.up_all ? 4 : nsplit);
  if (total_tiles > 2147483647LL) { set_error("sdb_gemm_tc: too many tiles"); return SDB_ERR_UNSUPPORTED; }
  p.m_tiles = (int)row_tiles;
  p.n_tiles = n_tiles;
  p.total_tiles = (int)total_tiles;
  {
    const int divs[4] = {p.n_tiles, p.m_tiles, p.tiles_w, p.tiles_h};
    for (int i = 0; i < 4; ++i) {
      const uint32_t d = (uint32_t)divs[i];
      if (d <= 1) { p.fd_mul[i] = 0; p.fd_shr[i] = 0; continue; }
      uint32_t lg = 0;
      while ((1u << lg) < d) ++lg;                       // ceil(log2 d)
      const uint32_t sh = 31 + lg;
      p.fd_mul[i] = (uint32_t)(((1ull << sh) + d - 1) / d);
      p.fd_shr[i] = sh - 32;
    }
  }
  const int num_sms = device_sm_count();
  const long long slots = num_sms / cg;                       // CTAs (or CTA pairs) resident at once
  const unsigned grid = (unsigned)((total_tiles < slots ? total_tiles : slots) * cg);
  {
    cudaError_t e = (cg == 2) ? launch_k(gemm_tc_kernel<2>, dim3(grid), dim3(GEMM_THREADS), (size_t)smem_bytes, stream, 2, p)
                              : launch_k(gemm_tc_kernel<1>, dim3(grid), dim3(GEMM_THREADS), (size_t)smem_bytes, stream, 1, p);
    if (e != cudaSuccess) { set_error("gemm_tc_kernel launch: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return SDB_ERR_CUDA; }
  }
  if ((rc = check_launch("gemm_tc_kernel"))) return rc;
  if (nsplit > 1) {
    const long long total = p.m_total * p.N;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    const bool vec4 = (p.N % 4 == 0) && (p.ldo % 4 == 0) && p.bias_mode != 2 &&
                      (p.residual == nullptr || (p.res_fp32 == 1 && p.ldr % 4 == 0 &&
                                                 (reinterpret_cast<uintptr_t>(p.residual) & 15u) == 0)) &&
                      (reinterpret_cast<uintptr_t>(p.workspace) & 15u) == 0 &&
                      (reinterpret_cast<uintptr_t>(p.out) & 15u) == 0 &&
                      (p.out2 == nullptr || (reinterpret_cast<uintptr_t>(p.out2) & 7u) == 0) &&
                      (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15u) == 0);
    if (vec4) {
      int blocks4 = (int)((total / 4 + 255) / 256);
      if (blocks4 > 148 * 8) blocks4 = 148 * 8;
      (void)launch_k(gemm_splitk_finalize4_kernel, dim3(blocks4), dim3(256), 0, stream, 1,
                     (const float*)p.workspace, nsplit, p.m_total, p.N, p.out, p.ldo, p.out_fp32, p.bias,
                     reinterpret_cast<const float*>(p.residual), p.ldr, p.act, p.out2, p.out_f16);
      if ((rc = check_launch("gemm_splitk_finalize4_kernel"))) return rc;
      return SDB_OK;
    }
    (void)launch_k(gemm_splitk_finalize_kernel, dim3(blocks), dim3(256), 0, stream, 1,
        (const float*)p.workspace, nsplit, p.m_total, p.N, p.out, p.ldo, p.out_fp32, p.bias, p.bias_mode,
        p.residual, p.res_fp32, p.ldr, p.act, p.out2, p.out_f16);
    if ((rc = check_launch("gemm_splitk_finalize_kernel"))) return rc;
  }
  return SDB_OK;
}
