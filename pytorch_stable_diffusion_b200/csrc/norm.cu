// HBM-bound normalisation kernels: GroupNorm (statistics + apply[+SiLU]), LayerNorm, row softmax.
// All reductions are fp32 per thread and fp64 across threads/blocks; activations are bf16 NHWC.
#include <cuda_fp16.h>
#include "common.cuh"
#include "host.h"
#include "../../include/sdb200.h"

namespace sdb {

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f, int f16 = 0) {
  uint4 u;
  u.x = pack16x2(f[0], f[1], f16); u.y = pack16x2(f[2], f[3], f16);
  u.z = pack16x2(f[4], f[5], f16); u.w = pack16x2(f[6], f[7], f16);
  return u;
}

// 8 consecutive channels of one pixel from a bf16 (is_fp32 = 0), fp32 (1) or IEEE-half (2) tensor.
__device__ __forceinline__ void load8(const void* base, long long elem_off, int is_fp32, float* f) {
  if (is_fp32 == 1) {
    const float* p = reinterpret_cast<const float*>(base) + elem_off;
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off));
    if (is_fp32 == 2) {            // input kind 2: IEEE half
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        f[2 * i] = v.x; f[2 * i + 1] = v.y;
      }
    } else {
      unpack8(u, f);
    }
  }
}

// Thread layout shared by stats and apply: V = (C0+C1)/8 8-channel vectors per pixel; a block holds
// `lanes` pixels side by side, thread t -> (vector t % V, pixel lane t / V). Each thread keeps the
// same 8 channels for its whole life, so per-channel state lives in registers.
// grid = (chunks, NB); every block walks pixels [chunk*ppc, (chunk+1)*ppc), four pixels per thread in
// flight. Statistics are deterministic: every block writes its per-group partial sums to
// stats[n][chunk][g][2] (no atomics, nothing to zero) and the apply kernel adds the chunks in order.
constexpr int GN_MAX_CHUNKS = 128;

__global__ void gn_stats_kernel(const void* __restrict__ x0, const void* __restrict__ x1,
                                double* __restrict__ stats, long long HW, int C0, int C1, int groups,
                                long long ppc, int V, int lanes, int x0_fp32, int x1_fp32) {
  extern __shared__ float s_part[];     // [lanes][ctot] sums, then [lanes][ctot] sums of squares
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.y;
  const int t = threadIdx.x;
  const int v = t % V;
  const int pl = t / V;
  const int V0 = C0 >> 3;
  const int ctot = C0 + C1;
  const int cpg = ctot / groups;
  float* s_sum = s_part;
  float* s_sq = s_part + lanes * ctot;
  if (pl < lanes) {
    const bool second = v >= V0;
    const void* src = second ? x1 : x0;
    const int src_fp32 = second ? x1_fp32 : x0_fp32;
    const long long base = second ? (long long)n * HW * C1 + (long long)(v - V0) * 8
                                  : (long long)n * HW * C0 + (long long)v * 8;
    const long long stride = second ? C1 : C0;
    float sum[8], sq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sum[j] = 0.f; sq[j] = 0.f; }
    const long long p_begin = (long long)blockIdx.x * ppc;
    const long long p_end = min(HW, p_begin + ppc);
    long long p = p_begin + pl;
    for (; p + 3LL * lanes < p_end; p += 4LL * lanes) {
      float f[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) load8(src, base + (p + (long long)u * lanes) * stride, src_fp32, f[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) { sum[j] += f[u][j]; sq[j] += f[u][j] * f[u][j]; }
    }
    for (; p < p_end; p += lanes) {
      float f[8];
      load8(src, base + p * stride, src_fp32, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { sum[j] += f[j]; sq[j] += f[j] * f[j]; }
    }
    float* ds = s_sum + pl * ctot + v * 8;
    float* dq = s_sq + pl * ctot + v * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) { ds[j] = sum[j]; dq[j] = sq[j]; }
  }
  __syncthreads();
  for (int g = t; g < groups; g += blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int l = 0; l < lanes; ++l) {
      const float* ps = s_sum + l * ctot + g * cpg;
      const float* pq = s_sq + l * ctot + g * cpg;
      for (int c = 0; c < cpg; ++c) { a += (double)ps[c]; b += (double)pq[c]; }
    }
    double* dst = stats + (((long long)n * GN_MAX_CHUNKS + blockIdx.x) * groups + g) * 2;
    dst[0] = a;
    dst[1] = b;
  }
}

// raw (unconverted) 8-channel vector of one pixel: one 16-byte word for a 16-bit source, two for fp32
template <bool ALL16> struct GnRaw { uint4 a; uint4 b; };
template <> struct GnRaw<true> { uint4 a; };

template <bool ALL16>
__device__ __forceinline__ void gn_raw_load(GnRaw<ALL16>& r, const char* base, unsigned elem_off, int kind) {
  if constexpr (ALL16) {
    r.a = __ldg(reinterpret_cast<const uint4*>(base + (size_t)elem_off * 2u));
  } else {
    if (kind == 1) {
      const uint4* q = reinterpret_cast<const uint4*>(base + (size_t)elem_off * 4u);
      r.a = __ldg(q);
      r.b = __ldg(q + 1);
    } else {
      r.a = __ldg(reinterpret_cast<const uint4*>(base + (size_t)elem_off * 2u));
    }
  }
}

template <bool ALL16>
__device__ __forceinline__ void gn_raw_unpack(const GnRaw<ALL16>& r, int kind, float* f) {
  if constexpr (!ALL16) {
    if (kind == 1) {
      f[0] = __uint_as_float(r.a.x); f[1] = __uint_as_float(r.a.y); f[2] = __uint_as_float(r.a.z);
      f[3] = __uint_as_float(r.a.w); f[4] = __uint_as_float(r.b.x); f[5] = __uint_as_float(r.b.y);
      f[6] = __uint_as_float(r.b.z); f[7] = __uint_as_float(r.b.w);
      return;
    }
  }
  if (kind == 2) {
    const uint32_t w[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = v.x; f[2 * i + 1] = v.y;
    }
  } else {
    unpack8(r.a, f);
  }
}

// The apply pass is a pure stream (read 2-4 bytes, write 2 bytes per element), so what decides its speed is how many
// bytes each SM keeps in flight. A thread issues the raw 16-byte loads of U pixels back to back (U = 8 for 16-bit
// sources, 4 for fp32: 32 registers of payload either way), converts afterwards, and the first batch is issued BEFORE
// the statistics prologue (two block-wide syncs and fp64 arithmetic) so that its latency hides there; <= 68 registers
// keep ~960 threads per SM resident. (The first form - load8 straight into fp32 registers, 92 registers, two blocks
// per SM, one-pixel tail loop - ran the 16 x 64 x 64 x 320 tensor at 2.4 TB/s: long-scoreboard stalls 8.3 of 14
// cycles per issue, 23 % of the warp slots occupied; ncu in profiles/r02r_ncu_full_summary.txt.)
template <bool ALL16>
__global__ void __launch_bounds__(320, 3)
gn_apply_kernel(const void* __restrict__ x0, const void* __restrict__ x1,
                const double* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                long long HW, int C0, int C1, int groups, float eps, int silu,
                long long ppc, int V, int lanes, int x0_fp32, int x1_fp32,
                int stat_chunks, int out_f16) {
  constexpr int U = ALL16 ? 8 : 4;
  __shared__ float s_mean[64];
  __shared__ float s_rstd[64];
  extern __shared__ double s_stat[];     // [stat_chunks][groups][2] partial statistics of this sample
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.y;
  const int t = threadIdx.x;
  const int ctot = C0 + C1;
  const int cpg = ctot / groups;
  const int v = t % V;
  const int pl = t / V;
  const int V0 = C0 >> 3;
  const int c0 = v * 8;
  const bool second = v >= V0;
  const int kind = second ? x1_fp32 : x0_fp32;
  const int stride = second ? C1 : C0;
  // everything below is 32-bit arithmetic relative to this thread's first element of the sample (the host checks
  // HW * channels < 2^31): pointer + pixel * stride
  const char* src;
  {
    const long long e = second ? (long long)n * HW * C1 + (long long)(v - V0) * 8 : (long long)n * HW * C0 + (long long)v * 8;
    src = reinterpret_cast<const char*>(second ? x1 : x0) + e * (kind == 1 ? 4 : 2);
  }
  const int p_begin = (int)(blockIdx.x * ppc);
  const int p_end = (int)min(HW, (long long)p_begin + ppc);
  const int pstep = U * lanes;
  int p = p_begin + pl;
  GnRaw<ALL16> raw[U];
  if (pl < lanes) {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (p + u * lanes < p_end) gn_raw_load<ALL16>(raw[u], src, (unsigned)(p + u * lanes) * (unsigned)stride, kind);
  }
  // all threads fetch the per-chunk partials in parallel (one L2 round trip), then one thread per group
  // adds them in chunk order: the result does not depend on scheduling
  {
    const double* sp = stats + (long long)n * GN_MAX_CHUNKS * groups * 2;
    const int total = stat_chunks * groups * 2;
    for (int i = t; i < total; i += blockDim.x) s_stat[i] = sp[i];
  }
  __syncthreads();
  for (int g = t; g < groups; g += blockDim.x) {
    double s = 0.0, ss = 0.0;
    for (int c = 0; c < stat_chunks; ++c) {
      s += s_stat[(c * groups + g) * 2];
      ss += s_stat[(c * groups + g) * 2 + 1];
    }
    const double cnt = (double)HW * cpg;
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[g] = (float)mean;
    s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  if (pl >= lanes) return;
  float sc[8], sh[8];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c0 + j) / cpg;
      sc[j] = s_rstd[g] * ga[j];
      sh[j] = be[j] - s_mean[g] * s_rstd[g] * ga[j];
    }
  }
  char* obase = reinterpret_cast<char*>(out) + ((long long)n * HW * ctot + c0) * 2;
  const unsigned ostride = (unsigned)ctot * 2u;
  while (true) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = p + u * lanes;
      if (pp < p_end) {
        float f[8];
        gn_raw_unpack<ALL16>(raw[u], kind, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float y = f[j] * sc[j] + sh[j];
          if (silu) y = silu_f(y);
          f[j] = y;
        }
        *reinterpret_cast<uint4*>(obase + (size_t)((unsigned)pp * ostride)) = pack8(f, out_f16);
      }
    }
    p += pstep;
    if (p >= p_end) break;
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (p + u * lanes < p_end) gn_raw_load<ALL16>(raw[u], src, (unsigned)(p + u * lanes) * (unsigned)stride, kind);
  }
}

// One warp per row; rows of up to 5*32*8 = 1280 channels are held in registers between the passes.
constexpr int LN_MAX_VEC_PER_LANE = 5;
__global__ void layernorm_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, void* __restrict__ out,
                                 long long rows, int C, float eps, int in_fp32, int out_fp32) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int V = C >> 3;
  float f[LN_MAX_VEC_PER_LANE][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC_PER_LANE; ++i) {
    const int v = lane + i * 32;
    if (v < V) {
      load8(x, row * C + v * 8, in_fp32, f[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += f[i][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC_PER_LANE; ++i) {
    const int v = lane + i * 32;
    if (v < V) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = f[i][j] - mean; sq += d * d; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC_PER_LANE; ++i) {
    const int v = lane + i * 32;
    if (v < V) {
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        y[j] = (f[i][j] - mean) * rstd * __ldg(gamma + v * 8 + j) + __ldg(beta + v * 8 + j);
      if (out_fp32 == 1) {
        float* o = reinterpret_cast<float*>(out) + row * C + v * 8;
        *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(y[4], y[5], y[6], y[7]);
      } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + row * C + v * 8;
        *reinterpret_cast<uint4*>(o) = pack8(y, out_fp32 == 2);
      }
    }
  }
}

// fp32 rows, float4 loads, two rows per warp iteration (more bytes in flight per SM), grid-stride over
// row pairs. VPL = float4 vectors per lane = ceil(C / 128).
template <int VPL>
__global__ void __launch_bounds__(256)
layernorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, void* __restrict__ out, long long rows, int C,
                     float eps, int out_fp32) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int V = C >> 2;
  const float invC = 1.0f / (float)C;
  for (long long pair = warp0; pair * 2 < rows; pair += nwarps) {
    float4 f[2][VPL];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const long long row = pair * 2 + r;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + i * 32;
        f[r][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v < V && row < rows) {
          const float* src = x + row * C + v * 4;
          asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(f[r][i].x), "=f"(f[r][i].y), "=f"(f[r][i].z), "=f"(f[r][i].w)
                       : "l"(src));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const long long row = pair * 2 + r;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) sum += (f[r][i].x + f[r][i].y) + (f[r][i].z + f[r][i].w);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum * invC;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        if (lane + i * 32 < V) {
          const float a = f[r][i].x - mean, b = f[r][i].y - mean, c = f[r][i].z - mean, d = f[r][i].w - mean;
          sq += (a * a + b * b) + (c * c + d * d);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = rsqrtf(sq * invC + eps);
      if (row >= rows) continue;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + i * 32;
        if (v < V) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + v);
          const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + v);
          float4 y;
          y.x = (f[r][i].x - mean) * rstd * g.x + b.x;
          y.y = (f[r][i].y - mean) * rstd * g.y + b.y;
          y.z = (f[r][i].z - mean) * rstd * g.z + b.z;
          y.w = (f[r][i].w - mean) * rstd * g.w + b.w;
          if (out_fp32 == 1) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + row * C + v * 4) = y;
          } else {
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + row * C + v * 4) =
                make_uint2(pack16x2(y.x, y.y, out_fp32 == 2), pack16x2(y.z, y.w, out_fp32 == 2));
          }
        }
      }
    }
  }
}

// IEEE-half rows (the token stream inside an attention block, engine.TOK_F16): 8-byte loads with the column mapping of
// the fp32 kernel (VPL = ceil(C / 128) vectors of 4 per lane), R rows per warp iteration kept as RAW words and
// converted on the fly - R = 4 keeps as many bytes in flight as the fp32 kernel's two rows.
template <int VPL, int R>
__global__ void __launch_bounds__(256)
layernorm_h16_kernel(const __half* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, void* __restrict__ out, long long rows, int C,
                     float eps, int out_fp32) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int V = C >> 2;
  const float invC = 1.0f / (float)C;
  auto unpack4 = [](const uint2& u, float* f) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  };
  for (long long grp = warp0; grp * R < rows; grp += nwarps) {
    uint2 raw[R][VPL];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = grp * R + r;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + i * 32;
        raw[r][i] = make_uint2(0u, 0u);
        if (v < V && row < rows) {
          const __half* src = x + row * C + v * 4;
          asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];"
                       : "=r"(raw[r][i].x), "=r"(raw[r][i].y)
                       : "l"(src));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = grp * R + r;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        float f[4];
        unpack4(raw[r][i], f);
        sum += (f[0] + f[1]) + (f[2] + f[3]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum * invC;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        if (lane + i * 32 < V) {
          float f[4];
          unpack4(raw[r][i], f);
          const float a = f[0] - mean, b = f[1] - mean, c = f[2] - mean, d = f[3] - mean;
          sq += (a * a + b * b) + (c * c + d * d);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = rsqrtf(sq * invC + eps);
      if (row >= rows) continue;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int v = lane + i * 32;
        if (v < V) {
          float f[4];
          unpack4(raw[r][i], f);
          const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + v);
          const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + v);
          float4 y;
          y.x = (f[0] - mean) * rstd * g.x + b.x;
          y.y = (f[1] - mean) * rstd * g.y + b.y;
          y.z = (f[2] - mean) * rstd * g.z + b.z;
          y.w = (f[3] - mean) * rstd * g.w + b.w;
          if (out_fp32 == 1) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + row * C + v * 4) = y;
          } else {
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + row * C + v * 4) =
                make_uint2(pack16x2(y.x, y.y, out_fp32 == 2), pack16x2(y.z, y.w, out_fp32 == 2));
          }
        }
      }
    }
  }
}

// One block per row, the row staged in shared memory (cols * 4 bytes).
__global__ void softmax_rows_kernel(const float* __restrict__ scores, __nv_bfloat16* __restrict__ probs,
                                    int cols, float scale_log2) {
  extern __shared__ float s_row[];
  __shared__ float s_red[32];
  const long long row = blockIdx.x;
  const float* src = scores + row * cols;
  const int t = threadIdx.x, nt = blockDim.x;
  float mx = -INFINITY;
  for (int i = t; i < cols; i += nt) {
    const float v = src[i] * scale_log2;
    s_row[i] = v;
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((t & 31) == 0) s_red[t >> 5] = mx;
  __syncthreads();
  if (t < 32) {
    float v = (t < (nt >> 5)) ? s_red[t] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (t == 0) s_red[0] = v;
  }
  __syncthreads();
  mx = s_red[0];
  __syncthreads();
  float sum = 0.f;
  for (int i = t; i < cols; i += nt) {
    const float e = exp2f(s_row[i] - mx);
    s_row[i] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((t & 31) == 0) s_red[t >> 5] = sum;
  __syncthreads();
  if (t < 32) {
    float v = (t < (nt >> 5)) ? s_red[t] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (t == 0) s_red[0] = v;
  }
  __syncthreads();
  const float inv = 1.0f / s_red[0];
  __nv_bfloat16* dst = probs + row * cols;
  for (int i = t; i < cols; i += nt) dst[i] = __float2bfloat16_rn(s_row[i] * inv);
}

static int gn_geometry(int NB, long long HW, int ctot, int* V, int* lanes, int* threads,
                       long long* ppc, int* chunks, int blocks_per_sm = 4) {
  if (ctot % 8 != 0) return -1;
  *V = ctot / 8;
  if (*V > 1024) return -1;
  *lanes = 256 / *V;
  if (*lanes < 1) *lanes = 1;
  *threads = *V * *lanes;
  // aim for ~4 resident blocks per SM across the batch
  long long want = (148LL * blocks_per_sm + NB - 1) / NB;
  long long max_chunks = (HW + *lanes * 4 - 1) / (*lanes * 4);
  if (want > max_chunks) want = max_chunks;
  if (want > GN_MAX_CHUNKS) want = GN_MAX_CHUNKS;
  if (want < 1) want = 1;
  *ppc = (HW + want - 1) / want;
  *chunks = (int)((HW + *ppc - 1) / *ppc);
  return 0;
}


// Per-group statistics from the per-slab, per-channel partial sums a producing GEMM epilogue left behind
// (sdb_gemm_args::gn_part): one block per (group, sample), fixed-order fp64 sums -> the same stats layout
// gn_apply_kernel reads, as a single "chunk". part0: [NB][K0][C0][2], part1: [NB][K1][C1][2] (channel concat).
__global__ void __launch_bounds__(256) gn_reduce_partials_kernel(const float* __restrict__ part0,
                                                                 const float* __restrict__ part1,
                                                                 double* __restrict__ stats, int K0, int K1, int C0,
                                                                 int C1, int groups) {
  pdl_trigger();
  pdl_wait();
  // The (slab, channel) pairs of this group are one flat index space walked by all 256 threads, four independent
  // loads in flight per thread: neighbouring threads read neighbouring channels of one slab row (8 B each), and a
  // group of 4 channels x 2048 slabs (the VAE at 512 x 512) costs 8 L2 round trips per thread, not 64 on 32 threads
  // (the first layout - 32 channel lanes x 8 slab lanes - left 7/8 of the block idle there: 281 us per launch).
  const int g = blockIdx.x, n = blockIdx.y;
  const int t = threadIdx.x;
  const int cx = t & 31, ky = t >> 5;
  const int cpg = (C0 + C1) / groups;
  const int cbeg = g * cpg, cend = cbeg + cpg;
  // channels of this group in source 0: [cbeg, min(cend, C0)); in source 1: [max(cbeg, C0) - C0, cend - C0)
  const int n0 = max(0, min(cend, C0) - cbeg), n1 = cpg - n0;
  double ds = 0.0, dq = 0.0;
  auto walk = [&](const float* base, int nch, int K, int C) {
    // base -> [K][C][2] of this sample, already offset to the group's first channel in this source
    const int tot = nch * K;
    auto at = [&](int i) {
      const int k = i / nch, cc = i - k * nch;
      return *reinterpret_cast<const float2*>(base + ((long long)k * C + cc) * 2);
    };
    int i = t;
    for (; i + 768 < tot; i += 1024) {
      const float2 v0 = at(i), v1 = at(i + 256), v2 = at(i + 512), v3 = at(i + 768);
      ds += ((double)v0.x + (double)v1.x) + ((double)v2.x + (double)v3.x);
      dq += ((double)v0.y + (double)v1.y) + ((double)v2.y + (double)v3.y);
    }
    for (; i < tot; i += 256) {
      const float2 v = at(i);
      ds += (double)v.x; dq += (double)v.y;
    }
  };
  if (n0 > 0) walk(part0 + ((long long)n * K0 * C0 + cbeg) * 2, n0, K0, C0);
  if (n1 > 0) walk(part1 + ((long long)n * K1 * C1 + (max(cbeg, C0) - C0)) * 2, n1, K1, C1);
  __shared__ double sh[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
    dq += __shfl_xor_sync(0xffffffffu, dq, o);
  }
  if (cx == 0) { sh[0][ky] = ds; sh[1][ky] = dq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double* dst = stats + ((long long)n * GN_MAX_CHUNKS * groups + g) * 2;
    dst[0] = ((sh[0][0] + sh[0][1]) + (sh[0][2] + sh[0][3])) + ((sh[0][4] + sh[0][5]) + (sh[0][6] + sh[0][7]));
    dst[1] = ((sh[1][0] + sh[1][1]) + (sh[1][2] + sh[1][3])) + ((sh[1][4] + sh[1][5]) + (sh[1][6] + sh[1][7]));
  }
}

// ------------------------------------------------------------------------------------------------
// One-pass GroupNorm (+SiLU): the tensor is read from HBM exactly once.
//
// A thread-block cluster of CS CTAs owns (sample n, slab of G adjacent groups = SC channels); CTA r of the
// cluster keeps pixels [r*PX, (r+1)*PX) x SC channels of the fp32 input resident in shared memory (cp.async,
// the whole slice in flight at once), reduces per-group {sum, sum of squares} in a fixed order (fp32 per lane,
// fp64 across lanes / warps / CTAs: bit-reproducible), exchanges the CTA partials through distributed shared
// memory, and normalises its slice out of shared memory into the bf16 output. Two 82 KB CTAs share an SM, so
// one CTA's load phase overlaps the other's store phase.
struct GnFusedParams {
  const float* x0;
  const float* x1;
  const float* gamma;
  const float* beta;
  __nv_bfloat16* out;
  int HW, C0, C1, cpg, G, SC, CS, PX, nslabs, silu;
  float eps;
  int out_f16;
};

constexpr int GNF_THREADS = 256;
constexpr int GNF_WARPS = GNF_THREADS / 32;
constexpr int GNF_MAX_SC = 128;

__global__ void __launch_bounds__(GNF_THREADS) gn_fused_kernel(const GnFusedParams p) {
  extern __shared__ __align__(16) uint8_t gnf_smem[];
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (p.CS > 1) ? (int)cluster_ctarank() : 0;
  const int cid = blockIdx.x / p.CS;
  const int n = cid / p.nslabs;
  const int cb = (cid - n * p.nslabs) * p.SC;            // first channel of the slab
  const int px0 = rank * p.PX;
  const int npx = max(0, min(p.PX, p.HW - px0));
  const int C = p.C0 + p.C1;
  const int SC = p.SC;
  float* data = reinterpret_cast<float*>(gnf_smem);                                   // [PX][SC]
  float* part = data + (size_t)p.PX * SC;                                            // [warps][SC][2]
  float* coef = part + GNF_WARPS * SC * 2;                                           // [SC][2]
  double* cta_stats = reinterpret_cast<double*>(coef + 2 * SC);                       // [G][2] (16-byte aligned)

  // ---- pass A: slice -> shared memory, 16 bytes per cp.async, everything in flight at once
  {
    const int V = SC >> 2;                               // 16-byte vectors per pixel
    const int da = GNF_THREADS / V, db = GNF_THREADS - da * V;
    int px = tid / V, v = tid - px * V;
    const uint32_t dbase = smem_u32(data);
    const long long row0 = (long long)n * p.HW + px0;
    while (px < npx) {
      const int c = cb + 4 * v;
      const float* src = (c < p.C0) ? p.x0 + (row0 + px) * p.C0 + c : p.x1 + (row0 + px) * p.C1 + (c - p.C0);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dbase + (uint32_t)((px * SC + 4 * v) * 4)), "l"(src)
                   : "memory");
      v += db; px += da;
      if (v >= V) { v -= V; ++px; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();

  // ---- pass B: per-channel partial sums: lane -> floats lane, lane + 32, ... of a pixel row (bank = lane),
  // warp -> pixels warp, warp + 8, ...
  {
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    for (int px = warp; px < npx; px += GNF_WARPS) {
      const float* row = data + px * SC;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = lane + 32 * k;
        if (c < SC) { const float x = row[c]; s[k] += x; q[k] = fmaf(x, x, q[k]); }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = lane + 32 * k;
      if (c < SC) { part[(warp * SC + c) * 2] = s[k]; part[(warp * SC + c) * 2 + 1] = q[k]; }
    }
  }
  __syncthreads();
  // one warp per group: cpg x warps partials summed in fp64, fixed order
  for (int g = warp; g < p.G; g += GNF_WARPS) {
    double ds = 0.0, dq = 0.0;
    const int cnt = p.cpg * GNF_WARPS;
    for (int i = lane; i < cnt; i += 32) {
      const int w = i / p.cpg, c = g * p.cpg + (i - w * p.cpg);
      ds += (double)part[(w * SC + c) * 2];
      dq += (double)part[(w * SC + c) * 2 + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ds += __shfl_xor_sync(0xffffffffu, ds, o);
      dq += __shfl_xor_sync(0xffffffffu, dq, o);
    }
    if (lane == 0) { cta_stats[2 * g] = ds; cta_stats[2 * g + 1] = dq; }
  }
  __syncthreads();
  if (p.CS > 1) cluster_sync_all();                     // every CTA's partials are published

  // ---- pass C: cluster-wide statistics through distributed shared memory, per-channel scale / shift
  if (tid < SC) {
    const int g = tid / p.cpg;
    double ds = 0.0, dq = 0.0;
    if (p.CS > 1) {
      const uint32_t local = smem_u32(cta_stats + 2 * g);
      for (int r = 0; r < p.CS; ++r) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
        double a, b;
        asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(remote));
        ds += a; dq += b;
      }
    } else {
      ds = cta_stats[2 * g]; dq = cta_stats[2 * g + 1];
    }
    const double cnt = (double)p.cpg * (double)p.HW;
    const double mean = ds / cnt;
    double var = dq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)p.eps));
    const float a = __ldg(p.gamma + cb + tid) * rstd;
    coef[2 * tid] = a;
    coef[2 * tid + 1] = __ldg(p.beta + cb + tid) - (float)mean * a;
  }
  __syncthreads();
  // peers may still be reading this CTA's partials: arrive now, wait before leaving
  if (p.CS > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");

  // ---- pass D: normalise out of shared memory, 8 channels (16 bytes of bf16) per thread and step
  {
    const int VO = SC >> 3;
    const int da = GNF_THREADS / VO, db = GNF_THREADS - da * VO;
    int px = tid / VO, v = tid - px * VO;
    const long long row0 = (long long)n * p.HW + px0;
    while (px < npx) {
      const float4 x0 = *reinterpret_cast<const float4*>(data + px * SC + 8 * v);
      const float4 x1 = *reinterpret_cast<const float4*>(data + px * SC + 8 * v + 4);
      const float4 c0 = *reinterpret_cast<const float4*>(coef + 16 * v);
      const float4 c1 = *reinterpret_cast<const float4*>(coef + 16 * v + 4);
      const float4 c2 = *reinterpret_cast<const float4*>(coef + 16 * v + 8);
      const float4 c3 = *reinterpret_cast<const float4*>(coef + 16 * v + 12);
      float y[8] = {fmaf(x0.x, c0.x, c0.y), fmaf(x0.y, c0.z, c0.w), fmaf(x0.z, c1.x, c1.y), fmaf(x0.w, c1.z, c1.w),
                    fmaf(x1.x, c2.x, c2.y), fmaf(x1.y, c2.z, c2.w), fmaf(x1.z, c3.x, c3.y), fmaf(x1.w, c3.z, c3.w)};
      if (p.silu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = silu_f(y[i]);
      }
      uint4 o;
      o.x = pack16x2(y[0], y[1], p.out_f16); o.y = pack16x2(y[2], y[3], p.out_f16);
      o.z = pack16x2(y[4], y[5], p.out_f16); o.w = pack16x2(y[6], y[7], p.out_f16);
      *reinterpret_cast<uint4*>(p.out + (row0 + px) * C + cb + 8 * v) = o;
      v += db; px += da;
      if (v >= VO) { v -= VO; ++px; }
    }
  }
  if (p.CS > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// (G, CS, PX) for the one-pass kernel, or false when no slab of the tensor fits the shared memory of a cluster.
static bool gn_fused_plan(long long HW, int C0, int C1, int groups, int* G_, int* CS_, int* PX_, size_t* smem_) {
  const int C = C0 + C1;
  if (groups <= 0 || C % groups != 0 || C0 % 4 != 0 || C1 % 4 != 0 || HW <= 0 || HW > (1 << 24)) return false;
  const int cpg = C / groups;
  bool found = false;
  size_t best = 0;
  for (int G = 1; G <= groups; G <<= 1) {
    if (groups % G != 0) break;
    const int SC = G * cpg;
    if (SC % 8 != 0 || SC > GNF_MAX_SC) continue;
    if (SC < 40 && G * 2 <= groups && (G * 2 * cpg) <= GNF_MAX_SC) continue;     // 160-byte runs at least, if possible
    for (int CS = 1; CS <= 8; CS <<= 1) {
      const long long PX = (HW + CS - 1) / CS;
      const size_t bytes = (size_t)PX * SC * 4 + (size_t)GNF_WARPS * SC * 8 + (size_t)SC * 8 + (size_t)G * 16 + 16;
      if (bytes > 200 * 1024) continue;
      // first plan at or below 84 KB (two CTAs per SM); otherwise the smallest footprint seen
      if (!found || (best > 84 * 1024 && bytes < best)) {
        found = true; best = bytes; *G_ = G; *CS_ = CS; *PX_ = (int)PX; *smem_ = bytes;
      }
      if (bytes <= 84 * 1024) return true;
    }
    if (found && best <= 84 * 1024) return true;
  }
  return found;
}

}  // namespace sdb

extern "C" long long sdb_groupnorm_stats_bytes(int NB, int groups) {
  return (long long)NB * sdb::GN_MAX_CHUNKS * groups * 2 * (long long)sizeof(double);
}

extern "C" int sdb_groupnorm_stats(const void* x0, const void* x1, double* stats, int NB,
                                   long long HW, int C0, int C1, int groups, int x0_fp32, int x1_fp32,
                                   void* stream) {
  using namespace sdb;
  const int ctot = C0 + C1;
  if (!x0 || !stats || NB <= 0 || HW <= 0 || groups <= 0 || groups > 64 || ctot % groups != 0 ||
      C0 % 8 != 0 || C1 % 8 != 0 || (C1 > 0 && !x1)) {
    set_error("sdb_groupnorm_stats: bad arguments (C0=%d C1=%d groups=%d)", C0, C1, groups);
    return SDB_ERR_ARG;
  }
  int V, lanes, threads, chunks; long long ppc;
  if (gn_geometry(NB, HW, ctot, &V, &lanes, &threads, &ppc, &chunks)) {
    set_error("sdb_groupnorm_stats: unsupported channel count %d", ctot);
    return SDB_ERR_UNSUPPORTED;
  }
  const size_t smem = (size_t)2 * lanes * ctot * sizeof(float);
  if (smem > 48 * 1024) {
    set_error("sdb_groupnorm_stats: %d channels need %zu bytes of shared memory", ctot, smem);
    return SDB_ERR_UNSUPPORTED;
  }
  cudaError_t le = launch_k(gn_stats_kernel, dim3(chunks, NB), dim3(threads), smem, (cudaStream_t)stream, 1,
                            x0, x1, stats, HW, C0, C1, groups, ppc, V, lanes, x0_fp32, x1_fp32);
  if (le != cudaSuccess) { set_error("gn_stats_kernel launch: %s", cudaGetErrorString(le)); (void)cudaGetLastError(); return SDB_ERR_CUDA; }
  return check_launch("gn_stats_kernel");
}

extern "C" int sdb_groupnorm_reduce_partials(const float* part0, const float* part1, double* stats, int NB, int K0,
                                            int K1, int C0, int C1, int groups, void* stream) {
  using namespace sdb;
  if (!part0 || !stats || NB <= 0 || K0 <= 0 || groups <= 0 || groups > 64 || (C0 + C1) % groups != 0 ||
      (C1 > 0 && (!part1 || K1 <= 0))) {
    set_error("sdb_groupnorm_reduce_partials: bad arguments");
    return SDB_ERR_ARG;
  }
  cudaError_t le = launch_k(gn_reduce_partials_kernel, dim3(groups, NB), dim3(256), 0, (cudaStream_t)stream, 1,
                            part0, part1, stats, K0, K1, C0, C1, groups);
  if (le != cudaSuccess) { set_error("gn_reduce_partials_kernel launch: %s", cudaGetErrorString(le)); (void)cudaGetLastError(); return SDB_ERR_CUDA; }
  return check_launch("gn_reduce_partials_kernel");
}

extern "C" int sdb_groupnorm_apply(const void* x0, const void* x1, const double* stats,
                                   const float* gamma, const float* beta, void* out, int NB,
                                   long long HW, int C0, int C1, int groups, float eps, int silu,
                                   int x0_fp32, int x1_fp32, int stat_chunks, int out_f16, void* stream) {
  using namespace sdb;
  const int ctot = C0 + C1;
  if (!x0 || !stats || !gamma || !beta || !out || NB <= 0 || HW <= 0 || groups <= 0 || groups > 64 ||
      ctot % groups != 0 || C0 % 8 != 0 || C1 % 8 != 0 || (C1 > 0 && !x1)) {
    set_error("sdb_groupnorm_apply: bad arguments");
    return SDB_ERR_ARG;
  }
  int V, lanes, threads, chunks; long long ppc;
  static int bps = 0;                   // resident blocks per SM the grid aims for (tuning switch, default 4)
  if (bps == 0) { const char* ev = getenv("SDB_GN_APPLY_BPS"); bps = (ev && atoi(ev) > 0) ? atoi(ev) : 4; }
  if (gn_geometry(NB, HW, ctot, &V, &lanes, &threads, &ppc, &chunks, bps)) {
    set_error("sdb_groupnorm_apply: unsupported channel count %d", ctot);
    return SDB_ERR_UNSUPPORTED;
  }
  if (threads > 320 || HW * (long long)ctot >= (1LL << 31)) {
    set_error("sdb_groupnorm_apply: %d channels x %lld pixels per sample is outside the kernel's range", ctot, HW);
    return SDB_ERR_UNSUPPORTED;
  }
  const int sc = stat_chunks > 0 ? stat_chunks : chunks;
  const size_t smem_apply = (size_t)sc * groups * 2 * sizeof(double);
  static PerDeviceOnce apply_once = {};
  if (first_use_on_device(apply_once)) {
    int rc = set_max_smem(gn_apply_kernel<true>, GN_MAX_CHUNKS * 64 * 2 * (int)sizeof(double), "sdb_groupnorm_apply");
    if (!rc) rc = set_max_smem(gn_apply_kernel<false>, GN_MAX_CHUNKS * 64 * 2 * (int)sizeof(double), "sdb_groupnorm_apply");
    if (rc) return rc;
  }
  const bool all16 = x0_fp32 != 1 && (C1 == 0 || x1_fp32 != 1);
  cudaError_t le = all16
      ? launch_k(gn_apply_kernel<true>, dim3(chunks, NB), dim3(threads), smem_apply, (cudaStream_t)stream, 1,
                 x0, x1, stats, gamma, beta, (__nv_bfloat16*)out, HW, C0, C1, groups, eps, silu, ppc, V,
                 lanes, x0_fp32, x1_fp32, sc, out_f16)
      : launch_k(gn_apply_kernel<false>, dim3(chunks, NB), dim3(threads), smem_apply, (cudaStream_t)stream, 1,
                 x0, x1, stats, gamma, beta, (__nv_bfloat16*)out, HW, C0, C1, groups, eps, silu, ppc, V,
                 lanes, x0_fp32, x1_fp32, sc, out_f16);
  if (le != cudaSuccess) { set_error("gn_apply_kernel launch: %s", cudaGetErrorString(le)); (void)cudaGetLastError(); return SDB_ERR_CUDA; }
  return check_launch("gn_apply_kernel");
}

// 0: no plan; 1: plan with a cluster of several CTAs (measured slower than stats + apply on B200, whose second
// read hits L2: 73 vs 46 us for 16 x 64 x 64 x 320); 2: single-CTA plan at two CTAs per SM (measured faster:
// 14.9 vs 18.3 us for 16 x 16 x 16 x 1280) - what the engines use.
extern "C" int sdb_groupnorm_fused_supported(long long HW, int C0, int C1, int groups) {
  int G, CS, PX; size_t smem;
  if (!sdb::gn_fused_plan(HW, C0, C1, groups, &G, &CS, &PX, &smem)) return 0;
  return (CS == 1 && smem <= 84 * 1024) ? 2 : 1;
}

extern "C" int sdb_groupnorm_fused(const float* x0, const float* x1, const float* gamma, const float* beta,
                                   void* out, int NB, long long HW, int C0, int C1, int groups, float eps,
                                   int silu, int out_f16, void* stream) {
  using namespace sdb;
  if (!x0 || !gamma || !beta || !out || NB <= 0 || (C1 > 0 && !x1)) {
    set_error("sdb_groupnorm_fused: bad arguments");
    return SDB_ERR_ARG;
  }
  GnFusedParams p;
  p.out_f16 = out_f16;
  size_t smem;
  if (!gn_fused_plan(HW, C0, C1, groups, &p.G, &p.CS, &p.PX, &smem)) {
    set_error("sdb_groupnorm_fused: no plan for HW=%lld C=%d+%d groups=%d", HW, C0, C1, groups);
    return SDB_ERR_UNSUPPORTED;
  }
  if (((reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(out)) & 15u) != 0) {
    set_error("sdb_groupnorm_fused: pointers must be 16-byte aligned");
    return SDB_ERR_UNSUPPORTED;
  }
  p.x0 = x0; p.x1 = x1; p.gamma = gamma; p.beta = beta; p.out = (__nv_bfloat16*)out;
  p.HW = (int)HW; p.C0 = C0; p.C1 = C1; p.cpg = (C0 + C1) / groups; p.SC = p.G * p.cpg;
  p.nslabs = groups / p.G; p.silu = silu; p.eps = eps;
  static PerDeviceOnce fused_once = {};
  if (first_use_on_device(fused_once)) {
    int rc = set_max_smem(gn_fused_kernel, 200 * 1024, "sdb_groupnorm_fused");
    if (rc) return rc;
  }
  cudaError_t e = launch_k(gn_fused_kernel, dim3((unsigned)((long long)NB * p.nslabs * p.CS)), dim3(GNF_THREADS), smem,
                           (cudaStream_t)stream, p.CS, p);
  if (e != cudaSuccess) {
    set_error("gn_fused_kernel launch: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return SDB_ERR_CUDA;
  }
  return check_launch("gn_fused_kernel");
}


extern "C" int sdb_layernorm(const void* x, const float* gamma, const float* beta, void* out,
                             long long rows, int C, float eps, int in_fp32, int out_fp32, void* stream) {
  using namespace sdb;
  if (!x || !gamma || !beta || !out || rows <= 0 || C % 8 != 0 || C > LN_MAX_VEC_PER_LANE * 256) {
    set_error("sdb_layernorm: bad arguments (C=%d)", C);
    return SDB_ERR_ARG;
  }
  const int warps = 8;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                         reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15u) == 0;
  if (in_fp32 == 1 && aligned && C % 4 == 0 && C <= 1280) {
    const int vpl = (C / 4 + 31) / 32;
    long long blocks = ((rows + 1) / 2 + warps - 1) / warps;
    const long long cap = 148LL * 8;
    if (blocks > cap) blocks = cap;
    const float* xf = reinterpret_cast<const float*>(x);
    cudaStream_t st = (cudaStream_t)stream;
#define SDB_LN(V) (void)launch_k(layernorm_f32_kernel<V>, dim3((unsigned)blocks), dim3(warps * 32), 0, st, 1, xf, gamma, beta, out, rows, C, eps, out_fp32)
    if (vpl <= 3) SDB_LN(3);
    else if (vpl <= 5) SDB_LN(5);
    else if (vpl <= 6) SDB_LN(6);
    else SDB_LN(10);
#undef SDB_LN
    return check_launch("layernorm_f32_kernel");
  }
  if (in_fp32 == 2 && aligned && C % 4 == 0 && C <= 1280) {
    const int vpl = (C / 4 + 31) / 32;
    const int rpw = vpl <= 6 ? 4 : 2;                         // rows per warp iteration
    long long blocks = ((rows + rpw - 1) / rpw + warps - 1) / warps;
    const long long cap = 148LL * 8;
    if (blocks > cap) blocks = cap;
    const __half* xh = reinterpret_cast<const __half*>(x);
    cudaStream_t st = (cudaStream_t)stream;
#define SDB_LNH(V, R) (void)launch_k(layernorm_h16_kernel<V, R>, dim3((unsigned)blocks), dim3(warps * 32), 0, st, 1, xh, gamma, beta, out, rows, C, eps, out_fp32)
    if (vpl <= 3) SDB_LNH(3, 4);
    else if (vpl <= 5) SDB_LNH(5, 4);
    else if (vpl <= 6) SDB_LNH(6, 4);
    else SDB_LNH(10, 2);
#undef SDB_LNH
    return check_launch("layernorm_h16_kernel");
  }
  const long long blocks = (rows + warps - 1) / warps;
  layernorm_kernel<<<(unsigned)blocks, warps * 32, 0, (cudaStream_t)stream>>>(
      x, gamma, beta, out, rows, C, eps, in_fp32, out_fp32);
  return check_launch("layernorm_kernel");
}

extern "C" int sdb_softmax_rows(const float* scores, void* probs, long long rows, int cols,
                                float scale, void* stream) {
  using namespace sdb;
  if (!scores || !probs || rows <= 0 || cols <= 0) {
    set_error("sdb_softmax_rows: bad arguments");
    return SDB_ERR_ARG;
  }
  if (cols > 24576) {
    set_error("sdb_softmax_rows: %d columns exceed the 24576 a row can hold in shared memory "
              "(VAE attention above 1248 x 1248 pixels)", cols);
    return SDB_ERR_UNSUPPORTED;
  }
  static PerDeviceOnce softmax_once = {};
  if (first_use_on_device(softmax_once)) {
    int rc = set_max_smem(softmax_rows_kernel, 100 * 1024, "sdb_softmax_rows");
    if (rc) return rc;
  }
  softmax_rows_kernel<<<(unsigned)rows, 256, cols * sizeof(float), (cudaStream_t)stream>>>(
      scores, (__nv_bfloat16*)probs, cols, scale * 1.4426950408889634f);
  return check_launch("softmax_rows_kernel");
}
