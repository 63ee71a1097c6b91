// Layout conversion, small direct convolutions, the time path, the fused CFG + DDPM step and the
// pre/post-processing kernels. None of these is GEMM-shaped; they are sized for coalesced HBM
// traffic (or are latency-bound at a few KiB, like the sampler step).
#include "common.cuh"
#include "host.h"
#include "../../include/sdb200.h"

namespace sdb {

static inline int grid_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = 148LL * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

__global__ void fill_zero_kernel(uint4* p, long long n16, unsigned char* tail, int ntail) {
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = i0; i < n16; i += step) p[i] = make_uint4(0, 0, 0, 0);
  if (i0 < ntail) tail[i0] = 0;
}

__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ x, void* __restrict__ out,
                                             int NB, int C, int H, int W, int repeat, float scale,
                                             int out_fp32) {
  const long long hw = (long long)H * W;
  const long long total = (long long)NB * repeat * hw * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long p = (i / C) % hw;
    const int n = (int)(i / (C * hw));
    const int ns = n % NB;
    const float v = x[((long long)ns * C + c) * hw + p] * scale;
    if (out_fp32) reinterpret_cast<float*>(out)[i] = v;
    else reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
  }
}

__global__ void nhwc_to_nchw_f32_kernel(const void* __restrict__ x, float* __restrict__ out, int NB,
                                        int C, int H, int W, int in_fp32) {
  const long long hw = (long long)H * W;
  const long long total = (long long)NB * hw * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % hw;
    const int c = (int)((i / hw) % C);
    const int n = (int)(i / (hw * C));
    const long long src = ((long long)n * hw + p) * C + c;
    out[i] = in_fp32 ? reinterpret_cast<const float*>(x)[src]
                     : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[src]);
  }
}

__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int NB, int H,
                                  int W, int V) {
  pdl_trigger();
  pdl_wait();
  const int Ho = 2 * H, Wo = 2 * W;
  const long long total = (long long)NB * Ho * Wo * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % V);
    long long r = i / V;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int n = (int)(r / Ho);
    out[i] = __ldg(x + (((long long)n * H + (ho >> 1)) * W + (wo >> 1)) * V + v);
  }
}

// Direct convolution, Cin <= 8: one thread = one output pixel x 8 output channels; the weights sit in
// shared memory transposed to [k*k*Cin][Cout] so a warp reads consecutive words.
template <int KS>
__global__ void conv_direct_kernel(const void* __restrict__ x, const float* __restrict__ w,
                                   const float* __restrict__ bias, void* __restrict__ out,
                                   __nv_bfloat16* __restrict__ out2, int NB, int H, int W, int Cin,
                                   int Cout, int out_fp32, int in_fp32, int out2_f16) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_w[];  // [KS*KS*Cin][CoutPad]
  const int CoutPad = (Cout + 7) & ~7;
  const int K = KS * KS * Cin;
  for (int i = threadIdx.x; i < K * CoutPad; i += blockDim.x) {
    const int co = i % CoutPad, k = i / CoutPad;
    s_w[i] = (co < Cout) ? w[(long long)co * K + k] : 0.f;
  }
  __syncthreads();
  const int G = CoutPad >> 3;
  const long long total = (long long)NB * H * W * G;
  constexpr int PAD = (KS - 1) / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long r = i / G;
    const int wo = (int)(r % W); r /= W;
    const int ho = (int)(r % H);
    const int n = (int)(r / H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = g * 8 + j;
      acc[j] = (bias != nullptr && co < Cout) ? bias[co] : 0.f;
    }
    for (int ky = 0; ky < KS; ++ky) {
      const int hi = ho + ky - PAD;
      if (hi < 0 || hi >= H) continue;
      for (int kx = 0; kx < KS; ++kx) {
        const int wi = wo + kx - PAD;
        if (wi < 0 || wi >= W) continue;
        const long long poff = (((long long)n * H + hi) * W + wi) * Cin;
        const float* wk = s_w + (long long)((ky * KS + kx) * Cin) * CoutPad + g * 8;
        for (int ci = 0; ci < Cin; ++ci) {
          const float a = in_fp32 ? reinterpret_cast<const float*>(x)[poff + ci]
                                  : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[poff + ci]);
          const float4 w0 = *reinterpret_cast<const float4*>(wk + ci * CoutPad);
          const float4 w1 = *reinterpret_cast<const float4*>(wk + ci * CoutPad + 4);
          acc[0] += a * w0.x; acc[1] += a * w0.y; acc[2] += a * w0.z; acc[3] += a * w0.w;
          acc[4] += a * w1.x; acc[5] += a * w1.y; acc[6] += a * w1.z; acc[7] += a * w1.w;
        }
      }
    }
    const long long obase = (((long long)n * H + ho) * W + wo) * Cout + g * 8;
    if ((Cout & 7) == 0) {
      // eight channels per thread, vector stores: consecutive threads write consecutive 32 (16) bytes - the scalar
      // form below took 258 us for the UNet stem (16 x 64 x 64 x 4 -> 320), 13 x its memory time
      const int h16 = out_fp32 ? out2_f16 : 0;        // the 16-bit shadow may be IEEE half; a 16-bit `out` is bf16
      const uint4 h = make_uint4(pack16x2(acc[0], acc[1], h16), pack16x2(acc[2], acc[3], h16),
                                 pack16x2(acc[4], acc[5], h16), pack16x2(acc[6], acc[7], h16));
      if (out_fp32) {
        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + obase);
        o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        if (out2 != nullptr) *reinterpret_cast<uint4*>(out2 + obase) = h;
      } else {
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + obase) = h;
      }
      continue;
    }
    for (int j = 0; j < 8; ++j) {
      if (g * 8 + j < Cout) {
        if (out_fp32) {
          reinterpret_cast<float*>(out)[obase + j] = acc[j];
          if (out2 != nullptr) reinterpret_cast<unsigned short*>(out2)[obase + j] = cvt16(acc[j], out2_f16);
        } else {
          reinterpret_cast<__nv_bfloat16*>(out)[obase + j] = __float2bfloat16_rn(acc[j]);
        }
      }
    }
  }
}

// Same convolution, 4 adjacent output pixels x 8 output channels per thread (W % 4 == 0, Cout % 8 == 0): the
// shared-memory weight reads bound the one-pixel form (two LDS.128 with a 2-way bank conflict per 8 FMAs:
// 258 us for the UNet stem, 13 x its memory time); here every weight vector feeds 32 FMAs.
template <int KS>
__global__ void __launch_bounds__(256) conv_direct4_kernel(const void* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, void* __restrict__ out,
                                                           __nv_bfloat16* __restrict__ out2, int NB, int H, int W,
                                                           int Cin, int Cout, int out_fp32, int in_fp32, int out2_f16) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_w[];  // [KS*KS*Cin][Cout]
  const int K = KS * KS * Cin;
  for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
    const int co = i % Cout, k = i / Cout;
    s_w[i] = w[(long long)co * K + k];
  }
  __syncthreads();
  const int G = Cout >> 3;
  const int W4 = W >> 2;
  const long long total = (long long)NB * H * W4 * G;
  constexpr int PAD = (KS - 1) / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    long long r = i / G;
    const int wq = (int)(r % W4); r /= W4;
    const int ho = (int)(r % H);
    const int n = (int)(r / H);
    const int wo0 = wq * 4;
    float acc[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float b = (bias != nullptr) ? bias[g * 8 + j] : 0.f;
#pragma unroll
      for (int px = 0; px < 4; ++px) acc[px][j] = b;
    }
    for (int ky = 0; ky < KS; ++ky) {
      const int hi = ho + ky - PAD;
      if (hi < 0 || hi >= H) continue;
      const long long rowoff = ((long long)n * H + hi) * W;
      for (int ci = 0; ci < Cin; ++ci) {
        // the KS + 3 input values of this row / channel that the 4 pixels touch
        float a[KS + 3];
#pragma unroll
        for (int t = 0; t < KS + 3; ++t) {
          const int wi = wo0 + t - PAD;
          float v = 0.f;
          if (wi >= 0 && wi < W) {
            const long long off = (rowoff + wi) * Cin + ci;
            v = in_fp32 ? reinterpret_cast<const float*>(x)[off]
                        : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[off]);
          }
          a[t] = v;
        }
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const float* wk = s_w + (long long)((ky * KS + kx) * Cin + ci) * Cout + g * 8;
          const float4 w0 = *reinterpret_cast<const float4*>(wk);
          const float4 w1 = *reinterpret_cast<const float4*>(wk + 4);
#pragma unroll
          for (int px = 0; px < 4; ++px) {
            const float av = a[px + kx];
            acc[px][0] = fmaf(av, w0.x, acc[px][0]); acc[px][1] = fmaf(av, w0.y, acc[px][1]);
            acc[px][2] = fmaf(av, w0.z, acc[px][2]); acc[px][3] = fmaf(av, w0.w, acc[px][3]);
            acc[px][4] = fmaf(av, w1.x, acc[px][4]); acc[px][5] = fmaf(av, w1.y, acc[px][5]);
            acc[px][6] = fmaf(av, w1.z, acc[px][6]); acc[px][7] = fmaf(av, w1.w, acc[px][7]);
          }
        }
      }
    }
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      const long long obase = (((long long)n * H + ho) * W + wo0 + px) * Cout + g * 8;
      const int h16 = out_fp32 ? out2_f16 : 0;
      const uint4 h = make_uint4(pack16x2(acc[px][0], acc[px][1], h16), pack16x2(acc[px][2], acc[px][3], h16),
                                 pack16x2(acc[px][4], acc[px][5], h16), pack16x2(acc[px][6], acc[px][7], h16));
      if (out_fp32) {
        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + obase);
        o[0] = make_float4(acc[px][0], acc[px][1], acc[px][2], acc[px][3]);
        o[1] = make_float4(acc[px][4], acc[px][5], acc[px][6], acc[px][7]);
        if (out2 != nullptr) *reinterpret_cast<uint4*>(out2 + obase) = h;
      } else {
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + obase) = h;
      }
    }
  }
}

// One warp per output element (r, n).
__global__ void small_linear_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                    const float* __restrict__ bias, float* __restrict__ out, int R,
                                    int K, int N, int act_in, int act_out) {
  const long long gw = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= (long long)R * N) return;
  const int n = (int)(gw % N);
  const int r = (int)(gw / N);
  const float* xr = x + (long long)r * K;
  const __nv_bfloat16* wr = w + (long long)n * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    float a = xr[k];
    if (act_in == SDB_ACT_SILU) a = a / (1.0f + expf(-a));
    acc += a * __bfloat162float(wr[k]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    if (bias) acc += bias[n];
    if (act_out == SDB_ACT_SILU) acc = acc / (1.0f + expf(-acc));
    out[(long long)r * N + n] = acc;
  }
}

__global__ void cfg_ddpm_step_kernel(float* __restrict__ latents, const float* __restrict__ eps,
                                     const float* __restrict__ noise, const float* __restrict__ coef,
                                     int step, float cfg_scale, int do_cfg,
                                     void* __restrict__ next_in, int NB, int C, int H, int W,
                                     int eps_nchw, int next_fp32) {
  pdl_trigger();
  pdl_wait();
  const long long hw = (long long)H * W;
  const long long total = (long long)NB * C * hw;
  const float sb = coef[step * 5 + 0];   // sqrt(1 - abar_t)
  const float sa = coef[step * 5 + 1];   // sqrt(abar_t)
  const float c_x0 = coef[step * 5 + 2];
  const float c_xt = coef[step * 5 + 3];
  const float sigma = coef[step * 5 + 4];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % hw;
    const int c = (int)((i / hw) % C);
    const int n = (int)(i / (hw * C));
    const long long e_idx = ((long long)n * hw + p) * C + c;   // NHWC position (next_in layout)
    const long long s_idx = eps_nchw ? i : e_idx;
    float e;
    if (do_cfg) {
      const float ec = eps[s_idx];
      const float eu = eps[s_idx + (long long)NB * hw * C];
      e = cfg_scale * (ec - eu) + eu;
    } else {
      e = eps[s_idx];
    }
    const float xt = latents[i];
    const float x0 = (xt - sb * e) / sa;
    float xn = c_x0 * x0 + c_xt * xt;
    if (sigma != 0.f && noise != nullptr) xn += sigma * noise[i];
    latents[i] = xn;
    if (next_in != nullptr) {
      if (next_fp32) {
        float* ni = reinterpret_cast<float*>(next_in);
        ni[e_idx] = xn;
        if (do_cfg) ni[e_idx + (long long)NB * hw * C] = xn;
      } else {
        __nv_bfloat16* ni = reinterpret_cast<__nv_bfloat16*>(next_in);
        const __nv_bfloat16 b = __float2bfloat16_rn(xn);
        ni[e_idx] = b;
        if (do_cfg) ni[e_idx + (long long)NB * hw * C] = b;
      }
    }
  }
}

// out[n][p][c] = y_flat[n][c*HW + p] + res[n][p][c]: a 32x32 shared-memory transpose of y viewed as
// [C][HW].
__global__ void vae_scramble_add_kernel(const __nv_bfloat16* __restrict__ y,
                                        const float* __restrict__ res, float* __restrict__ out,
                                        __nv_bfloat16* __restrict__ out2, long long HW, int C) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const __nv_bfloat16* yn = y + (long long)n * HW * C;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j;
    const long long p = p0 + tx;
    tile[j][tx] = (c < C && p < HW) ? __bfloat162float(yn[(long long)c * HW + p]) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const long long p = p0 + j;
    const int c = c0 + tx;
    if (p < HW && c < C) {
      const long long o = ((long long)n * HW + p) * C + c;
      const float v = tile[tx][j] + res[o];
      out[o] = v;
      if (out2 != nullptr) out2[o] = __float2bfloat16_rn(v);
    }
  }
}

__global__ void vae_encode_tail_kernel(const float* __restrict__ moments, const float* __restrict__ noise,
                                       float* __restrict__ out, int NB, int H, int W) {
  const long long hw = (long long)H * W;
  const long long total = (long long)NB * 4 * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % hw;
    const int c = (int)((i / hw) % 4);
    const int n = (int)(i / (hw * 4));
    const float* m = moments + ((long long)n * hw + p) * 8;
    const float mean = m[c];
    float lv = m[4 + c];
    lv = fminf(fmaxf(lv, -30.f), 20.f);
    const float stdev = sqrtf(expf(lv));
    out[i] = (mean + stdev * noise[i]) * 0.18215f;
  }
}

// fp32 -> 16-bit shadow (bf16, or IEEE half when f16), four elements per thread when n % 4 == 0
__global__ void f32_to_bf16_kernel(const float* __restrict__ x, unsigned short* __restrict__ out, long long n, int f16) {
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, step = (long long)gridDim.x * blockDim.x;
  if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
    for (long long i = i0; i < (n >> 2); i += step) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
      reinterpret_cast<uint2*>(out)[i] = make_uint2(pack16x2(v.x, v.y, f16), pack16x2(v.z, v.w, f16));
    }
    return;
  }
  for (long long i = i0; i < n; i += step) out[i] = cvt16(x[i], f16);
}

__global__ void axpby_kernel(const float* __restrict__ x, const float* __restrict__ y,
                             float* __restrict__ out, float a, float b, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = a * x[i] + b * y[i];
}

__global__ void image_to_uint8_kernel(const float* __restrict__ x, unsigned char* __restrict__ out,
                                      long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    // rescale (-1,1) -> (0,255) in the reference's operation order: x -= -1; x *= 255/2; x += 0
    // (explicit round-to-nearest intrinsics: no FMA contraction, bit-identical to the three torch ops)
    float v = __fmul_rn(__fsub_rn(x[i], -1.0f), 127.5f);
    v = __fadd_rn(v, 0.0f);
    v = fminf(fmaxf(v, 0.f), 255.f);
    out[i] = (unsigned char)v;  // truncation, as torch's float -> uint8 cast
  }
}

__global__ void uint8_to_image_kernel(const unsigned char* __restrict__ x, void* __restrict__ out,
                                      long long n, int out_fp32) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    // x -= 0; x *= 2/255; x += -1 (sd/pipeline.py:295-301) as three separately rounded fp32 operations
    float v = __fadd_rn(__fmul_rn((float)x[i], 2.0f / 255.0f), -1.0f);
    if (out_fp32) reinterpret_cast<float*>(out)[i] = v;
    else reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
  }
}

__global__ void clip_embed_kernel(const long long* __restrict__ tokens, const float* __restrict__ table,
                                  const float* __restrict__ pos, float* __restrict__ out, int NB,
                                  int T, int T_pad, int D, int vocab) {
  const long long total = (long long)NB * T_pad * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const int t = (int)((i / D) % T_pad);
    const int b = (int)(i / ((long long)D * T_pad));
    float v = 0.f;
    if (t < T) {
      long long tok = tokens[(long long)b * T + t];
      if (tok < 0) tok = 0;
      if (tok >= vocab) tok = vocab - 1;
      v = table[tok * D + d] + pos[(long long)t * D + d];
    }
    out[i] = v;
  }
}

// Pack-time composition of affine maps in double precision: C[M, N] = A[M, K] . B[K, N] (row-major, A / B fp32 or
// fp64, C fp64). Used once per model load to fold linear_geglu_2 . linear_geglu_1[:4C] and conv_output into one
// matrix (sd/diffusion.py:355-381 has no non-linearity between them); products of two fp32 values are exact in fp64.
// 64 x 64 tile per block, 4 x 4 outputs per thread, K walked 16 at a time through shared memory.
template <typename TA, typename TB>
__global__ void __launch_bounds__(256) matmul_f64_kernel(const TA* __restrict__ A, const TB* __restrict__ B,
                                                         double* __restrict__ C, int M, int N, int K) {
  __shared__ double sa[16][64 + 1];
  __shared__ double sb[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, kk = i & 15;          // A tile: 64 rows x 16 k
      const int gm = m0 + r, gk = k0 + kk;
      sa[kk][r] = (gm < M && gk < K) ? (double)A[(long long)gm * K + gk] : 0.0;
      const int kb = i >> 6, c = i & 63;          // B tile: 16 k x 64 columns
      const int gk2 = k0 + kb, gn = n0 + c;
      sb[kb][c] = (gk2 < K && gn < N) ? (double)B[(long long)gk2 * N + gn] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sa[kk][ty * 4 + i]; b[i] = sb[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
      if (gm < M && gn < N) C[(long long)gm * N + gn] = acc[i][j];
    }
}

// One pass of Pillow's 8-bit separable resampler (ImagingResampleHorizontal_8bpc / Vertical_8bpc, what
// PIL.Image.resize runs for the reference's input_image.resize((WIDTH, HEIGHT)), sd/pipeline.py:156) over uint8 NHWC:
// out[.., o, ..] = clip8((2^21 + sum_k src[.., lo_o + k, ..] * coef[o][k]) >> 22) with the host-computed fixed-point
// taps (22 fractional bits) and windows [lo_o, lo_o + cnt_o). axis 0 = along W, 1 = along H. With `img` non-null the
// pass also writes the pre-processed fp32 value x * (2/255) - 1 (sd/pipeline.py:162-173), so resize + rescale of the
// last pass is one kernel. Bit-exact against Pillow (tests/test_kernels_gpu.py).
__global__ void resample_u8_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst,
                                   float* __restrict__ img, int NB, int H, int W, int C, int out_size, int axis,
                                   const int* __restrict__ bounds, const int* __restrict__ coef, int ksize) {
  const int HO = axis ? out_size : H, WO = axis ? W : out_size;
  const long long total = (long long)NB * HO * WO * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int x = (int)((i / C) % WO);
    const int y = (int)((i / ((long long)C * WO)) % HO);
    const int n = (int)(i / ((long long)C * WO * HO));
    const int o = axis ? y : x;
    const int lo = bounds[2 * o], cnt = bounds[2 * o + 1];
    const int* k = coef + (long long)o * ksize;
    int acc = 1 << 21;
    if (axis) {
      const unsigned char* p = src + (((long long)n * H + lo) * W + x) * C + c;
      for (int j = 0; j < cnt; ++j) acc += (int)p[(long long)j * W * C] * k[j];
    } else {
      const unsigned char* p = src + (((long long)n * H + y) * W + lo) * C + c;
      for (int j = 0; j < cnt; ++j) acc += (int)p[(long long)j * C] * k[j];
    }
    acc >>= 22;
    const unsigned char v = (unsigned char)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
    dst[i] = v;
    if (img != nullptr) img[i] = __fadd_rn(__fmul_rn((float)v, 2.0f / 255.0f), -1.0f);
  }
}

}  // namespace sdb

using namespace sdb;
#define SDB_STREAM ((cudaStream_t)stream)

extern "C" int sdb_fill_zero(void* ptr, long long bytes, void* stream) {
  if (!ptr || bytes < 0) { set_error("sdb_fill_zero: bad arguments"); return SDB_ERR_ARG; }
  if (bytes == 0) return SDB_OK;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0) {
    cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, SDB_STREAM);
    if (e != cudaSuccess) { set_error("sdb_fill_zero: %s", cudaGetErrorString(e)); return SDB_ERR_CUDA; }
    return SDB_OK;
  }
  const long long n16 = bytes / 16;
  const int ntail = (int)(bytes % 16);
  fill_zero_kernel<<<grid_for(n16 > 0 ? n16 : 1, 256), 256, 0, SDB_STREAM>>>(
      (uint4*)ptr, n16, (unsigned char*)ptr + n16 * 16, ntail);
  return check_launch("fill_zero_kernel");
}

extern "C" int sdb_nchw_f32_to_nhwc(const float* x, void* out, int NB, int C, int H, int W,
                                    int repeat, float scale, int out_fp32, void* stream) {
  if (!x || !out || NB <= 0 || C <= 0 || H <= 0 || W <= 0 || repeat <= 0) {
    set_error("sdb_nchw_f32_to_nhwc: bad arguments"); return SDB_ERR_ARG;
  }
  const long long total = (long long)NB * repeat * C * H * W;
  nchw_f32_to_nhwc_bf16_kernel<<<grid_for(total, 256), 256, 0, SDB_STREAM>>>(
      x, out, NB, C, H, W, repeat, scale, out_fp32);
  return check_launch("nchw_f32_to_nhwc_bf16_kernel");
}

extern "C" int sdb_nhwc_to_nchw_f32(const void* x, float* out, int NB, int C, int H, int W,
                                    int in_fp32, void* stream) {
  if (!x || !out || NB <= 0 || C <= 0 || H <= 0 || W <= 0) {
    set_error("sdb_nhwc_to_nchw_f32: bad arguments"); return SDB_ERR_ARG;
  }
  const long long total = (long long)NB * C * H * W;
  nhwc_to_nchw_f32_kernel<<<grid_for(total, 256), 256, 0, SDB_STREAM>>>(x, out, NB, C, H, W, in_fp32);
  return check_launch("nhwc_to_nchw_f32_kernel");
}

extern "C" int sdb_upsample2x_nhwc(const void* x, void* out, int NB, int H, int W, int C, void* stream) {
  if (!x || !out || NB <= 0 || H <= 0 || W <= 0 || C % 8 != 0) {
    set_error("sdb_upsample2x_nhwc: bad arguments"); return SDB_ERR_ARG;
  }
  const long long total = (long long)NB * 4 * H * W * (C / 8);
  (void)launch_k(upsample2x_kernel, dim3(grid_for(total, 256)), dim3(256), 0, SDB_STREAM, 1, (const uint4*)x, (uint4*)out, NB, H, W,
                                                                  C / 8);
  return check_launch("upsample2x_kernel");
}

extern "C" int sdb_conv_direct(const void* x, const float* w, const float* bias, void* out, void* out2,
                               int NB, int H, int W, int Cin, int Cout, int ksize, int out_fp32,
                               int in_fp32, int out2_f16, void* stream) {
  if (!x || !w || !out || NB <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cin > 8 || Cout <= 0 ||
      (ksize != 1 && ksize != 3) || (out2 && !out_fp32)) {
    set_error("sdb_conv_direct: bad arguments (Cin=%d Cout=%d k=%d)", Cin, Cout, ksize);
    return SDB_ERR_ARG;
  }
  const int CoutPad = (Cout + 7) & ~7;
  const size_t smem = (size_t)ksize * ksize * Cin * CoutPad * sizeof(float);
  if (smem > 160 * 1024) { set_error("sdb_conv_direct: weights too large"); return SDB_ERR_UNSUPPORTED; }
  const long long total = (long long)NB * H * W * (CoutPad / 8);
  static PerDeviceOnce direct_once = {};
  if (first_use_on_device(direct_once)) {
    int rc = set_max_smem(conv_direct_kernel<1>, 160 * 1024, "sdb_conv_direct");
    if (!rc) rc = set_max_smem(conv_direct_kernel<3>, 160 * 1024, "sdb_conv_direct");
    if (!rc) rc = set_max_smem(conv_direct4_kernel<1>, 160 * 1024, "sdb_conv_direct");
    if (!rc) rc = set_max_smem(conv_direct4_kernel<3>, 160 * 1024, "sdb_conv_direct");
    if (rc) return rc;
  }
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if ((W & 3) == 0 && (Cout & 7) == 0) {
    const long long total4 = (long long)NB * H * (W / 4) * (Cout / 8);
    long long blocks4 = (total4 + 255) / 256;
    if (blocks4 > 148 * 4) blocks4 = 148 * 4;
    const size_t smem4 = (size_t)ksize * ksize * Cin * Cout * sizeof(float);
    if (ksize == 1)
      (void)launch_k(conv_direct4_kernel<1>, dim3((unsigned)blocks4), dim3(256), smem4, SDB_STREAM, 1,
          x, w, bias, out, (__nv_bfloat16*)out2, NB, H, W, Cin, Cout, out_fp32, in_fp32, out2_f16);
    else
      (void)launch_k(conv_direct4_kernel<3>, dim3((unsigned)blocks4), dim3(256), smem4, SDB_STREAM, 1,
          x, w, bias, out, (__nv_bfloat16*)out2, NB, H, W, Cin, Cout, out_fp32, in_fp32, out2_f16);
    return check_launch("conv_direct4_kernel");
  }
  if (ksize == 1)
    (void)launch_k(conv_direct_kernel<1>, dim3((unsigned)blocks), dim3(256), smem, SDB_STREAM, 1,
        x, w, bias, out, (__nv_bfloat16*)out2, NB, H, W, Cin, Cout, out_fp32, in_fp32, out2_f16);
  else
    (void)launch_k(conv_direct_kernel<3>, dim3((unsigned)blocks), dim3(256), smem, SDB_STREAM, 1,
        x, w, bias, out, (__nv_bfloat16*)out2, NB, H, W, Cin, Cout, out_fp32, in_fp32, out2_f16);
  return check_launch("conv_direct_kernel");
}

extern "C" int sdb_small_linear(const float* x, const void* w, const float* bias, float* out, int R,
                                int K, int N, int act_in, int act_out, void* stream) {
  if (!x || !w || !out || R <= 0 || K <= 0 || N <= 0) {
    set_error("sdb_small_linear: bad arguments"); return SDB_ERR_ARG;
  }
  const long long warps = (long long)R * N;
  const long long blocks = (warps + 7) / 8;
  small_linear_kernel<<<(unsigned)blocks, 256, 0, SDB_STREAM>>>(x, (const __nv_bfloat16*)w, bias, out, R, K,
                                                                N, act_in, act_out);
  return check_launch("small_linear_kernel");
}

extern "C" int sdb_cfg_ddpm_step(float* latents, const float* eps, const float* noise,
                                 const float* coef, int step, float cfg_scale, int do_cfg,
                                 void* next_in, int NB, int C, int H, int W, int eps_nchw,
                                 int next_fp32, void* stream) {
  if (!latents || !eps || !coef || step < 0 || NB <= 0 || C <= 0 || H <= 0 || W <= 0) {
    set_error("sdb_cfg_ddpm_step: bad arguments"); return SDB_ERR_ARG;
  }
  const long long total = (long long)NB * C * H * W;
  (void)launch_k(cfg_ddpm_step_kernel, dim3(grid_for(total, 256)), dim3(256), 0, SDB_STREAM, 1,
      latents, eps, noise, coef, step, cfg_scale, do_cfg, next_in, NB, C, H, W, eps_nchw, next_fp32);
  return check_launch("cfg_ddpm_step_kernel");
}

extern "C" int sdb_vae_attn_scramble_add(const void* y, const float* res, float* out, void* out2, int NB,
                                         long long HW, int C, void* stream) {
  if (!y || !res || !out || NB <= 0 || HW <= 0 || C <= 0) {
    set_error("sdb_vae_attn_scramble_add: bad arguments"); return SDB_ERR_ARG;
  }
  dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)NB);
  vae_scramble_add_kernel<<<grid, dim3(32, 8), 0, SDB_STREAM>>>(
      (const __nv_bfloat16*)y, res, out, (__nv_bfloat16*)out2, HW, C);
  return check_launch("vae_scramble_add_kernel");
}

extern "C" int sdb_vae_encode_tail(const float* moments, const float* noise, float* out, int NB, int H,
                                   int W, void* stream) {
  if (!moments || !noise || !out || NB <= 0 || H <= 0 || W <= 0) {
    set_error("sdb_vae_encode_tail: bad arguments"); return SDB_ERR_ARG;
  }
  const long long total = (long long)NB * 4 * H * W;
  vae_encode_tail_kernel<<<grid_for(total, 256), 256, 0, SDB_STREAM>>>(moments, noise, out, NB, H, W);
  return check_launch("vae_encode_tail_kernel");
}

extern "C" int sdb_f32_to_bf16(const float* x, void* out, long long n, int f16, void* stream) {
  if (!x || !out || n <= 0) { set_error("sdb_f32_to_bf16: bad arguments"); return SDB_ERR_ARG; }
  f32_to_bf16_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, SDB_STREAM>>>(x, (unsigned short*)out, n, f16);
  return check_launch("f32_to_bf16_kernel");
}

extern "C" int sdb_axpby(const float* x, const float* y, float* out, float a, float b, long long n,
                         void* stream) {
  if (!x || !y || !out || n <= 0) { set_error("sdb_axpby: bad arguments"); return SDB_ERR_ARG; }
  axpby_kernel<<<grid_for(n, 256), 256, 0, SDB_STREAM>>>(x, y, out, a, b, n);
  return check_launch("axpby_kernel");
}

extern "C" int sdb_image_to_uint8(const float* x, unsigned char* out, long long n, void* stream) {
  if (!x || !out || n <= 0) { set_error("sdb_image_to_uint8: bad arguments"); return SDB_ERR_ARG; }
  image_to_uint8_kernel<<<grid_for(n, 256), 256, 0, SDB_STREAM>>>(x, out, n);
  return check_launch("image_to_uint8_kernel");
}

extern "C" int sdb_uint8_to_image(const unsigned char* x, void* out, long long n, int out_fp32, void* stream) {
  if (!x || !out || n <= 0) { set_error("sdb_uint8_to_image: bad arguments"); return SDB_ERR_ARG; }
  uint8_to_image_kernel<<<grid_for(n, 256), 256, 0, SDB_STREAM>>>(x, out, n, out_fp32);
  return check_launch("uint8_to_image_kernel");
}

extern "C" int sdb_clip_embed(const long long* tokens, const float* table, const float* pos, void* out,
                              int NB, int T, int T_pad, int D, int vocab, void* stream) {
  if (!tokens || !table || !pos || !out || NB <= 0 || T <= 0 || T_pad < T || D <= 0 || vocab <= 0) {
    set_error("sdb_clip_embed: bad arguments"); return SDB_ERR_ARG;
  }
  const long long total = (long long)NB * T_pad * D;
  clip_embed_kernel<<<grid_for(total, 256), 256, 0, SDB_STREAM>>>(tokens, table, pos, (float*)out, NB,
                                                                  T, T_pad, D, vocab);
  return check_launch("clip_embed_kernel");
}

extern "C" int sdb_matmul_f64(const void* A, int a_f64, const void* B, int b_f64, double* C, int M, int N, int K,
                              void* stream) {
  using namespace sdb;
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) { set_error("sdb_matmul_f64: bad arguments"); return SDB_ERR_ARG; }
  const dim3 grid((N + 63) / 64, (M + 63) / 64);
  if (a_f64 && b_f64) matmul_f64_kernel<double, double><<<grid, 256, 0, SDB_STREAM>>>((const double*)A, (const double*)B, C, M, N, K);
  else if (a_f64) matmul_f64_kernel<double, float><<<grid, 256, 0, SDB_STREAM>>>((const double*)A, (const float*)B, C, M, N, K);
  else if (b_f64) matmul_f64_kernel<float, double><<<grid, 256, 0, SDB_STREAM>>>((const float*)A, (const double*)B, C, M, N, K);
  else matmul_f64_kernel<float, float><<<grid, 256, 0, SDB_STREAM>>>((const float*)A, (const float*)B, C, M, N, K);
  return check_launch("matmul_f64_kernel");
}

extern "C" int sdb_resample_u8(const unsigned char* src, unsigned char* dst, float* img, int NB, int H, int W, int C,
                               int out_size, int axis, const int* bounds, const int* coef, int ksize, void* stream) {
  using namespace sdb;
  if (!src || !dst || !bounds || !coef || NB <= 0 || H <= 0 || W <= 0 || C <= 0 || out_size <= 0 || ksize <= 0 ||
      (axis != 0 && axis != 1)) {
    set_error("sdb_resample_u8: bad arguments");
    return SDB_ERR_ARG;
  }
  const long long total = (long long)NB * (axis ? out_size : H) * (axis ? W : out_size) * C;
  resample_u8_kernel<<<grid_for(total, 256), 256, 0, SDB_STREAM>>>(src, dst, img, NB, H, W, C, out_size, axis, bounds,
                                                                    coef, ksize);
  return check_launch("resample_u8_kernel");
}

// Device-to-device copy as a stream-ordered memcpy node (legal under CUDA-graph capture): duplicating the shared prefix
// of a classifier-free-guidance pair (engine.UNetEngine.forward_nhwc, cfg_pairs).
extern "C" int sdb_copy_bytes(void* dst, const void* src, long long bytes, void* stream) {
  using namespace sdb;
  if (!dst || !src || bytes <= 0) { set_error("sdb_copy_bytes: bad arguments"); return SDB_ERR_ARG; }
  cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, SDB_STREAM);
  if (e != cudaSuccess) { set_error("sdb_copy_bytes: %s", cudaGetErrorString(e)); return SDB_ERR_CUDA; }
  return SDB_OK;
}
