// (each thread holds 64 columns so that 16 warps fit the register file; figures are scaled to 128 columns)
// Micro-benchmark: cycles per 128-column softmax row block (the attention inner loop on registers only) as a
// function of the number of warps per SM sub-partition and of the instruction mix.
//   (mode 0's max pass carries an extra FADD per element to stay loop-variant)
//   mode 0: max + FFMA + MUFU.EX2 + FADD row sum + pack          (plain)
//   mode 1: FFMA + 3/4 MUFU + 1/4 polynomial + pack, max tracked (streaming path of attn2_tc_kernel)
//   mode 2: as 1 without the max tracking
//   mode 3: FFMA + MUFU + pack only
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_issue softmax_issue.cu ; run on a B200.
#include <cstdio>
#include "../../pytorch_stable_diffusion_b200/csrc/common.cuh"
using namespace sdb;

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.0551716685f, f, 0.2426111400f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, float sl2, float mb, unsigned* sink, long long* cyc) {
  extern __shared__ float dummy[];
  float sv[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) sv[i] = -0.01f * (float)((threadIdx.x * 7 + i * 13) & 255);
  unsigned acc = 0;
  float lsum = 0.f, bm = -1e30f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[32];
    mb += 1e-4f;                       // loop-variant: nothing below can be hoisted out of the loop
    if (MODE == 0) {
      float m0 = -1e30f, m1 = -1e30f, m2 = -1e30f, m3 = -1e30f;
#pragma unroll
      for (int i = 0; i < 64; i += 4) { m0 = fmaxf(m0, sv[i] + mb); m1 = fmaxf(m1, sv[i + 1] + mb); m2 = fmaxf(m2, sv[i + 2] + mb); m3 = fmaxf(m3, sv[i + 3] + mb); }
      const float mbb = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * sl2 + mb;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const float p0 = ex2a(fmaf(sv[i], sl2, -mbb)), p1 = ex2a(fmaf(sv[i + 1], sl2, -mbb));
        l0 += p0; l1 += p1;
        pk[i >> 1] = pack_bf16x2(p0, p1);
      }
      lsum += l0 + l1;
    } else {
      float b0 = bm, b1 = bm, b2 = bm, b3 = bm;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        const float x0 = fmaf(sv[i], sl2, -mb), x1 = fmaf(sv[i + 1], sl2, -mb), x2 = fmaf(sv[i + 2], sl2, -mb), x3 = fmaf(sv[i + 3], sl2, -mb);
        if (MODE == 1) { b0 = fmaxf(b0, x0); b1 = fmaxf(b1, x1); b2 = fmaxf(b2, x2); b3 = fmaxf(b3, x3); }
        const float p0 = ex2a(x0), p1 = ex2a(x1), p2 = ex2a(x2);
        const float p3 = (MODE == 3) ? ex2a(x3) : exp2_poly(x3);
        pk[i >> 1] = pack_bf16x2(p0, p1);
        pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
      }
      bm = fmaxf(fmaxf(b0, b1), fmaxf(b2, b3));
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= pk[i];
    // keep the inputs changing so nothing is hoisted out of the loop
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(lsum) + __float_as_uint(bm);
}

template <int MODE>
void run(int warps, int iters, unsigned* sink, long long* cyc) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<MODE><<<148, warps * 32, 200 * 1024>>>(iters, 1.4427f, -3.0f, sink, cyc);
  cudaDeviceSynchronize();
  k<MODE><<<148, warps * 32, 200 * 1024>>>(iters, 1.4427f, -3.0f, sink, cyc);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  // row blocks per SMSP per iteration = warps / 4
  printf("mode %d  warps/SMSP %d : %8.1f cycles per iteration, %7.1f cycles per (128-col row block x warp)\n", MODE, warps / 4,
         2.0 * (double)c / iters, 2.0 * (double)c / iters / (warps / 4));
}

int main() {
  unsigned* sink; long long* cyc;
  cudaMalloc(&sink, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int w : {4, 8, 16}) { run<0>(w, 200, sink, cyc); run<1>(w, 200, sink, cyc); run<2>(w, 200, sink, cyc); run<3>(w, 200, sink, cyc); }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
