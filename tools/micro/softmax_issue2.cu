// Micro-benchmark 2: which instruction of the attention exponential loop costs what (2 warps per SMSP, registers
// only, 64 columns per thread, figures scaled to a 128-column row block per warp).
#include <cstdio>
#include "../../pytorch_stable_diffusion_b200/csrc/common.cuh"
using namespace sdb;
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float poly(float x, bool clamp) {
  if (clamp) x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(0.0551716685f, f, 0.2426111400f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// V: 0 FFMA+MUFU+pack | 1 FFMA+MUFU, no pack | 2 FADD+MUFU+pack | 3 MUFU only (x = sv + mb folded: FADD) no pack
//    4 1/4 poly +pack | 5 1/2 poly + pack | 6 1/4 poly no clamp + pack | 7 all poly + pack | 8 FFMA + pack only (no exp)
//    9 1/4 poly, FADD instead of FFMA, no clamp
template <int V>
__global__ void __launch_bounds__(512, 1) k(int iters, float sl2, float mb, unsigned* sink, long long* cyc) {
  extern __shared__ float dummy[];
  float sv[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) sv[i] = -0.01f * (float)((threadIdx.x * 7 + i * 13) & 255);
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    mb += 1e-4f;
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
      float x[4], p[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) x[e] = (V == 2 || V == 3 || V == 9) ? (sv[i + e] - mb) : fmaf(sv[i + e], sl2, -mb);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        bool usep = (V == 4 || V == 6 || V == 9) ? (e == 3) : (V == 5) ? (e & 1) : (V == 7);
        if (V == 8) p[e] = x[e];
        else p[e] = usep ? poly(x[e], !(V == 6 || V == 9)) : ex2a(x[e]);
      }
      if (V == 1 || V == 3) { acc ^= __float_as_uint(p[0]) ^ __float_as_uint(p[1]) ^ __float_as_uint(p[2]) ^ __float_as_uint(p[3]); }
      else { acc ^= pack_bf16x2(p[0], p[1]) ^ pack_bf16x2(p[2], p[3]); }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int V>
void run(const char* what, unsigned* sink, long long* cyc) {
  const int iters = 400, warps = 8;
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int r = 0; r < 2; ++r) { k<V><<<148, warps * 32, 200 * 1024>>>(iters, 1.4427f, -3.0f, sink, cyc); cudaDeviceSynchronize(); }
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("V%d %-44s %7.1f cycles per 128-col row block per warp (2 warps/SMSP)\n", V, what, 2.0 * (double)c / iters / 2);
}
int main() {
  unsigned* sink; long long* cyc;
  cudaMalloc(&sink, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  run<0>("FFMA + MUFU + pack", sink, cyc);
  run<1>("FFMA + MUFU", sink, cyc);
  run<2>("FADD + MUFU + pack", sink, cyc);
  run<3>("FADD + MUFU", sink, cyc);
  run<4>("FFMA + 3/4 MUFU + 1/4 poly + pack", sink, cyc);
  run<5>("FFMA + 1/2 MUFU + 1/2 poly + pack", sink, cyc);
  run<6>("FFMA + 3/4 MUFU + 1/4 poly (no clamp) + pack", sink, cyc);
  run<7>("FFMA + all poly + pack", sink, cyc);
  run<8>("FFMA + pack (no exponential)", sink, cyc);
  run<9>("FADD + 3/4 MUFU + 1/4 poly (no clamp) + pack", sink, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
