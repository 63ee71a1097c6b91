// Micro-benchmark: TMEM -> register bandwidth (tcgen05.ld 32x32b.x32) with 4 or 8 warps, and the cost of
// a softmax-like pass (128 fp32 per row: max + FFMA + EX2 + pack) with and without the TMEM traffic.
#include <cstdio>
#include "../../pytorch_stable_diffusion_b200/csrc/common.cuh"
using namespace sdb;

__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float pz = fmaf(f, 0.05508868f, 0.24260405f);
  pz = fmaf(pz, f, 0.69327624f);
  pz = fmaf(pz, f, 0.99992894f);
  return __int_as_float(__float_as_int(pz) + (__float_as_int(t) << 23));
}
template <int POLY_OF_8>
__device__ __forceinline__ void exp_pass(const uint32_t* v, float m, uint32_t* pk) {
#pragma unroll
  for (int i = 0; i < 128; i += 2) {
    const float a0 = fmaf(__uint_as_float(v[i]), 0.1f, -m), a1 = fmaf(__uint_as_float(v[i + 1]), 0.1f, -m);
    const bool poly = ((i >> 1) & 7) < POLY_OF_8;
    float p0, p1;
    if (poly) { p0 = ex2_poly(a0); p1 = ex2_poly(a1); }
    else {
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(a0));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(a1));
    }
    pk[i >> 1] = pack_bf16x2(p0, p1);
  }
}
__global__ void __launch_bounds__(256, 1) tmem_rate_kernel(int iters, int mode, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[128];
    if (mode & 1) {
      tmem_ld32(addr + 0, v + 0);
      tmem_ld32(addr + 32, v + 32);
      tmem_ld32(addr + 64, v + 64);
      tmem_ld32(addr + 96, v + 96);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < 128; ++i) v[i] = __float_as_uint(acc + (float)i);
    }
    if (mode & 2) {   // softmax-like math
      float m = -1e30f;
      if (!(mode & 64)) {
#pragma unroll
        for (int i = 0; i < 128; ++i) m = fmaxf(m, __uint_as_float(v[i]));
      } else m = acc;
      uint32_t pk[64];
      const int pf = (mode >> 3) & 7;
      if (pf == 0) exp_pass<0>(v, m, pk);
      else if (pf == 1) exp_pass<1>(v, m, pk);
      else if (pf == 2) exp_pass<2>(v, m, pk);
      else if (pf == 3) exp_pass<3>(v, m, pk);
      else if (pf == 4) exp_pass<4>(v, m, pk);
      else exp_pass<8>(v, m, pk);
      if (mode & 4) {
        tmem_st32(addr + 0, pk + 0);
        tmem_st32(addr + 32, pk + 32);
        tmem_st_wait();
      }
#pragma unroll
      for (int i = 0; i < 64; ++i) acc += __uint_as_float(pk[i]) * 1e-30f;
    } else {
#pragma unroll
      for (int i = 0; i < 128; ++i) acc += __uint_as_float(v[i]) * 1e-30f;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; float* sink;
  cudaMalloc(&d, 16); cudaMalloc(&sink, 4096);
  for (int threads = 128; threads <= 256; threads += 128)
    for (int mode : {2 + 64, 2 + 64 + 8, 2 + 64 + 16, 2 + 64 + 24, 2 + 64 + 32, 2 + 64 + 40, 2, 2 + 16}) {
      const int iters = 2000;
      tmem_rate_kernel<<<1, threads>>>(iters, mode, d, sink);
      cudaDeviceSynchronize();
      tmem_rate_kernel<<<1, threads>>>(iters, mode, d, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h;
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("%d warps  exp pass, poly %d/8, max %s: %8.1f clk per 128x128 block   %s\n", threads / 32, (mode >> 3) & 7, (mode & 64) ? "off" : "on ", (double)h / iters, cudaGetErrorString(e));
    }
  return 0;
}
