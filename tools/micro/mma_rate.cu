// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) as a function of N.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu   Run on a B200.
#include <cstdio>
#include "../../pytorch_stable_diffusion_b200/csrc/common.cuh"
using namespace sdb;

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n, int iters, int kper, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t ring[8];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&done_bar, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&ring[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)n);
    const uint64_t a_desc = make_kmajor_sw128_desc(smem_u32(smem));
    const uint64_t b_desc = make_kmajor_sw128_desc(smem_u32(smem) + 16384);
    const long long t0 = clock64();
    int s = 0;
    for (int i = 0; i < iters; ++i) {
      if (mode & 2) {                       // a wait that always passes + the fence the real loop has
        mbar_wait(&done_bar, 1, 98);
        tc_fence_after();
      }
      if (mode & 4) {                       // A operand from TMEM (columns 256..), as the P.V product of attention
        for (int k = 0; k < kper; ++k) mma_ts(tmem, tmem + 256 + 8 * (k & 7), b_desc + (uint64_t)(2 * (k & 3)), idesc, 1u);
      } else {
        for (int k = 0; k < kper; ++k) mma_ss(tmem, a_desc + (uint64_t)(2 * (k & 3)), b_desc + (uint64_t)(2 * (k & 3)), idesc, 1u);
      }
      if (mode & 1) {                       // per-k-block commit to a ring barrier (nobody waits on it)
        tc_commit(&ring[s]);
        if (++s == 8) s = 0;
      }
    }
    tc_commit(&bar);
    const long long t1 = clock64();
    mbar_wait(&bar, 0, 99);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int ns[] = {48, 64, 96, 128, 160, 256};
  for (int mode = 0; mode < 8; mode += 4)
  for (int n : ns) {
    const int iters = 250, kper = 8;
    mma_rate_kernel<<<1, 128, 60 * 1024>>>(n, iters, kper, mode, d);
    cudaDeviceSynchronize();
    mma_rate_kernel<<<1, 128, 60 * 1024>>>(n, iters, kper, mode, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("mode=%d (commit/kblock=%d wait+fence=%d) N=%3d  issue %.1f clk/MMA   complete %.1f clk/MMA   (ideal %d)  %s\n", mode, mode & 1, (mode >> 1) & 1, n, (double)h[0] / (iters * kper),
           (double)h[1] / (iters * kper), n / 2, cudaGetErrorString(e));
  }
  return 0;
}
