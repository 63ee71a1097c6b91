// Micro-benchmark: is the ~88-cycle cost of a small-N tcgen05.mma (M=128, K=16) a THROUGHPUT floor or the latency of a
// dependent accumulation chain? Issues the same number of MMAs into 1, 2 or 4 different TMEM accumulators in
// round-robin order (the tensor pipe sees the same work; only the dependency between consecutive instructions
// changes), operands from shared memory (SS) or A from TMEM (TS, as attention's P.V).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_chain mma_chain.cu   Run on a B200.
#include <cstdio>
#include "../../pytorch_stable_diffusion_b200/csrc/common.cuh"
using namespace sdb;

__global__ void __launch_bounds__(128, 1) mma_chain_kernel(int n, int iters, int nacc, int ts, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)n);
    const uint64_t a_desc = make_kmajor_sw128_desc(smem_u32(smem));
    const uint64_t b_desc = make_kmajor_sw128_desc(smem_u32(smem) + 16384);
    const uint32_t stride = 64;            // accumulators at TMEM columns 0, 64, 128, 192 (n <= 64)
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t d = tmem + (uint32_t)((k % nacc) * stride);
        if (ts) mma_ts(d, tmem + 256 + 8 * (k & 7), b_desc + (uint64_t)(2 * (k & 3)), idesc, 1u);
        else mma_ss(d, a_desc + (uint64_t)(2 * (k & 3)), b_desc + (uint64_t)(2 * (k & 3)), idesc, 1u);
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
  cudaFuncSetAttribute(mma_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int ns[] = {16, 48, 64};
  for (int ts = 0; ts < 2; ++ts)
    for (int n : ns)
      for (int nacc = 1; nacc <= 4; nacc *= 2) {
        const int iters = 250;
        mma_chain_kernel<<<1, 128, 60 * 1024>>>(n, iters, nacc, ts, d);
        cudaDeviceSynchronize();
        mma_chain_kernel<<<1, 128, 60 * 1024>>>(n, iters, nacc, ts, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%s N=%2d accumulators=%d  issue %.1f clk/MMA  complete %.1f clk/MMA  (pipe floor N/2 = %d)  %s\n",
               ts ? "TS" : "SS", n, nacc, (double)h[0] / (iters * 8), (double)h[1] / (iters * 8), n / 2, cudaGetErrorString(e));
      }
  return 0;
}
