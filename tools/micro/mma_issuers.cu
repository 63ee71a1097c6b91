// Micro-benchmark: does the ~90-120 cycle cost of a small tcgen05.mma belong to the ISSUING THREAD or to the tensor
// pipe? W warps (one elected thread each) issue independent MMA streams (own accumulator, own operands) at the same
// time; if the aggregate rate scales with W the cost is per issuing thread and a kernel with two tiles should give each
// tile its own issuer warp. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_issuers mma_issuers.cu
#include <cstdio>
#include "../../pytorch_stable_diffusion_b200/csrc/common.cuh"
using namespace sdb;

__global__ void __launch_bounds__(256, 1) mma_issuers_kernel(int n, int iters, int nwarps, int ts, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp < nwarps && lane == 0) {
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)n);
    const uint64_t a_desc = make_kmajor_sw128_desc(smem_u32(smem));
    const uint64_t b_desc = make_kmajor_sw128_desc(smem_u32(smem) + 16384);
    const uint32_t d = tmem + (uint32_t)(warp * 64);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (ts) mma_ts(d, tmem + 256 + 8 * (k & 7), b_desc + (uint64_t)(2 * (k & 3)), idesc, 1u);
        else mma_ss(d, a_desc + (uint64_t)(2 * (k & 3)), b_desc + (uint64_t)(2 * (k & 3)), idesc, 1u);
      }
    }
    tc_commit(&bar[warp]);
    t1 = clock64();
    mbar_wait(&bar[warp], 0, 99);
    t2 = clock64();
    out[warp * 2] = t1 - t0;
    out[warp * 2 + 1] = t2 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(mma_issuers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int ns[] = {48, 64, 128, 256};
  for (int ts = 0; ts < 2; ++ts)
    for (int n : ns)
      for (int w = 1; w <= 4; w *= 2) {
        if (n == 256 && w > 1) continue;      // accumulators are 64 columns apart
        if (n == 128 && w > 2) continue;
        const int iters = 250;
        for (int rep = 0; rep < 2; ++rep) {
          cudaMemset(d, 0, 64);
          mma_issuers_kernel<<<1, 256, 60 * 1024>>>(n, iters, w, ts, d);
          cudaDeviceSynchronize();
        }
        long long h[8];
        cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        long long worst = 0;
        for (int i = 0; i < w; ++i) worst = h[2 * i + 1] > worst ? h[2 * i + 1] : worst;
        printf("%s N=%3d issuing warps=%d  per-warp issue %.1f clk/MMA  aggregate %.1f clk/MMA  (pipe floor N/2 = %d)  %s\n",
               ts ? "TS" : "SS", n, w, (double)h[0] / (iters * 8), (double)worst / (iters * 8 * w), n / 2,
               cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
