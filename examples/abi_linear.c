/*
 * examples/abi_linear.c - the C boundary without Python or torch: a C99 host program that links libsdb200.so and the
 * CUDA runtime only, runs one nn.Linear (out = x . W^T + b, sd/attention.py:12 style) through sdb_gemm_tc and checks
 * it against a host-side double-precision loop on the same bf16-rounded operands.
 *
 *   gcc -std=c99 -O2 -Iinclude -I/usr/local/cuda/include examples/abi_linear.c \
 *       -Lpytorch_stable_diffusion_b200/csrc -lsdb200 -L/usr/local/cuda/lib64 -lcudart -lm \
 *       -Wl,-rpath,$PWD/pytorch_stable_diffusion_b200/csrc -o /tmp/abi_linear && /tmp/abi_linear
 *
 * tests/test_kernels_gpu.py::test_c_host_program builds and runs it on the GPU box.
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sdb200.h"

static uint16_t to_bf16(float f) {           /* round to nearest even */
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float from_bf16(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

int main(void) {
  enum { M = 300, K = 320, N = 96 };        /* ragged row count, K a multiple of 64 */
  uint16_t* x = malloc(sizeof(uint16_t) * M * K);
  uint16_t* w = malloc(sizeof(uint16_t) * N * K);
  float* b = malloc(sizeof(float) * N);
  float* y = malloc(sizeof(float) * M * N);
  uint32_t s = 12345u;
  for (int i = 0; i < M * K; ++i) { s = s * 1664525u + 1013904223u; x[i] = to_bf16((float)(s >> 8) / 8388608.0f - 1.0f); }
  for (int i = 0; i < N * K; ++i) { s = s * 1664525u + 1013904223u; w[i] = to_bf16(((float)(s >> 8) / 8388608.0f - 1.0f) * 0.056f); }
  for (int i = 0; i < N; ++i) b[i] = 0.01f * (float)i;

  if (sdb_abi_version() != SDB_ABI_VERSION || sdb_args_size(0) != (int)sizeof(sdb_gemm_args)) {
    fprintf(stderr, "header / library mismatch: version %d vs %d, sdb_gemm_args %d vs %zu bytes\n", sdb_abi_version(),
            SDB_ABI_VERSION, sdb_args_size(0), sizeof(sdb_gemm_args));
    return 2;
  }
  void *dx, *dw, *db, *dy;
  if (cudaMalloc(&dx, sizeof(uint16_t) * M * K) || cudaMalloc(&dw, sizeof(uint16_t) * N * K) ||
      cudaMalloc(&db, sizeof(float) * N) || cudaMalloc(&dy, sizeof(float) * M * N)) {
    fprintf(stderr, "cudaMalloc failed (no CUDA device?)\n");
    return 3;
  }
  cudaMemcpy(dx, x, sizeof(uint16_t) * M * K, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, w, sizeof(uint16_t) * N * K, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b, sizeof(float) * N, cudaMemcpyHostToDevice);

  sdb_gemm_args a;
  memset(&a, 0, sizeof(a));                   /* zero = "choose" / "off" for every optional field */
  a.kind = SDB_GEMM_LINEAR;
  a.a0 = dx; a.w = dw; a.bias = (const float*)db; a.out = dy;
  a.M = M; a.C0 = K; a.Cout = N;
  a.out_fp32 = 1;
  int rc = sdb_gemm_tc(&a, NULL);             /* NULL = the default stream */
  if (rc != SDB_OK) { fprintf(stderr, "sdb_gemm_tc: %d (%s)\n", rc, sdb_last_error()); return 4; }
  if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "kernel failed\n"); return 5; }
  cudaMemcpy(y, dy, sizeof(float) * M * N, cudaMemcpyDeviceToHost);

  double max_err = 0.0, max_ref = 0.0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = b[n];
      for (int k = 0; k < K; ++k) acc += (double)from_bf16(x[m * K + k]) * (double)from_bf16(w[n * K + k]);
      const double e = fabs(acc - (double)y[m * N + n]);
      if (e > max_err) max_err = e;
      if (fabs(acc) > max_ref) max_ref = fabs(acc);
    }
  unsigned int fault = 0;
  sdb_read_fault(&fault);
  printf("abi_linear: %d x %d x %d, max |err| / max |ref| = %.3e, watchdog fault word = 0x%x, kernels launched = %llu\n",
         M, K, N, max_err / max_ref, fault, sdb_launch_count());
  cudaFree(dx); cudaFree(dw); cudaFree(db); cudaFree(dy);
  free(x); free(w); free(b); free(y);
  return (max_err / max_ref < 1e-5 && fault == 0) ? 0 : 1;
}
