// Host-side runtime glue for the C-ABI library (error reporting, TMA descriptor encoding).
#include "host.h"
#include <stdarg.h>
#include <atomic>
#include <stdlib.h>
#include "common.cuh"
#include "../../include/sdb200.h"

namespace sdb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};   // entry points may be called from one thread per GPU

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return SDB_ERR_CUDA;
  }
  return SDB_OK;
}

int pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* ev = getenv("SDB_PDL");
    on = (ev && ev[0] == '0') ? 0 : 1;
  }
  return on;
}

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
    else (void)cudaGetLastError();
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, const char* what) {
  return make_tmap(out, base, 2, 128, rank, dims, strides_bytes, box, what);
}

int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int swizzle_bytes, int rank,
              const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, const char* what) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) {
    set_error("%s: cuTensorMapEncodeTiled unavailable (no CUDA driver)", what);
    return SDB_ERR_DRIVER;
  }
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) {
    set_error("%s: base pointer not 16-byte aligned", what);
    return SDB_ERR_ARG;
  }
  for (int i = 0; i + 1 < rank; ++i) {
    if (gs[i] % 16 != 0) {
      set_error("%s: stride %d = %llu bytes is not a multiple of 16", what, i,
                (unsigned long long)gs[i]);
      return SDB_ERR_ARG;
    }
  }
  const CUtensorMapDataType dt = (elem_bytes == 4) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = (swizzle_bytes == 128) ? CU_TENSOR_MAP_SWIZZLE_128B
                              : (swizzle_bytes == 64) ? CU_TENSOR_MAP_SWIZZLE_64B
                              : (swizzle_bytes == 32) ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base),
                   gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,%llu,%llu,%llu] "
              "box=[%u,%u,%u,%u,%u]",
              what, (int)r, rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
              (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0),
              (unsigned long long)(rank > 4 ? gd[4] : 0), bx[0], rank > 1 ? bx[1] : 0,
              rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0);
    return SDB_ERR_DRIVER;
  }
  return SDB_OK;
}

}  // namespace sdb

extern "C" {

const char* sdb_last_error(void) { return sdb::g_err; }

int sdb_abi_version(void) { return SDB_ABI_VERSION; }

// Kernels launched by this library in this process so far (launches recorded into a CUDA graph
// count once, at capture).
unsigned long long sdb_launch_count(void) { return sdb::g_launches.load(std::memory_order_relaxed); }

// sizeof() of the argument structs as THIS build sees them: a binding compares it with its own struct size
// before the first call (a short struct would make the entry point read past its end).
int sdb_args_size(int which) {
  if (which == 0) return (int)sizeof(sdb_gemm_args);
  if (which == 1) return (int)sizeof(sdb_attn_args);
  return SDB_ERR_ARG;
}

// Reads and clears the device fault word (mbarrier watchdog). Synchronises the device.
int sdb_read_fault(unsigned int* out) {
  unsigned int v = 0;
  cudaError_t e = cudaMemcpyFromSymbol(&v, sdb::g_sdb_fault, sizeof(v));
  if (e != cudaSuccess) {
    sdb::set_error("sdb_read_fault: %s", cudaGetErrorString(e));
    return SDB_ERR_CUDA;
  }
  if (v != 0) {
    unsigned int z = 0;
    cudaMemcpyToSymbol(sdb::g_sdb_fault, &z, sizeof(z));
  }
  *out = v;
  return SDB_OK;
}

}  // extern "C"
