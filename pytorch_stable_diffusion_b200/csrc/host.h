// Host-side helpers shared by the C-ABI entry points: status codes, last-error text,
// lazily resolved cuTensorMapEncodeTiled (no link-time dependency on libcuda, so the library
// loads on a GPU-less box), and tensor-map construction.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define SDB_OK 0
#define SDB_ERR_ARG (-1)
#define SDB_ERR_UNSUPPORTED (-2)
#define SDB_ERR_CUDA (-3)
#define SDB_ERR_DRIVER (-4)

namespace sdb {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

// bf16 tensor map, 128B swizzle, zero OOB fill. dims[0] is the contiguous dimension.
// strides_bytes has rank-1 entries (stride of dims[1..rank-1]).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, const char* what);

// 1 unless SDB_PDL=0: launches carry cudaLaunchAttributeProgrammaticStreamSerialization.
int pdl_enabled();

// Launch with optional cluster dimension and (by default) programmatic dependent launch. Kernels launched
// through here call pdl_wait() before their first global-memory access.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster_x;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: every entry point that needs more than
// 48 KiB of dynamic shared memory configures its kernels once per device (one process or thread per GPU).
struct PerDeviceOnce { bool done[64]; };
// true when the calling entry point still has to configure the current device (and marks it configured)
inline bool first_use_on_device(PerDeviceOnce& o) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (o.done[dev]) return false;
  o.done[dev] = true;
  return true;
}
template <typename K>
inline int set_max_smem(K kernel, int bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(MaxDynamicSharedMemorySize = %d): %s", what, bytes, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return SDB_ERR_CUDA;
  }
  return SDB_OK;
}

// General form: elem_bytes 2 (16-bit) or 4 (fp32); swizzle_bytes 0 / 32 / 64 / 128.
int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int swizzle_bytes, int rank,
              const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, const char* what);

}  // namespace sdb
