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

// General form: elem_bytes 2 (16-bit) or 4 (fp32); swizzle_bytes 0 / 32 / 64 / 128.
int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int swizzle_bytes, int rank,
              const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, const char* what);

}  // namespace sdb
