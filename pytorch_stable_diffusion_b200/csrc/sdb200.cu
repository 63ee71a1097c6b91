// Single translation unit of libsdb200.so (keeps the device-global watchdog word and the host
// helpers in one object without relocatable device code).
#include "host.cu"
#include "gemm_tc.cu"
#include "attn_tc.cu"
#include "norm.cu"
#include "elementwise.cu"
