// Host-side shared state of libttvdm_sm100.so: error text, launch counter, driver entry points.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/ttvdm.h"

namespace ttvdm {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;
extern int g_num_sms;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// cuTensorMapEncodeTiled resolved through the runtime (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();
int ensure_init();
void reset_init();

// bf16 tensor map, 128-byte swizzle, zero OOB fill. dims/strides innermost first; strides in BYTES for dims 1..rank-1.
int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, bool swizzle128 = true, const uint32_t* elem_strides = nullptr);
// cross-attention through the tcgen05 flash kernel (attn_flash.cu); returns -1 if the call's shape is not covered
int launch_attn_cross_tc(const ttvdm_xattn_params* p, cudaStream_t stream);

#define TTVDM_CHECK_LAUNCH(name)                                                                  \
  do {                                                                                            \
    cudaError_t e__ = cudaGetLastError();                                                         \
    if (e__ != cudaSuccess) return ::ttvdm::fail(TTVDM_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__)); \
    ::ttvdm::g_launches.fetch_add(1, std::memory_order_relaxed);                                  \
  } while (0)

}  // namespace ttvdm
