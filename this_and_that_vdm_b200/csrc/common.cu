#include "common.h"

#include <cstring>
#include <mutex>

namespace ttvdm {

thread_local char g_err[512] = {0};
std::atomic<uint64_t> g_launches{0};
int g_num_sms = 0;
static EncodeTiledFn g_encode = nullptr;
static std::mutex g_init_mutex;
static bool g_inited = false;
static int g_init_status = 0;

EncodeTiledFn encode_tiled_fn() { return g_encode; }

static void do_init() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    g_init_status = fail(TTVDM_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    return;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    g_init_status = fail(TTVDM_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    return;
  }
  if (prop.major != 10) {
    g_init_status = fail(TTVDM_ERR_ARCH, "libttvdm_sm100 needs an sm_100 device, found sm_%d%d", prop.major, prop.minor);
    return;
  }
  g_num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr) {
    g_init_status = fail(TTVDM_ERR_CUDA, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed");
    return;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  g_init_status = 0;
}

int ensure_init() {
  std::lock_guard<std::mutex> lock(g_init_mutex);
  if (!g_inited) {
    do_init();
    g_inited = g_init_status == 0;  // a failed init (no device yet, wrong arch) is retried by the next call
  }
  return g_init_status;
}

void reset_init() {
  std::lock_guard<std::mutex> lock(g_init_mutex);
  g_inited = false;
  g_encode = nullptr;
  g_num_sms = 0;
  g_init_status = 0;
}

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, bool swizzle128, const uint32_t* elem_strides) {
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = elem_strides ? elem_strides[i] : 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(TTVDM_ERR_SHAPE, "tensor map base not 16B aligned");
  for (int i = 0; i + 1 < rank; ++i)
    if (gstr[i] % 16 != 0) return fail(TTVDM_ERR_SHAPE, "tensor map stride %d (%llu B) not a multiple of 16", i,
                                       (unsigned long long)gstr[i]);
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                        bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TTVDM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace ttvdm

extern "C" {

int ttvdm_init(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return ttvdm::fail(TTVDM_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  return ttvdm::ensure_init();
}

int ttvdm_last_error(char* buf, size_t n) {
  if (!buf || n == 0) return TTVDM_ERR_SHAPE;
  strncpy(buf, ttvdm::g_err, n - 1);
  buf[n - 1] = 0;
  return 0;
}

// Drops the process-wide immutable state (driver entry point, SM count). The library owns no device memory, streams or
// events, so there is nothing else to release; the next call (or ttvdm_init) initialises again.
int ttvdm_destroy(void) {
  ttvdm::reset_init();
  return 0;
}

int ttvdm_abi_version(void) { return 4; }

uint64_t ttvdm_launch_count(void) { return ttvdm::g_launches.load(); }
}
