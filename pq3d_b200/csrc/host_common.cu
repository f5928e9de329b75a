#include "host_common.h"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace pq3d {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error() { return g_err; }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("PQ3D_PDL");
    on = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return on == 1;
}

static thread_local int g_priority = 0;
int launch_priority() { return g_priority; }
void set_launch_priority(int p) { g_priority = p; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error(PQ3D_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(PQ3D_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu] strides [%llu,%llu] "
                     "box [%u,%u,%u] base %p",
                     (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                     (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
                     (unsigned long long)(rank > 2 ? strides_bytes[1] : 0), box[0], rank > 1 ? box[1] : 0,
                     rank > 2 ? box[2] : 0, base);
  }
  return PQ3D_OK;
}

}  // namespace pq3d

extern "C" const char* pq3d_last_error(void) { return pq3d::last_error(); }
extern "C" int pq3d_abi_version(void) { return 1; }
extern "C" int pq3d_set_launch_priority(int priority) {
  int least = 0, greatest = 0;
  if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) { least = 0; greatest = 0; }
  if (priority < greatest) priority = greatest;        // numerically lower = more urgent
  if (priority > least) priority = least;
  pq3d::set_launch_priority(priority);
  return PQ3D_OK;
}
