// Host-side plumbing shared by every C-ABI entry point: error reporting (return code +
// pq3d_last_error(), never exit() — contrast the reference's CUDA_CHECK_ERRORS,
// modules/third_party/pointnet2/_ext_src/include/cuda_utils.h:29-39) and TMA tensor-map encoding
// through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#define PQ3D_OK 0
#define PQ3D_ERR_INVALID -1
#define PQ3D_ERR_CUDA -2
#define PQ3D_ERR_UNSUPPORTED -3

namespace pq3d {

int set_error(int code, const char* fmt, ...);
const char* last_error();

#define PQ3D_CHECK_ARG(cond, ...) \
  do {                            \
    if (!(cond)) return ::pq3d::set_error(PQ3D_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define PQ3D_CUDA(call)                                                                        \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return ::pq3d::set_error(PQ3D_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                               __FILE__, __LINE__);                                            \
  } while (0)

// bf16 tensor map, 128-byte swizzle, zero OOB fill.  dims/strides innermost-first; strides in BYTES
// for dims 1..rank-1 (dim 0 is contiguous).  Returns PQ3D_OK or an error code.
int make_tmap(CUtensorMap* out, const void* base, int elem_bytes /*2 = bf16, 4 = fp32*/, int rank,
              const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);
inline int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap(out, base, 2, rank, dims, strides_bytes, box);
}

// Programmatic dependent launch (PDL): every kernel in this library starts with
// `griddepcontrol.launch_dependents; griddepcontrol.wait;` after its private set-up, so launching with
// the programmatic-stream-serialization attribute lets kernel N+1's set-up (barrier init, TMEM
// allocation, tensor-map prefetch, launch latency) overlap kernel N's tail.  PQ3D_PDL=0 disables it.
bool pdl_enabled();
// Launch priority of the calling thread's next kernels (cudaLaunchAttributePriority; 0 = the device default = lowest,
// negative = more urgent; recorded in CUDA-graph kernel nodes at capture).  Set through pq3d_set_launch_priority.
int launch_priority();
void set_launch_priority(int p);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (launch_priority() != 0) {
    attr[1].id = cudaLaunchAttributePriority;
    attr[1].val.priority = launch_priority();
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t stream, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[3];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  int n = 1;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (launch_priority() != 0) {
    attr[n].id = cudaLaunchAttributePriority;
    attr[n].val.priority = launch_priority();
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// SM count of the CURRENT device (cached per device index: one process may drive several GPUs).
inline int sm_count() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = dev >= 0 && dev < 64 ? dev : 0;
  if (n[slot] == 0) cudaDeviceGetAttribute(&n[slot], cudaDevAttrMultiProcessorCount, dev);
  return n[slot];
}

}  // namespace pq3d
