// Voxel -> segment pooling: torch_scatter.scatter_mean(f, point2segment, dim=0, dim_size=max_seg)
// per scene and per feature scale, stacked over the batch
// (modules/vision/pcd_mask3d_encoder.py:144-154; torch_scatter is a third-party dependency that is not vendored in
// the reference tree — README pins torch-scatter 2.1.1; its scatter_mean is: sum = scatter_add(src), count =
// scatter_add(ones).clamp(min=1), out = sum / count).
//
// HBM-bound segmented reduction with UNSORTED indices.  The B200 plan has two parts:
//   1. pq3d_segment_csr   index work, once per batch (the five feature scales share it): a STABLE counting sort of
//                         the voxels by (scene, segment) -> perm[] (voxel ids grouped by segment, ascending inside a
//                         segment) + offsets[].  Three small kernels; all integer, deterministic.
//   2. pq3d_segment_mean  one warp per (scene, segment): walks its voxel list in ascending voxel order and adds the
//                         feature rows in fp32 — rows are >= 384 contiguous bytes, every load a full 128-byte line —
//                         then divides by max(count, 1).  Reading each feature row exactly once is the algorithmic
//                         traffic: Nv*C*4 bytes in, B*max_seg*C*(4 [+2]) out.
// Summation order = ascending voxel index = the order torch's CPU scatter_add_/index_add_ uses, so the result is
// BIT-EXACT against the CPU oracle (no atomics on floating point anywhere).
#include "host_common.h"
#include "ptx.cuh"

namespace pq3d {

constexpr int kChunk = 1024;        // voxels per warp-chunk of the counting sort
constexpr int kMaxScenes = 64;

struct SceneTable {
  int32_t voxel_start[kMaxScenes + 1];   // prefix of voxels per scene
  int32_t chunk_start[kMaxScenes + 1];   // prefix of chunks per scene
  int32_t B;
};

__device__ __forceinline__ void locate_chunk(const SceneTable& t, int chunk, int& b, int& v0, int& v1) {
  b = 0;
  while (b + 1 < t.B && chunk >= t.chunk_start[b + 1]) ++b;
  v0 = t.voxel_start[b] + (chunk - t.chunk_start[b]) * kChunk;
  v1 = min(v0 + kChunk, t.voxel_start[b + 1]);
}

// pass A: per-chunk histogram of segment ids -> chunk_hist[chunk][seg]
__global__ void __launch_bounds__(32) seg_hist_kernel(const int64_t* __restrict__ p2s, SceneTable t, int max_seg,
                                                      int32_t* __restrict__ chunk_hist) {
  extern __shared__ int32_t hist[];
  pdl_sync();
  int b, v0, v1;
  locate_chunk(t, blockIdx.x, b, v0, v1);
  for (int s = threadIdx.x; s < max_seg; s += 32) hist[s] = 0;
  __syncwarp();
  for (int v = v0 + threadIdx.x; v < v1; v += 32) {
    const int64_t s = p2s[v];
    if (s >= 0 && s < max_seg) atomicAdd(&hist[static_cast<int>(s)], 1);
  }
  __syncwarp();
  int32_t* out = chunk_hist + static_cast<int64_t>(blockIdx.x) * max_seg;
  for (int s = threadIdx.x; s < max_seg; s += 32) out[s] = hist[s];
}

// pass B1: per (scene, segment) exclusive scan over that scene's chunks (in place) + segment totals
__global__ void seg_scan_chunks_kernel(SceneTable t, int max_seg, int32_t* __restrict__ chunk_hist,
                                       int32_t* __restrict__ counts /* [B*max_seg] */) {
  pdl_sync();
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= t.B * max_seg) return;
  const int b = g / max_seg, s = g % max_seg;
  int run = 0;
  for (int c = t.chunk_start[b]; c < t.chunk_start[b + 1]; ++c) {
    int32_t* p = chunk_hist + static_cast<int64_t>(c) * max_seg + s;
    const int v = *p;
    *p = run;
    run += v;
  }
  counts[g] = run;
}

// pass B2: exclusive scan of the B*max_seg totals -> offsets[0 .. G] (single block)
__global__ void __launch_bounds__(1024) seg_scan_totals_kernel(const int32_t* __restrict__ counts, int G,
                                                               int32_t* __restrict__ offsets) {
  __shared__ int32_t warp_sum[32];
  __shared__ int32_t carry;
  pdl_sync();
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < G; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < G ? counts[i] : 0;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, d);
      if (static_cast<int>(lane_id()) >= d) x += y;
    }
    if (lane_id() == 31) warp_sum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_sum[threadIdx.x];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, d);
        if (static_cast<int>(lane_id()) >= d) w += y;
      }
      warp_sum[threadIdx.x] = w;
    }
    __syncthreads();
    const int before = carry + (threadIdx.x >= 32 ? warp_sum[(threadIdx.x >> 5) - 1] : 0) + (x - v);
    if (i < G) offsets[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[G] = carry;
}

// pass C: stable placement.  One warp per chunk walks its voxels in order, 32 at a time; lanes holding the same
// segment are ranked by lane (match.any), so voxels of a segment keep their ascending order.
__global__ void __launch_bounds__(32) seg_place_kernel(const int64_t* __restrict__ p2s, SceneTable t, int max_seg,
                                                       const int32_t* __restrict__ chunk_hist,
                                                       const int32_t* __restrict__ offsets, int32_t* __restrict__ perm) {
  extern __shared__ int32_t cursor[];
  pdl_sync();
  int b, v0, v1;
  locate_chunk(t, blockIdx.x, b, v0, v1);
  const int32_t* base = chunk_hist + static_cast<int64_t>(blockIdx.x) * max_seg;
  const int32_t* off = offsets + static_cast<int64_t>(b) * max_seg;
  for (int s = threadIdx.x; s < max_seg; s += 32) cursor[s] = off[s] + base[s];
  __syncwarp();
  const unsigned lane = lane_id();
  for (int v = v0; v < v1; v += 32) {
    const int i = v + lane;
    int64_t s64 = i < v1 ? p2s[i] : -1;
    const bool ok = s64 >= 0 && s64 < max_seg;
    const int s = ok ? static_cast<int>(s64) : -1 - static_cast<int>(lane);   // distinct keys for inactive lanes
    const unsigned peers = __match_any_sync(0xffffffffu, s);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    const int leader = __ffs(peers) - 1;
    int start = 0;
    if (ok && static_cast<int>(lane) == leader) {
      start = cursor[s];
      cursor[s] = start + __popc(peers);
    }
    start = __shfl_sync(0xffffffffu, start, leader);
    if (ok) perm[start + rank] = i;
    __syncwarp();
  }
}

// one warp per (scene, segment): fp32 sum in ascending voxel order, / max(count, 1)
__global__ void __launch_bounds__(256) seg_mean_kernel(const float* __restrict__ feat, int64_t ld,
                                                       const int32_t* __restrict__ perm,
                                                       const int32_t* __restrict__ offsets, int G, int C,
                                                       float* __restrict__ out32, int64_t ld32,
                                                       __nv_bfloat16* __restrict__ out16, int64_t ld16, int K16) {
  pdl_sync();
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= G) return;
  const int lane = lane_id();
  const int o0 = offsets[g], n = offsets[g + 1] - o0;
  const float inv_den = static_cast<float>(n > 1 ? n : 1);
  const int C4 = C >> 2;
  for (int c4 = lane; c4 < ((C4 + 31) & ~31); c4 += 32) {
    const bool act = c4 < C4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = 0;
    for (; k + 4 <= n; k += 4) {                       // four rows in flight; the adds stay in voxel order
      const int i0 = perm[o0 + k], i1 = perm[o0 + k + 1], i2 = perm[o0 + k + 2], i3 = perm[o0 + k + 3];
      if (act) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(feat + i0 * ld) + c4);
        const float4 b = __ldg(reinterpret_cast<const float4*>(feat + i1 * ld) + c4);
        const float4 c = __ldg(reinterpret_cast<const float4*>(feat + i2 * ld) + c4);
        const float4 d = __ldg(reinterpret_cast<const float4*>(feat + i3 * ld) + c4);
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
        acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
        acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w;
        acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
      }
    }
    for (; k < n; ++k) {
      const int i0 = perm[o0 + k];
      if (act) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(feat + i0 * ld) + c4);
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
      }
    }
    if (act) {
      acc.x = acc.x / inv_den; acc.y = acc.y / inv_den; acc.z = acc.z / inv_den; acc.w = acc.w / inv_den;   // IEEE division
      if (out32 != nullptr) reinterpret_cast<float4*>(out32 + g * ld32)[c4] = acc;
      if (out16 != nullptr) {
        uint2 u;
        u.x = pack_bf16x2(acc.x, acc.y);
        u.y = pack_bf16x2(acc.z, acc.w);
        reinterpret_cast<uint2*>(out16 + g * ld16)[c4] = u;
      }
    }
  }
  if (out16 != nullptr)                                 // zero the K padding of the GEMM operand
    for (int c = C + lane; c < K16; c += 32) out16[g * ld16 + c] = __float2bfloat16_rn(0.f);
}

}  // namespace pq3d

using namespace pq3d;

static int fill_table(SceneTable& t, const int64_t* voxel_offsets_host, int B) {
  t.B = B;
  t.voxel_start[0] = 0;
  t.chunk_start[0] = 0;
  for (int b = 0; b < B; ++b) {
    const int64_t n = voxel_offsets_host[b + 1] - voxel_offsets_host[b];
    if (n < 0 || voxel_offsets_host[b + 1] > 0x7fffffff) return -1;
    t.voxel_start[b + 1] = static_cast<int32_t>(voxel_offsets_host[b + 1]);
    t.chunk_start[b + 1] = t.chunk_start[b] + static_cast<int32_t>((n + kChunk - 1) / kChunk);
  }
  return t.chunk_start[B];
}

extern "C" int64_t pq3d_segment_csr_workspace_bytes(const int64_t* voxel_offsets_host, int B, int max_seg) {
  if (voxel_offsets_host == nullptr || B < 1 || B > kMaxScenes || max_seg < 1) return -1;
  SceneTable t;
  const int chunks = fill_table(t, voxel_offsets_host, B);
  if (chunks < 0) return -1;
  return (static_cast<int64_t>(chunks > 0 ? chunks : 1) * max_seg + static_cast<int64_t>(B) * max_seg) * 4;
}

// p2s: int64 [Nv_total] segment id of every voxel, scenes concatenated (scene b = voxels
// [voxel_offsets_host[b], voxel_offsets_host[b+1]), a HOST array of B+1 prefix sums starting at 0).  Ids outside
// [0, max_seg) are skipped.  Outputs: perm int32 [Nv_total] (only the first offsets[B*max_seg] entries are written),
// offsets int32 [B*max_seg + 1].  workspace: pq3d_segment_csr_workspace_bytes bytes, caller-owned.
extern "C" int pq3d_segment_csr(const int64_t* p2s, const int64_t* voxel_offsets_host, int B, int max_seg, int32_t* perm,
                                int32_t* offsets, void* workspace, int64_t workspace_bytes, void* stream) {
  PQ3D_CHECK_ARG(p2s && voxel_offsets_host && perm && offsets && workspace, "pq3d_segment_csr: null argument");
  PQ3D_CHECK_ARG(B >= 1 && B <= kMaxScenes && max_seg >= 1, "pq3d_segment_csr: B=%d (1..%d), max_seg=%d", B, kMaxScenes,
                 max_seg);
  PQ3D_CHECK_ARG(voxel_offsets_host[0] == 0, "pq3d_segment_csr: voxel_offsets must start at 0");
  PQ3D_CHECK_ARG(static_cast<int64_t>(max_seg) * 4 <= 200 * 1024,
                 "pq3d_segment_csr: max_seg=%d needs more than 200 KB of shared-memory counters", max_seg);
  SceneTable t;
  const int chunks = fill_table(t, voxel_offsets_host, B);
  PQ3D_CHECK_ARG(chunks >= 0, "pq3d_segment_csr: voxel_offsets must be non-decreasing and below 2^31");
  const int64_t need = pq3d_segment_csr_workspace_bytes(voxel_offsets_host, B, max_seg);
  PQ3D_CHECK_ARG(workspace_bytes >= need, "pq3d_segment_csr: workspace %lld < %lld bytes", (long long)workspace_bytes,
                 (long long)need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int32_t* chunk_hist = reinterpret_cast<int32_t*>(workspace);
  int32_t* counts = chunk_hist + static_cast<int64_t>(chunks > 0 ? chunks : 1) * max_seg;
  const size_t smem = static_cast<size_t>(max_seg) * 4;
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    PQ3D_CUDA(cudaFuncSetAttribute(seg_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PQ3D_CUDA(cudaFuncSetAttribute(seg_place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int G = B * max_seg;
  if (chunks > 0)
    PQ3D_CUDA(launch_kernel(seg_hist_kernel, dim3(chunks), dim3(32), smem, st, p2s, t, max_seg, chunk_hist));
  PQ3D_CUDA(launch_kernel(seg_scan_chunks_kernel, dim3((G + 255) / 256), dim3(256), 0, st, t, max_seg, chunk_hist, counts));
  PQ3D_CUDA(launch_kernel(seg_scan_totals_kernel, dim3(1), dim3(1024), 0, st, static_cast<const int32_t*>(counts), G,
                          offsets));
  if (chunks > 0)
    PQ3D_CUDA(launch_kernel(seg_place_kernel, dim3(chunks), dim3(32), smem, st, p2s, t, max_seg,
                            static_cast<const int32_t*>(chunk_hist), static_cast<const int32_t*>(offsets), perm));
  return PQ3D_OK;
}

// feat: fp32 [Nv_total, C] (leading dimension ld, multiple of 4, 16-byte aligned), perm / offsets from
// pq3d_segment_csr with G = B*max_seg segments.  out32 (optional): fp32 [G, C] (ld32); out16 (optional): bf16
// [G, K16] (ld16) — the next Linear's GEMM operand, columns [C, K16) zero-filled.
extern "C" int pq3d_segment_mean(const float* feat, int64_t ld, const int32_t* perm, const int32_t* offsets, int G, int C,
                                 float* out32, int64_t ld32, void* out16, int64_t ld16, int K16, void* stream) {
  PQ3D_CHECK_ARG(feat && perm && offsets && (out32 || out16), "pq3d_segment_mean: null argument");
  PQ3D_CHECK_ARG(G > 0 && C > 0 && C % 4 == 0 && ld % 4 == 0 && ld >= C && (reinterpret_cast<uintptr_t>(feat) & 15) == 0,
                 "pq3d_segment_mean: C=%d and ld=%lld must be multiples of 4 with a 16-byte aligned base", C, (long long)ld);
  PQ3D_CHECK_ARG(out32 == nullptr || (ld32 % 4 == 0 && ld32 >= C && (reinterpret_cast<uintptr_t>(out32) & 15) == 0),
                 "pq3d_segment_mean: out32 alignment");
  PQ3D_CHECK_ARG(out16 == nullptr || (ld16 % 4 == 0 && K16 >= C && ld16 >= K16 && (reinterpret_cast<uintptr_t>(out16) & 7) == 0),
                 "pq3d_segment_mean: out16 alignment / K16=%d < C=%d", K16, C);
  const int warps_per_block = 8;
  PQ3D_CUDA(launch_kernel(seg_mean_kernel, dim3((G + warps_per_block - 1) / warps_per_block), dim3(32 * warps_per_block), 0,
                          reinterpret_cast<cudaStream_t>(stream), feat, ld, perm, offsets, G, C, out32, ld32,
                          reinterpret_cast<__nv_bfloat16*>(out16), ld16, K16));
  return PQ3D_OK;
}
