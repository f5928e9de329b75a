// Training-step companions that are not part of the backward arithmetic itself: the per-step refresh of the kernels'
// operand copies of the parameters (one launch for every weight, bias and affine of the decoder).
//
// The reference keeps fp32 parameters and lets autocast re-cast each weight to bf16 inside every nn.Linear call of
// every step (torch autocast's weight cache, trainer/build.py:57-75 with mixed precision); here the bf16 operand
// copies live in buffers laid out for the GEMMs (stacked over layers / memories, plus the transposed copies the
// dgrad GEMMs take) and are rewritten once per optimizer step by pq3d_pack_segments.
#include "host_common.h"
#include "ptx.cuh"

namespace pq3d {

// One segment: src fp32 [rows, cols] (dense, ld = cols) ->
//   dst_c [rows, cols] with pitch ld_c (bf16, or fp32 when flags & 1), optional
//   dst_t [cols, rows] with pitch ld_t (bf16), optional.
// Eight int64 words per segment: src, dst_c, dst_t, rows, cols, ld_c, ld_t, flags (1: dst_c is fp32; 2: vector path).
constexpr int kSegWords = 8;

__global__ void __launch_bounds__(256) pack_segments_kernel(const int64_t* __restrict__ segs,
                                                            const int32_t* __restrict__ tile_start, int n_seg) {
  pdl_sync();
  __shared__ float tile[64][65];
  // segment of this block: last s with tile_start[s] <= blockIdx.x
  int lo = 0, hi = n_seg - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tile_start[mid] <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
  }
  const int64_t* sg = segs + static_cast<int64_t>(lo) * kSegWords;
  const float* src = reinterpret_cast<const float*>(sg[0]);
  void* dst_c = reinterpret_cast<void*>(sg[1]);
  __nv_bfloat16* dst_t = reinterpret_cast<__nv_bfloat16*>(sg[2]);
  const int rows = static_cast<int>(sg[3]), cols = static_cast<int>(sg[4]);
  const int64_t ld_c = sg[5], ld_t = sg[6];
  const bool c_fp32 = (sg[7] & 1) != 0;
  const int tiles_c = (cols + 63) / 64;
  const int local = static_cast<int>(blockIdx.x) - tile_start[lo];
  const int r0 = (local / tiles_c) * 64, c0 = (local % tiles_c) * 64;
  if (sg[7] & 2) {
    // vector path (host guarantees rows, cols, pitches multiples of 4 and 16-byte aligned bases): float4 loads,
    // 8-byte bf16x4 stores in both orientations
    const int q = threadIdx.x & 15, rr = threadIdx.x >> 4;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int i = rr + it * 16, r = r0 + i, c = c0 + q * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows && c < cols) {
        v = __ldg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(r) * cols + c));
        if (dst_c != nullptr) {
          if (c_fp32) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst_c) + r * ld_c + c) = v;
          } else {
            uint2 o;
            o.x = pack_bf16x2(v.x, v.y); o.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dst_c) + r * ld_c + c) = o;
          }
        }
      }
      tile[i][q * 4] = v.x; tile[i][q * 4 + 1] = v.y; tile[i][q * 4 + 2] = v.z; tile[i][q * 4 + 3] = v.w;
    }
    if (dst_t == nullptr) return;
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int i = rr + it * 16, c = c0 + i, r = r0 + q * 4;
      if (c < cols && r < rows) {
        uint2 o;
        o.x = pack_bf16x2(tile[q * 4][i], tile[q * 4 + 1][i]);
        o.y = pack_bf16x2(tile[q * 4 + 2][i], tile[q * 4 + 3][i]);
        *reinterpret_cast<uint2*>(dst_t + c * ld_t + r) = o;
      }
    }
    return;
  }
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 64; i += 8) {
    const int r = r0 + i;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int c = c0 + tx + jj * 32;
      float v = 0.f;
      if (r < rows && c < cols) {
        v = __ldg(src + static_cast<int64_t>(r) * cols + c);
        if (dst_c != nullptr) {
          if (c_fp32) reinterpret_cast<float*>(dst_c)[r * ld_c + c] = v;
          else reinterpret_cast<__nv_bfloat16*>(dst_c)[r * ld_c + c] = __float2bfloat16_rn(v);
        }
      }
      tile[i][tx + jj * 32] = v;
    }
  }
  if (dst_t == nullptr) return;
  __syncthreads();
  for (int i = ty; i < 64; i += 8) {
    const int c = c0 + i;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int r = r0 + tx + jj * 32;
      if (c < cols && r < rows) dst_t[c * ld_t + r] = __float2bfloat16_rn(tile[tx + jj * 32][i]);
    }
  }
}

}  // namespace pq3d

using namespace pq3d;

extern "C" int pq3d_pack_segments(const int64_t* segs_dev, const int32_t* tile_start_dev, int n_seg, int total_tiles,
                                  void* stream) {
  PQ3D_CHECK_ARG(segs_dev && tile_start_dev && n_seg > 0 && total_tiles > 0, "pq3d_pack_segments: bad argument");
  PQ3D_CUDA(launch_kernel(pack_segments_kernel, dim3(total_tiles), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                          segs_dev, tile_start_dev, n_seg));
  return PQ3D_OK;
}
