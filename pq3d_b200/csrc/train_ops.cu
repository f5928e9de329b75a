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

// ------------------------------------------------------------------------------------------------
// add_layernorm_train: the training-mode residual block tail,
//   out = sum_g w[b,g] * LN_g(residual + dropout(y_g))       (post-norm: query_encoder.py:304-305, 386-387, 449-450)
// with w[b,g] = 1/G (plain mean over the parallel memories, :153) or the memory-dropout weights keep[b,g] / #kept[b]
// (:145-152).  dropout(y) = y * keep / (1 - p), keep from the counter RNG at element (g*R + row)*D + col.
// Same outputs as pq3d_add_layernorm.  One warp per row.
// ------------------------------------------------------------------------------------------------
constexpr int kLnTrainMaxGroups = 4;

template <int NV>
__global__ void add_layernorm_train_kernel(const float* __restrict__ y, int64_t y_group_stride,
                                           const float* __restrict__ residual, const float* __restrict__ gamma,
                                           const float* __restrict__ beta, int G, float eps, int R,
                                           const float* __restrict__ pos, float* __restrict__ out_f32,
                                           __nv_bfloat16* __restrict__ out_bf16, __nv_bfloat16* __restrict__ out_pos_bf16,
                                           uint32_t drop_thresh, float drop_scale, const uint32_t* __restrict__ seed,
                                           uint32_t site, const float* __restrict__ row_w, int rows_per_scene) {
  pdl_sync();
  constexpr int D = NV * 128;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const int lane = threadIdx.x & 31;
  const int64_t base = static_cast<int64_t>(row) * D;
  const uint32_t key = drop_thresh != 0 ? drop_key(__ldg(seed), site) : 0u;
  float4 res[NV], acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    acc[i] = make_float4(0, 0, 0, 0);
    res[i] = residual != nullptr ? __ldg(reinterpret_cast<const float4*>(residual + base) + i * 32 + lane)
                                 : make_float4(0, 0, 0, 0);
  }
  constexpr float inv_d = 1.f / static_cast<float>(D);
  for (int g = 0; g < G; ++g) {
    const float w = row_w != nullptr ? __ldg(row_w + (row / rows_per_scene) * G + g) : 1.f / static_cast<float>(G);
    float4 x[NV];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 v = y != nullptr ? __ldg(reinterpret_cast<const float4*>(y + g * y_group_stride + base) + i * 32 + lane)
                              : make_float4(0, 0, 0, 0);
      if (drop_thresh != 0) {
        const uint32_t e = static_cast<uint32_t>((static_cast<int64_t>(g) * R + row) * D + (i * 32 + lane) * 4);
        v.x = drop_keep(key, e, drop_thresh) ? v.x * drop_scale : 0.f;
        v.y = drop_keep(key, e + 1, drop_thresh) ? v.y * drop_scale : 0.f;
        v.z = drop_keep(key, e + 2, drop_thresh) ? v.z * drop_scale : 0.f;
        v.w = drop_keep(key, e + 3, drop_thresh) ? v.w * drop_scale : 0.f;
      }
      x[i] = make_float4(res[i].x + v.x, res[i].y + v.y, res[i].z + v.z, res[i].w + v.w);
      sum += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = x[i].x - mean, b = x[i].y - mean, c = x[i].z - mean, d = x[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * inv_d + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + g * D) + i * 32 + lane);
      const float4 be = __ldg(reinterpret_cast<const float4*>(beta + g * D) + i * 32 + lane);
      acc[i].x += w * ((x[i].x - mean) * rstd * ga.x + be.x);
      acc[i].y += w * ((x[i].y - mean) * rstd * ga.y + be.y);
      acc[i].z += w * ((x[i].z - mean) * rstd * ga.z + be.z);
      acc[i].w += w * ((x[i].w - mean) * rstd * ga.w + be.w);
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 o = acc[i];
    const int64_t off = base + (i * 32 + lane) * 4;
    if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + off) = o;
    if (out_bf16 != nullptr)
      *reinterpret_cast<uint2*>(out_bf16 + off) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    if (out_pos_bf16 != nullptr) {
      const float4 pp = __ldg(reinterpret_cast<const float4*>(pos + base) + i * 32 + lane);
      *reinterpret_cast<uint2*>(out_pos_bf16 + off) =
          make_uint2(pack_bf16x2(o.x + pp.x, o.y + pp.y), pack_bf16x2(o.z + pp.z, o.w + pp.w));
    }
  }
}

// x = dropout(x) in place on a bf16 array (the FFN's hidden activations, query_encoder.py:384): element idx kept with
// probability 1 - p and scaled by 1 / (1 - p).  n multiple of 8.
__global__ void dropout_bf16_kernel(__nv_bfloat16* __restrict__ x, int64_t n8, uint32_t drop_thresh, float drop_scale,
                                    const uint32_t* __restrict__ seed, uint32_t site) {
  pdl_sync();
  const uint32_t key = drop_key(__ldg(seed), site);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    uint4 u = reinterpret_cast<uint4*>(x)[i];
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
    const uint32_t e = static_cast<uint32_t>(i * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = drop_keep(key, e + 2 * j, drop_thresh) ? __bfloat162float(h[j].x) * drop_scale : 0.f;
      const float b = drop_keep(key, e + 2 * j + 1, drop_thresh) ? __bfloat162float(h[j].y) * drop_scale : 0.f;
      h[j] = __floats2bfloat162_rn(a, b);
    }
    reinterpret_cast<uint4*>(x)[i] = u;
  }
}

// Backward of pq3d_mask_head_finalize (modules/heads/mask_head.py:36-40): mask_logits = raw / (sum_m valid_m + 1e-8),
// rows of padded segments overwritten by a constant -> d_raw = d_logits / (cnt + 1e-8), 0 on padded segments.  Emits the
// bf16 operand of the two products that follow (d_q = d_raw^T k, d_k = d_raw q): [B*S, Np] with zero pad columns.
// masks: uint8 [n_mem + 1][B*S], memory validity masks first (1 = ignore), the segment padding mask last.
__global__ void mask_head_finalize_bwd_kernel(const float* __restrict__ d_logits, const uint8_t* __restrict__ masks,
                                              int n_mem, __nv_bfloat16* __restrict__ d_raw, int64_t rows, int N, int Np) {
  pdl_sync();
  const int64_t total = rows * Np;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i % Np);
    const int64_t row = i / Np;
    float v = 0.f;
    if (n < N && masks[n_mem * rows + row] == 0) {
      int cnt = 0;
      for (int m = 0; m < n_mem; ++m) cnt += (masks[m * rows + row] == 0) ? 1 : 0;
      v = __fdiv_rn(d_logits[row * N + n], static_cast<float>(cnt) + 1e-8f);
    }
    d_raw[i] = __float2bfloat16_rn(v);
  }
}

// Backward of pq3d_gate_mix (structure 'gate', query_encoder.py:166-170): out = (1 - g) q + g u, g = sigmoid(gl):
//   d_u = d g,  d_q = d (1 - g),  d_gl = d (u - q) g (1 - g)   (fp32, plus the bf16 copy of d_gl the dgrad GEMM takes)
__global__ void gate_mix_bwd_kernel(const float* __restrict__ gl, const float* __restrict__ q, const float* __restrict__ u,
                                    const float* __restrict__ d_out, float* __restrict__ d_gl,
                                    __nv_bfloat16* __restrict__ d_gl16, float* __restrict__ d_u, float* __restrict__ d_q,
                                    int64_t n) {
  pdl_sync();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float g = __fdiv_rn(1.f, 1.f + expf(-gl[i]));
    const float d = d_out[i];
    const float dg = d * (u[i] - q[i]) * g * (1.f - g);
    d_gl[i] = dg;
    d_gl16[i] = __float2bfloat16_rn(dg);
    d_u[i] = d * g;
    d_q[i] = d * (1.f - g);
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

extern "C" int pq3d_add_layernorm_train(const float* y, int64_t y_group_stride, const float* residual, const float* gamma,
                                        const float* beta, int G, float eps, int R, int D, const float* pos,
                                        float* out_f32, void* out_bf16, void* out_pos_bf16, float drop_p,
                                        const uint32_t* seed_dev, uint32_t site, const float* row_w, int rows_per_scene,
                                        void* stream) {
  PQ3D_CHECK_ARG((y || residual) && gamma && beta, "pq3d_add_layernorm_train: null argument");
  PQ3D_CHECK_ARG(G >= 1 && G <= kLnTrainMaxGroups && R > 0 && D % 128 == 0 && D <= 1024,
                 "pq3d_add_layernorm_train: bad shape G=%d R=%d D=%d", G, R, D);
  PQ3D_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || seed_dev != nullptr),
                 "pq3d_add_layernorm_train: dropout needs p in [0,1) and a device seed");
  PQ3D_CHECK_ARG(row_w == nullptr || (rows_per_scene > 0 && R % rows_per_scene == 0),
                 "pq3d_add_layernorm_train: row weights need rows_per_scene dividing R");
  PQ3D_CHECK_ARG(out_pos_bf16 == nullptr || pos != nullptr, "pq3d_add_layernorm_train: out_pos_bf16 needs pos");
  PQ3D_CHECK_ARG(static_cast<int64_t>(G) * R * D < (int64_t(1) << 32), "pq3d_add_layernorm_train: too many elements");
  const int warps = 2;
  const dim3 grid((R + warps - 1) / warps), block(warps * 32);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const uint32_t thresh = drop_p > 0.f ? drop_threshold(drop_p) : 0u;
  const float scale = 1.f / (1.f - drop_p);
  cudaError_t err = cudaSuccess;
#define PQ3D_LNT_CASE(NV)                                                                                            \
  case NV:                                                                                                           \
    err = launch_kernel(add_layernorm_train_kernel<NV>, grid, block, 0, st, y, y_group_stride, residual, gamma, beta, \
                        G, eps, R, pos, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16),                         \
                        reinterpret_cast<__nv_bfloat16*>(out_pos_bf16), thresh, scale, seed_dev, site, row_w,        \
                        rows_per_scene);                                                                             \
    break;
  switch (D / 128) {
    PQ3D_LNT_CASE(1) PQ3D_LNT_CASE(2) PQ3D_LNT_CASE(3) PQ3D_LNT_CASE(4) PQ3D_LNT_CASE(5) PQ3D_LNT_CASE(6)
    PQ3D_LNT_CASE(7) PQ3D_LNT_CASE(8)
  }
#undef PQ3D_LNT_CASE
  PQ3D_CUDA(err);
  return PQ3D_OK;
}

extern "C" int pq3d_dropout_bf16(void* x, int64_t n, float drop_p, const uint32_t* seed_dev, uint32_t site,
                                 void* stream) {
  PQ3D_CHECK_ARG(x && seed_dev && n > 0 && n % 8 == 0 && n < (int64_t(1) << 32) && drop_p > 0.f && drop_p < 1.f &&
                     (reinterpret_cast<uintptr_t>(x) & 15) == 0,
                 "pq3d_dropout_bf16: bad argument (n multiple of 8, 0 < p < 1, 16-byte aligned)");
  int64_t blocks = (n / 8 + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  PQ3D_CUDA(launch_kernel(dropout_bf16_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0,
                          reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<__nv_bfloat16*>(x), n / 8,
                          drop_threshold(drop_p), 1.f / (1.f - drop_p), seed_dev, site));
  return PQ3D_OK;
}

extern "C" int pq3d_mask_head_finalize_bwd(const float* d_logits, const uint8_t* masks, int n_mem, void* d_raw_bf16, int B,
                                           int S, int N, int Np, void* stream) {
  PQ3D_CHECK_ARG(d_logits && masks && d_raw_bf16 && n_mem >= 1 && B > 0 && S > 0 && N > 0 && Np >= N,
                 "pq3d_mask_head_finalize_bwd: bad argument");
  const int64_t rows = static_cast<int64_t>(B) * S;
  int64_t blocks = (rows * Np + 255) / 256;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  PQ3D_CUDA(launch_kernel(mask_head_finalize_bwd_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0,
                          reinterpret_cast<cudaStream_t>(stream), d_logits, masks, n_mem,
                          reinterpret_cast<__nv_bfloat16*>(d_raw_bf16), rows, N, Np));
  return PQ3D_OK;
}

extern "C" int pq3d_gate_mix_bwd(const float* gate_logits, const float* query, const float* update, const float* d_out,
                                 float* d_gate_logits, void* d_gate_logits_bf16, float* d_update, float* d_query, int64_t n,
                                 void* stream) {
  PQ3D_CHECK_ARG(gate_logits && query && update && d_out && d_gate_logits && d_gate_logits_bf16 && d_update && d_query &&
                     n > 0, "pq3d_gate_mix_bwd: bad argument");
  int64_t blocks = (n + 255) / 256;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  PQ3D_CUDA(launch_kernel(gate_mix_bwd_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0,
                          reinterpret_cast<cudaStream_t>(stream), gate_logits, query, update, d_out, d_gate_logits,
                          reinterpret_cast<__nv_bfloat16*>(d_gate_logits_bf16), d_update, d_query, n));
  return PQ3D_OK;
}
