// HBM-bound companions of the tensor-core kernels: memory ingest (fp32 -> bf16, +pos), residual +
// LayerNorm (+ mean over memories, + bf16 operand emission), mask bit-packing and the mask-head
// finalisation.  All are streaming kernels: 16-byte accesses, one warp per 768-wide row.
#include "host_common.h"
#include "ptx.cuh"

namespace pq3d {

// ------------------------------------------------------------------------------------------------
// ingest: x_v = bf16(feat), x_k = bf16(feat + pos); rows s in [S, S_pitch) are zero-filled so that
// masked P (= 0) times padded V stays 0.   CrossAttentionLayer.with_pos_embed
// (modules/grounding/query_encoder.py:285-286,298-300).
// ------------------------------------------------------------------------------------------------
__global__ void ingest_kernel(const float* __restrict__ feat, const float* __restrict__ pos,
                              __nv_bfloat16* __restrict__ xk, __nv_bfloat16* __restrict__ xv, int B, int S,
                              int S_pitch, int D) {
  pdl_sync();
  const int vec_per_row = D / 8;
  const int64_t total = static_cast<int64_t>(B) * S_pitch * vec_per_row;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vec_per_row);
    const int64_t row = i / vec_per_row;
    const int s = static_cast<int>(row % S_pitch);
    const int b = static_cast<int>(row / S_pitch);
    uint4 ok = make_uint4(0, 0, 0, 0), ov = make_uint4(0, 0, 0, 0);
    if (s < S) {
      const int64_t src = (static_cast<int64_t>(b) * S + s) * D + v * 8;
      const float4 f0 = __ldg(reinterpret_cast<const float4*>(feat + src));
      const float4 f1 = __ldg(reinterpret_cast<const float4*>(feat + src) + 1);
      ov.x = pack_bf16x2(f0.x, f0.y); ov.y = pack_bf16x2(f0.z, f0.w);
      ov.z = pack_bf16x2(f1.x, f1.y); ov.w = pack_bf16x2(f1.z, f1.w);
      if (pos != nullptr) {
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos + src));
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos + src) + 1);
        ok.x = pack_bf16x2(f0.x + p0.x, f0.y + p0.y); ok.y = pack_bf16x2(f0.z + p0.z, f0.w + p0.w);
        ok.z = pack_bf16x2(f1.x + p1.x, f1.y + p1.y); ok.w = pack_bf16x2(f1.z + p1.z, f1.w + p1.w);
      }
    }
    const int64_t dst = row * D + v * 8;
    if (xv != nullptr) *reinterpret_cast<uint4*>(xv + dst) = ov;
    if (xk != nullptr) *reinterpret_cast<uint4*>(xk + dst) = (pos != nullptr) ? ok : ov;
  }
}

// Several memories that share ONE positional table (the scene memories: mv / pc / voxel all add fts_pos,
// model/query3d_unified.py:139-155): pos is read once per element instead of once per memory — per (token, feature)
// 4*(n+1) bytes in instead of 8*n, 4*n bytes out.  Memory m's operands live at xk + m*mem_stride, xv + m*mem_stride.
struct IngestMany {
  const float* feat[4];
  int32_t n;
};
__global__ void ingest_many_kernel(IngestMany in, const float* __restrict__ pos, __nv_bfloat16* __restrict__ xk,
                                   __nv_bfloat16* __restrict__ xv, int64_t mem_stride, int B, int S, int S_pitch, int D) {
  pdl_sync();
  const int vec_per_row = D / 8;
  const int64_t total = static_cast<int64_t>(B) * S_pitch * vec_per_row;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vec_per_row);
    const int64_t row = i / vec_per_row;
    const int s = static_cast<int>(row % S_pitch);
    const int b = static_cast<int>(row / S_pitch);
    const int64_t dst = row * D + v * 8;
    if (s >= S) {
      for (int m = 0; m < in.n; ++m) {
        *reinterpret_cast<uint4*>(xv + m * mem_stride + dst) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(xk + m * mem_stride + dst) = make_uint4(0, 0, 0, 0);
      }
      continue;
    }
    const int64_t src = (static_cast<int64_t>(b) * S + s) * D + v * 8;
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos + src));
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos + src) + 1);
    float4 f0[4], f1[4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
      if (m < in.n) {                                   // every load in flight before the first use
        f0[m] = __ldg(reinterpret_cast<const float4*>(in.feat[m] + src));
        f1[m] = __ldg(reinterpret_cast<const float4*>(in.feat[m] + src) + 1);
      }
#pragma unroll
    for (int m = 0; m < 4; ++m)
      if (m < in.n) {
        uint4 ov, ok;
        ov.x = pack_bf16x2(f0[m].x, f0[m].y); ov.y = pack_bf16x2(f0[m].z, f0[m].w);
        ov.z = pack_bf16x2(f1[m].x, f1[m].y); ov.w = pack_bf16x2(f1[m].z, f1[m].w);
        ok.x = pack_bf16x2(f0[m].x + p0.x, f0[m].y + p0.y); ok.y = pack_bf16x2(f0[m].z + p0.z, f0[m].w + p0.w);
        ok.z = pack_bf16x2(f1[m].x + p1.x, f1[m].y + p1.y); ok.w = pack_bf16x2(f1[m].z + p1.z, f1[m].w + p1.w);
        *reinterpret_cast<uint4*>(xv + m * mem_stride + dst) = ov;
        *reinterpret_cast<uint4*>(xk + m * mem_stride + dst) = ok;
      }
  }
}

// ------------------------------------------------------------------------------------------------
// add_layernorm: out = (1/G) * sum_g LN_g(residual + y_g)   (post-norm residual blocks,
// query_encoder.py:304-305,449-450,386-387; the mean over memories is parallel_ca's eval branch,
// :153).  Optionally emits the bf16 operands of the next GEMMs: bf16(out) and bf16(out + pos).
// One warp per row; two-pass variance in fp32.
// ------------------------------------------------------------------------------------------------
constexpr int kLnMaxVec = 8;   // D <= 1024
constexpr int kLnMaxGroups = 4;

// NV = D / 128 float4 chunks per lane.  Every global load (residual, all groups' y, affines of group 0)
// is issued before the first reduction, so a row costs one memory round trip, not one per stage.
template <int NV>
__global__ void add_layernorm_kernel(const float* __restrict__ y, int64_t y_group_stride,
                                     const float* __restrict__ residual, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, int G, float eps, int R,
                                     const float* __restrict__ pos, float* __restrict__ out_f32,
                                     __nv_bfloat16* __restrict__ out_bf16, __nv_bfloat16* __restrict__ out_pos_bf16) {
  pdl_sync();
  constexpr int D = NV * 128;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const int lane = threadIdx.x & 31;
  const int64_t base = static_cast<int64_t>(row) * D;
  float4 res[NV], acc[NV], pp[NV];
  float4 yv[kLnMaxGroups][NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    acc[i] = make_float4(0, 0, 0, 0);
    res[i] = residual != nullptr ? __ldg(reinterpret_cast<const float4*>(residual + base) + i * 32 + lane)
                                 : make_float4(0, 0, 0, 0);
    pp[i] = (pos != nullptr && out_pos_bf16 != nullptr) ? __ldg(reinterpret_cast<const float4*>(pos + base) + i * 32 + lane)
                                                        : make_float4(0, 0, 0, 0);
  }
#pragma unroll
  for (int g = 0; g < kLnMaxGroups; ++g) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      yv[g][i] = (g < G && y != nullptr) ? __ldg(reinterpret_cast<const float4*>(y + g * y_group_stride + base) + i * 32 + lane)
                                         : make_float4(0, 0, 0, 0);
  }
  constexpr float inv_d = 1.f / static_cast<float>(D);
#pragma unroll
  for (int g = 0; g < kLnMaxGroups; ++g) {
    if (g < G) {
      float4 ga[NV], be[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        ga[i] = __ldg(reinterpret_cast<const float4*>(gamma + g * D) + i * 32 + lane);
        be[i] = __ldg(reinterpret_cast<const float4*>(beta + g * D) + i * 32 + lane);
      }
      float4 x[NV];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        x[i] = make_float4(res[i].x + yv[g][i].x, res[i].y + yv[g][i].y, res[i].z + yv[g][i].z, res[i].w + yv[g][i].w);
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
        acc[i].x += (x[i].x - mean) * rstd * ga[i].x + be[i].x;
        acc[i].y += (x[i].y - mean) * rstd * ga[i].y + be[i].y;
        acc[i].z += (x[i].z - mean) * rstd * ga[i].z + be[i].z;
        acc[i].w += (x[i].w - mean) * rstd * ga[i].w + be[i].w;
      }
    }
  }
  const float inv_g = 1.f / static_cast<float>(G);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float4 o = acc[i];
    if (G > 1) { o.x *= inv_g; o.y *= inv_g; o.z *= inv_g; o.w *= inv_g; }
    const int64_t off = base + (i * 32 + lane) * 4;
    if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + off) = o;
    if (out_bf16 != nullptr)
      *reinterpret_cast<uint2*>(out_bf16 + off) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    if (out_pos_bf16 != nullptr)
      *reinterpret_cast<uint2*>(out_pos_bf16 + off) =
          make_uint2(pack_bf16x2(o.x + pp[i].x, o.y + pp[i].y), pack_bf16x2(o.z + pp[i].z, o.w + pp[i].w));
  }
}

// ------------------------------------------------------------------------------------------------
// pack_mask: bool bytes [rows, S] (1 = ignore) -> bits [rows, W], W = 4*ceil(S/128) words; bits past
// S are set.  unmask_full_rows implements `attn_mask[attn_mask.all(-1)] = False`
// (query_encoder.py:83): a query that would see nothing sees everything.  One warp per row.
// ------------------------------------------------------------------------------------------------
__global__ void pack_mask_kernel(const uint8_t* __restrict__ mask, uint32_t* __restrict__ bits, int64_t rows, int S,
                                 int W, int unmask_full_rows, uint8_t* __restrict__ mask_fixed,
                                 int32_t* __restrict__ active_tiles, int64_t rows_per_batch) {
  pdl_sync();
  __shared__ int s_last;
  if (threadIdx.x == 0) s_last = 0;
  // one block per row; warp w packs words w, w + nwarps, ...; the all-masked test is a block-wide AND
  const int64_t row = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint8_t* src = mask + row * S;
  int all_masked = 1;
  if (unmask_full_rows) {
    for (int s = threadIdx.x; s < S; s += blockDim.x) all_masked &= (src[s] != 0);
    all_masked = __syncthreads_and(all_masked);
  } else {
    all_masked = 0;
  }
  int last = 0;   // 1 + index of the last 128-key tile of this row that has a visible key
  for (int w = wid; w < W; w += nw) {
    const int s = w * 32 + lane;
    bool mk = (s < S) ? (src[s] != 0) : true;
    if (all_masked && s < S) mk = false;
    const uint32_t word = __ballot_sync(0xffffffffu, mk);
    if (lane == 0) bits[row * W + w] = word;
    if (mask_fixed != nullptr && s < S) mask_fixed[row * S + s] = mk ? 1 : 0;
    if (word != 0xffffffffu) last = max(last, w / 4 + 1);
  }
  if (active_tiles != nullptr) {
    // trailing fully-masked key tiles (padding of ragged scenes) are skipped by the attention kernel
    if (!unmask_full_rows) __syncthreads();   // s_last initialised (the fix-up path already synchronised)
    if (lane == 0 && last > 0) atomicMax(&s_last, last);
    __syncthreads();
    if (threadIdx.x == 0 && s_last > 0) atomicMax(&active_tiles[row / rows_per_batch], s_last);
  }
}

// ------------------------------------------------------------------------------------------------
// mask_head_finalize (modules/heads/mask_head.py:36-43): raw = sum_m valid_m * (k_m . q_m) arrives
// as [B, S, N] fp32;  mask_logits = raw / (sum_m valid_m + 1e-8);  mask_logits[seg_pad] = -1e6;
// attn_mask[b, n, s] = sigmoid(mask_logits[b, s, n]) < 0.5   (transposed, bool bytes).
// 32x32 tiles through shared memory so both the [S,N] read/write and the [N,S] write coalesce.
// ------------------------------------------------------------------------------------------------
__global__ void mask_head_finalize_kernel(const float* __restrict__ raw, const uint8_t* const* __restrict__ mem_masks,
                                          int n_mem, const uint8_t* __restrict__ seg_masks,
                                          float* __restrict__ mask_logits, uint8_t* __restrict__ attn_mask, int S,
                                          int N) {
  pdl_sync();
  __shared__ uint8_t tile[32][33];
  const int b = blockIdx.z;
  const int s0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int s = s0 + i, n = n0 + threadIdx.x;
    uint8_t am = 0;
    if (s < S && n < N) {
      int cnt = 0;
      for (int m = 0; m < n_mem; ++m) cnt += (mem_masks[m][static_cast<int64_t>(b) * S + s] == 0) ? 1 : 0;
      const int64_t idx = (static_cast<int64_t>(b) * S + s) * N + n;
      float v = __fdiv_rn(raw[idx], static_cast<float>(cnt) + 1e-8f);
      if (seg_masks[static_cast<int64_t>(b) * S + s] != 0) v = -1e6f;
      mask_logits[idx] = v;
      const float sg = __fdiv_rn(1.f, 1.f + expf(-v));
      am = sg < 0.5f ? 1 : 0;
    }
    tile[i][threadIdx.x] = am;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, s = s0 + threadIdx.x;
    if (n < N && s < S) attn_mask[(static_cast<int64_t>(b) * N + n) * S + s] = tile[threadIdx.x][i];
  }
}

// gate mix for structure 'gate' (query_encoder.py:166-170): out = (1 - sigmoid(g)) * q + sigmoid(g) * u
__global__ void gate_mix_kernel(const float* __restrict__ gate_logits, const float* __restrict__ query,
                                const float* __restrict__ update, float* __restrict__ out, int64_t n) {
  pdl_sync();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float g = __fdiv_rn(1.f, 1.f + expf(-gate_logits[i]));
    out[i] = (1.f - g) * query[i] + g * update[i];
  }
}

// plain fp32 -> bf16 cast (+ optional add), for operands that do not pass through a LayerNorm
__global__ void cast_bf16_kernel(const float* __restrict__ x, const float* __restrict__ add,
                                 __nv_bfloat16* __restrict__ out, int64_t n4) {
  pdl_sync();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    if (add != nullptr) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(add) + i);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}


// ------------------------------------------------------------------------------------------------
// fourier_pos: CoordinateEncoder's Fourier features (model/query3d_unified.py:15-27 ->
// modules/third_party/mask3d/position_embedding.py:38-43,127-156): xyz' = (xyz - min) / (max - min),
// proj = 2*pi*xyz' . gauss_B[3, D/2], out = [sin(proj), cos(proj)] emitted as the bf16 GEMM operand
// of feat_proj (the reference computes this part in fp32 with autocast disabled).
// ------------------------------------------------------------------------------------------------
__global__ void fourier_pos_kernel(const float* __restrict__ xyz, int xyz_stride, const float* __restrict__ cmin,
                                   const float* __restrict__ cmax, const float* __restrict__ gauss_B,
                                   __nv_bfloat16* __restrict__ out, int B, int L, int half) {
  pdl_sync();
  const int64_t total = static_cast<int64_t>(B) * L * half;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % half);
    const int64_t pt = i / half;
    const int b = static_cast<int>(pt / L);
    const float* p = xyz + pt * xyz_stride;
    float proj = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float lo = cmin[b * 3 + d], hi = cmax[b * 3 + d];
      float x = __fdiv_rn((p[d] - lo) * 1.0f, hi - lo) + 0.0f;   // shift_scale_points with dst range [0,1]
      x *= 6.283185307179586f;
      proj = fmaf(x, gauss_B[d * half + j], proj);
    }
    float sn, cs;
    sincosf(proj, &sn, &cs);
    out[pt * (2 * half) + j] = __float2bfloat16_rn(sn);
    out[pt * (2 * half) + half + j] = __float2bfloat16_rn(cs);
  }
}

// ------------------------------------------------------------------------------------------------
// pairwise_locs: calc_pairwise_locs(centers, None, 'center', spatial_dist_norm=True, spatial_dim=5)
// (modules/utils.py:38-68).  One block per scene: pass 1 block-reduces the scene's max distance,
// pass 2 writes [d/max, dz/d, d2/d, dy/d2, dx/d2].
// ------------------------------------------------------------------------------------------------
__global__ void pairwise_locs_kernel(const float* __restrict__ centers, int c_stride, float* __restrict__ out, int N,
                                     float eps) {
  pdl_sync();
  __shared__ float red[32];
  __shared__ float s_max;
  const int b = blockIdx.x;
  const float* c = centers + static_cast<int64_t>(b) * N * c_stride;
  float mx = 0.f;
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) {
    const int a = i / N, t = i % N;
    const float dx = c[a * c_stride] - c[t * c_stride], dy = c[a * c_stride + 1] - c[t * c_stride + 1],
                dz = c[a * c_stride + 2] - c[t * c_stride + 2];
    mx = fmaxf(mx, sqrtf(dx * dx + dy * dy + dz * dz + eps));
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) s_max = v;
  }
  __syncthreads();
  const float maxd = s_max;
  float* o = out + static_cast<int64_t>(b) * N * N * 5;
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) {
    const int a = i / N, t = i % N;
    const float dx = c[a * c_stride] - c[t * c_stride], dy = c[a * c_stride + 1] - c[t * c_stride + 1],
                dz = c[a * c_stride + 2] - c[t * c_stride + 2];
    const float d = sqrtf(dx * dx + dy * dy + dz * dz + eps);
    const float d2 = sqrtf(dx * dx + dy * dy + eps);
    o[i * 5 + 0] = __fdiv_rn(d, maxd);
    o[i * 5 + 1] = __fdiv_rn(dz, d);
    o[i * 5 + 2] = __fdiv_rn(d2, d);
    o[i * 5 + 3] = __fdiv_rn(dy, d2);
    o[i * 5 + 4] = __fdiv_rn(dx, d2);
  }
}

}  // namespace pq3d

using namespace pq3d;

static inline int grid_for(int64_t work_items, int threads) {
  int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

extern "C" int pq3d_ingest_memory(const float* feat, const float* pos, void* xk, void* xv, int B, int S, int S_pitch,
                                  int D, void* stream) {
  PQ3D_CHECK_ARG(feat && (xk || xv), "pq3d_ingest_memory: null argument");
  PQ3D_CHECK_ARG(B > 0 && S > 0 && S_pitch >= S && D % 8 == 0, "pq3d_ingest_memory: bad shape B=%d S=%d pitch=%d D=%d",
                 B, S, S_pitch, D);
  PQ3D_CHECK_ARG((reinterpret_cast<uintptr_t>(feat) & 15) == 0 && (reinterpret_cast<uintptr_t>(pos) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(xk) & 15) == 0 && (reinterpret_cast<uintptr_t>(xv) & 15) == 0,
                 "pq3d_ingest_memory: pointers must be 16-byte aligned");
  const int64_t total = static_cast<int64_t>(B) * S_pitch * (D / 8);
  PQ3D_CUDA(launch_kernel(ingest_kernel, dim3(grid_for(total, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                          feat, pos, reinterpret_cast<__nv_bfloat16*>(xk), reinterpret_cast<__nv_bfloat16*>(xv), B, S,
                          S_pitch, D));
  return PQ3D_OK;
}

// n_mem (2..4) memories with the same (B, S, D) shape and ONE shared positional table: feats is a HOST array of device
// pointers; memory m's operands are written at xk + m*mem_stride, xv + m*mem_stride (elements).
extern "C" int pq3d_ingest_memories(int n_mem, const float* const* feats, const float* pos, void* xk, void* xv,
                                    int64_t mem_stride, int B, int S, int S_pitch, int D, void* stream) {
  PQ3D_CHECK_ARG(feats && pos && xk && xv && n_mem >= 1 && n_mem <= 4, "pq3d_ingest_memories: bad argument (n_mem=%d)", n_mem);
  PQ3D_CHECK_ARG(B > 0 && S > 0 && S_pitch >= S && D % 8 == 0 && mem_stride % 8 == 0,
                 "pq3d_ingest_memories: bad shape B=%d S=%d pitch=%d D=%d", B, S, S_pitch, D);
  IngestMany in;
  in.n = n_mem;
  for (int m = 0; m < 4; ++m) {
    in.feat[m] = m < n_mem ? feats[m] : nullptr;
    PQ3D_CHECK_ARG(m >= n_mem || (feats[m] != nullptr && (reinterpret_cast<uintptr_t>(feats[m]) & 15) == 0),
                   "pq3d_ingest_memories: feature table %d null or not 16-byte aligned", m);
  }
  PQ3D_CHECK_ARG((reinterpret_cast<uintptr_t>(pos) & 15) == 0 && (reinterpret_cast<uintptr_t>(xk) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(xv) & 15) == 0, "pq3d_ingest_memories: pointers must be 16-byte aligned");
  const int64_t total = static_cast<int64_t>(B) * S_pitch * (D / 8);
  PQ3D_CUDA(launch_kernel(ingest_many_kernel, dim3(grid_for(total, 256)), dim3(256), 0,
                          reinterpret_cast<cudaStream_t>(stream), in, pos, reinterpret_cast<__nv_bfloat16*>(xk),
                          reinterpret_cast<__nv_bfloat16*>(xv), mem_stride, B, S, S_pitch, D));
  return PQ3D_OK;
}

extern "C" int pq3d_add_layernorm(const float* y, int64_t y_group_stride, const float* residual, const float* gamma,
                                  const float* beta, int G, float eps, int R, int D, const float* pos, float* out_f32,
                                  void* out_bf16, void* out_pos_bf16, void* stream) {
  PQ3D_CHECK_ARG((y || residual) && gamma && beta, "pq3d_add_layernorm: null argument");
  PQ3D_CHECK_ARG(G >= 1 && R > 0 && D % 128 == 0 && D <= 128 * kLnMaxVec, "pq3d_add_layernorm: bad shape G=%d R=%d D=%d",
                 G, R, D);
  PQ3D_CHECK_ARG(G <= kLnMaxGroups, "pq3d_add_layernorm: G=%d exceeds %d", G, kLnMaxGroups);
  const int warps = 2;
  const dim3 grid((R + warps - 1) / warps), block(warps * 32);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  __nv_bfloat16* op16 = reinterpret_cast<__nv_bfloat16*>(out_pos_bf16);
#define PQ3D_LN_CASE(NV)                                                                                          \
  case NV:                                                                                                        \
    err = launch_kernel(add_layernorm_kernel<NV>, grid, block, 0, st, y, y_group_stride, residual, gamma, beta, G, \
                        eps, R, pos, out_f32, o16, op16);                                                          \
    break;
  cudaError_t err = cudaSuccess;
  switch (D / 128) {
    PQ3D_LN_CASE(1) PQ3D_LN_CASE(2) PQ3D_LN_CASE(3) PQ3D_LN_CASE(4) PQ3D_LN_CASE(5) PQ3D_LN_CASE(6) PQ3D_LN_CASE(7)
    PQ3D_LN_CASE(8)
  }
#undef PQ3D_LN_CASE
  PQ3D_CUDA(err);
  return PQ3D_OK;
}

extern "C" int pq3d_pack_mask(const uint8_t* mask, uint32_t* bits, int64_t rows, int S, int unmask_full_rows,
                              uint8_t* mask_fixed, int32_t* active_tiles, int64_t rows_per_batch, void* stream) {
  PQ3D_CHECK_ARG(mask && bits && rows > 0 && S > 0, "pq3d_pack_mask: bad argument");
  PQ3D_CHECK_ARG(active_tiles == nullptr || (rows_per_batch > 0 && rows % rows_per_batch == 0),
                 "pq3d_pack_mask: rows=%lld must be a multiple of rows_per_batch=%lld", (long long)rows,
                 (long long)rows_per_batch);
  if (active_tiles != nullptr)
    PQ3D_CUDA(cudaMemsetAsync(active_tiles, 0, sizeof(int32_t) * (rows / rows_per_batch),
                              reinterpret_cast<cudaStream_t>(stream)));
  const int W = ((S + 127) / 128) * 4;
  const int threads = S >= 1024 ? 256 : (S >= 256 ? 128 : 32);
  PQ3D_CUDA(launch_kernel(pack_mask_kernel, dim3(static_cast<unsigned>(rows)), dim3(threads), 0,
                          reinterpret_cast<cudaStream_t>(stream), mask, bits, rows, S, W, unmask_full_rows, mask_fixed,
                          active_tiles, rows_per_batch));
  return PQ3D_OK;
}

extern "C" int pq3d_mask_head_finalize(const float* raw, const uint8_t* const* mem_masks_dev, int n_mem,
                                       const uint8_t* seg_masks, float* mask_logits, uint8_t* attn_mask, int B, int S,
                                       int N, void* stream) {
  PQ3D_CHECK_ARG(raw && mem_masks_dev && seg_masks && mask_logits && attn_mask && n_mem >= 1,
                 "pq3d_mask_head_finalize: bad argument");
  dim3 grid((S + 31) / 32, (N + 31) / 32, B), block(32, 8);
  PQ3D_CUDA(launch_kernel(mask_head_finalize_kernel, grid, block, 0, reinterpret_cast<cudaStream_t>(stream), raw,
                          mem_masks_dev, n_mem, seg_masks, mask_logits, attn_mask, S, N));
  return PQ3D_OK;
}

extern "C" int pq3d_gate_mix(const float* gate_logits, const float* query, const float* update, float* out, int64_t n,
                             void* stream) {
  PQ3D_CHECK_ARG(gate_logits && query && update && out && n > 0, "pq3d_gate_mix: bad argument");
  PQ3D_CUDA(launch_kernel(gate_mix_kernel, dim3(grid_for(n, 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                          gate_logits, query, update, out, n));
  return PQ3D_OK;
}

extern "C" int pq3d_cast_bf16(const float* x, const float* add, void* out, int64_t n, void* stream) {
  PQ3D_CHECK_ARG(x && out && n > 0 && n % 4 == 0, "pq3d_cast_bf16: n=%lld must be a positive multiple of 4", (long long)n);
  PQ3D_CUDA(launch_kernel(cast_bf16_kernel, dim3(grid_for(n / 4, 256)), dim3(256), 0,
                          reinterpret_cast<cudaStream_t>(stream), x, add, reinterpret_cast<__nv_bfloat16*>(out), n / 4));
  return PQ3D_OK;
}

extern "C" int pq3d_fourier_pos(const float* xyz, int xyz_stride, const float* coord_min, const float* coord_max,
                                const float* gauss_B, void* out_bf16, int B, int L, int d_pos, void* stream) {
  PQ3D_CHECK_ARG(xyz && coord_min && coord_max && gauss_B && out_bf16, "pq3d_fourier_pos: null argument");
  PQ3D_CHECK_ARG(B > 0 && L > 0 && d_pos > 0 && d_pos % 2 == 0 && xyz_stride >= 3, "pq3d_fourier_pos: bad shape");
  const int64_t total = static_cast<int64_t>(B) * L * (d_pos / 2);
  PQ3D_CUDA(launch_kernel(fourier_pos_kernel, dim3(grid_for(total, 256)), dim3(256), 0,
                          reinterpret_cast<cudaStream_t>(stream), xyz, xyz_stride, coord_min, coord_max, gauss_B,
                          reinterpret_cast<__nv_bfloat16*>(out_bf16), B, L, d_pos / 2));
  return PQ3D_OK;
}

extern "C" int pq3d_pairwise_locs(const float* centers, int c_stride, float* out, int B, int N, float eps,
                                  void* stream) {
  PQ3D_CHECK_ARG(centers && out && B > 0 && N > 0 && c_stride >= 3, "pq3d_pairwise_locs: bad argument");
  PQ3D_CUDA(launch_kernel(pairwise_locs_kernel, dim3(B), dim3(512), 0, reinterpret_cast<cudaStream_t>(stream), centers,
                          c_stride, out, N, eps));
  return PQ3D_OK;
}
