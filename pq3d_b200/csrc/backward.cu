// Streaming companions of the backward pass.  The GEMM-shaped work of the backward (dgrad, wgrad, the
// four attention-backward products) runs on the same tcgen05 kernels as the forward
// (pq3d_linear_bf16 / pq3d_bgemm_bf16); what is left is HBM-bound: transposes + casts that put operands into
// the K-major layout those kernels take, LayerNorm backward, bias gradients, the softmax backward and the
// spatial-bias backward.  Reference: torch autograd through modules/grounding/query_encoder.py and
// torch/nn/functional.py:6630-6654 in the reference's training step (trainer/query3d_trainer.py:18-28).
#include "host_common.h"
#include "ptx.cuh"

namespace pq3d {

// ------------------------------------------------------------------------------------------------
// transpose_cast: out_t[b][c][r] = bf16(scale * in[b][r][c] * (gate[b][r][c] > 0)), r < R; zero for R <= r < Rp.
// Optionally also the un-transposed bf16 copy.  Two batch levels (b1, b2) with independent strides so that
// per-(scene, head) slices of packed tensors transpose in one launch.  32x32 tiles through shared memory.
// ------------------------------------------------------------------------------------------------
struct TransposeParams {
  const void* in;
  const __nv_bfloat16* gate;
  __nv_bfloat16* out_t;
  __nv_bfloat16* out_c;
  int64_t ld_in, in_b1, in_b2;
  int64_t ld_gate, gate_b1, gate_b2;
  int64_t ld_t, t_b1, t_b2;
  int64_t ld_c, c_b1, c_b2;
  int R, C, Rp, B2;
  int in_fp32;
  float scale;
};

__global__ void transpose_cast_kernel(const TransposeParams p) {
  pdl_sync();
  __shared__ float tile[32][33];
  const int b1 = blockIdx.z / p.B2, b2 = blockIdx.z % p.B2;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < p.R && c < p.C) {
      const int64_t off = b1 * p.in_b1 + b2 * p.in_b2 + static_cast<int64_t>(r) * p.ld_in + c;
      v = p.in_fp32 ? reinterpret_cast<const float*>(p.in)[off]
                    : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.in)[off]);
      v *= p.scale;
      if (p.gate != nullptr) {
        const float gt = __bfloat162float(p.gate[b1 * p.gate_b1 + b2 * p.gate_b2 + static_cast<int64_t>(r) * p.ld_gate + c]);
        if (!(gt > 0.f)) v = 0.f;
      }
      if (p.out_c != nullptr)
        p.out_c[b1 * p.c_b1 + b2 * p.c_b2 + static_cast<int64_t>(r) * p.ld_c + c] = __float2bfloat16_rn(v);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  if (p.out_t == nullptr) return;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < p.C && r < p.Rp)
      p.out_t[b1 * p.t_b1 + b2 * p.t_b2 + static_cast<int64_t>(c) * p.ld_t + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

// Vectorised variant (16-byte global accesses, 64x64 tiles) for the common case: C, Rp, every pitch and batch stride
// multiples of 8 elements and 16-byte aligned bases.  Same semantics as transpose_cast_kernel.
constexpr int kTcPitch = 66;   // bf16 elements per shared-memory tile row (33 words: conflict-light both ways)

__global__ void __launch_bounds__(256) transpose_cast_vec_kernel(const TransposeParams p) {
  pdl_sync();
  __shared__ __align__(16) __nv_bfloat16 tile[64 * kTcPitch];      // tile[c][r]
  const int b1 = blockIdx.z / p.B2, b2 = blockIdx.z % p.B2;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int t = threadIdx.x;
  const int chunk = t & 7;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int rl = (t >> 3) + it * 32;
    const int r = r0 + rl, c = c0 + chunk * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (r < p.R && c < p.C) {
      const int64_t off = b1 * p.in_b1 + b2 * p.in_b2 + static_cast<int64_t>(r) * p.ld_in + c;
      if (p.in_fp32) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.in) + off));
        const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.in) + off) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.in) + off));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
        for (int j = 0; j < 4; ++j) { v[2 * j] = __bfloat162float(h[j].x); v[2 * j + 1] = __bfloat162float(h[j].y); }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= p.scale;
      if (p.gate != nullptr) {
        const uint4 g = __ldg(reinterpret_cast<const uint4*>(p.gate + b1 * p.gate_b1 + b2 * p.gate_b2 +
                                                             static_cast<int64_t>(r) * p.ld_gate + c));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (!(__bfloat162float(h[j].x) > 0.f)) v[2 * j] = 0.f;
          if (!(__bfloat162float(h[j].y) > 0.f)) v[2 * j + 1] = 0.f;
        }
      }
      if (p.out_c != nullptr) {
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(p.out_c + b1 * p.c_b1 + b2 * p.c_b2 + static_cast<int64_t>(r) * p.ld_c + c) = o;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) tile[(chunk * 8 + j) * kTcPitch + rl] = __float2bfloat16_rn(v[j]);
  }
  __syncthreads();
  if (p.out_t == nullptr) return;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int cl = (t >> 3) + it * 32;
    const int c = c0 + cl, r = r0 + chunk * 8;
    if (c < p.C && r < p.Rp) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(&tile[cl * kTcPitch + chunk * 8]);
      uint4 o;
      o.x = src[0]; o.y = src[1]; o.z = src[2]; o.w = src[3];
      *reinterpret_cast<uint4*>(p.out_t + b1 * p.t_b1 + b2 * p.t_b2 + static_cast<int64_t>(c) * p.ld_t + r) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// colsum: out[c] (+)= sum_r in[r][c] * (gate[r][c] > 0)  (bias gradients).  Each block takes 64 columns x one chunk of
// rows: a warp reads one row segment of 64 columns per step (2 columns per lane: 128 B of bf16 / 256 B of fp32,
// coalesced), the 8 warps stride over the chunk's rows; partial sums meet in shared memory and leave with one atomicAdd
// per column (the host zero-fills `out` first unless accumulating).  C must be even.
// ------------------------------------------------------------------------------------------------
__global__ void colsum_kernel(const void* __restrict__ in, int in_fp32, int64_t ld, const __nv_bfloat16* __restrict__ gate,
                              int64_t ld_gate, float* __restrict__ out, int R, int C, int rows_per_block, float scale) {
  pdl_sync();
  __shared__ float part[8][64];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + lane * 2;
  const int r_begin = blockIdx.y * rows_per_block;
  const int r_end = min(R, r_begin + rows_per_block);
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    for (int r = r_begin + wid; r < r_end; r += 8) {
      const int64_t off = static_cast<int64_t>(r) * ld + c;
      float v0, v1;
      if (in_fp32) {
        const float2 v = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(in) + off);
        v0 = v.x; v1 = v.y;
      } else {
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(in) + off);
        v0 = __bfloat162float(v.x); v1 = __bfloat162float(v.y);
      }
      if (gate != nullptr) {
        const __nv_bfloat162 gt = *reinterpret_cast<const __nv_bfloat162*>(gate + static_cast<int64_t>(r) * ld_gate + c);
        if (!(__bfloat162float(gt.x) > 0.f)) v0 = 0.f;
        if (!(__bfloat162float(gt.y) > 0.f)) v1 = 0.f;
      }
      s0 += v0; s1 += v1;
    }
  }
  part[wid][lane * 2] = s0;
  part[wid][lane * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int cc = blockIdx.x * 64 + threadIdx.x;
    if (cc < C) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
      atomicAdd(&out[cc], t * scale);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// layernorm_bwd: forward was out = (1/G) sum_g LN_g(x_g), x_g = residual + y_g.  Given d_out:
//   d_x_g   = rstd_g * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)),  dxhat = (d_out / G) * gamma_g
//   d_res   = sum_g d_x_g          d_gamma_g += sum_rows (d_out / G) * xhat      d_beta_g += sum_rows (d_out / G)
// One warp per row, statistics recomputed (nothing but x is saved).  Affine gradients: per-block partial sums in
// shared memory, then one atomicAdd per column per block (caller zero-initialises d_gamma / d_beta).
// ------------------------------------------------------------------------------------------------
constexpr int kLnbWarps = 4;

template <int NV>
__global__ void __launch_bounds__(kLnbWarps * 32)
layernorm_bwd_kernel(const float* __restrict__ y, int64_t y_group_stride, const float* __restrict__ residual,
                     const float* __restrict__ gamma, const float* __restrict__ d_out, int G, float eps, int R,
                     float* __restrict__ d_x, int64_t dx_group_stride, __nv_bfloat16* __restrict__ d_x16,
                     float* __restrict__ d_res, float* __restrict__ d_gamma, float* __restrict__ d_beta,
                     uint32_t drop_thresh, float drop_scale, const uint32_t* __restrict__ seed, uint32_t site,
                     const float* __restrict__ row_w, int rows_per_scene) {
  pdl_sync();
  constexpr int D = NV * 128;
  const uint32_t key = drop_thresh != 0 ? drop_key(__ldg(seed), site) : 0u;
  // per-warp partial affine gradients meet here once per group: no atomics inside the row loop
  __shared__ float s_dg[kLnbWarps][D], s_db[kLnbWarps][D];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float inv_g = 1.f / static_cast<float>(G);
  constexpr float inv_d = 1.f / static_cast<float>(D);
  for (int g = 0; g < G; ++g) {
    float4 pg[NV], pb[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) pg[i] = pb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 ga[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) ga[i] = __ldg(reinterpret_cast<const float4*>(gamma + g * D) + i * 32 + lane);
    for (int row = blockIdx.x * kLnbWarps + wid; row < R; row += gridDim.x * kLnbWarps) {
      const int64_t base = static_cast<int64_t>(row) * D;
      float4 x[NV], go[NV];
      uint32_t km[NV];                  // dropout keep bits of this lane's 4 columns per chunk
      float sum = 0.f;
      const float wrow = row_w != nullptr ? __ldg(row_w + (row / rows_per_scene) * G + g) : inv_g;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float4 v = residual != nullptr ? __ldg(reinterpret_cast<const float4*>(residual + base) + i * 32 + lane)
                                       : make_float4(0, 0, 0, 0);
        km[i] = 0xfu;
        if (y != nullptr) {
          float4 t = __ldg(reinterpret_cast<const float4*>(y + g * y_group_stride + base) + i * 32 + lane);
          if (drop_thresh != 0) {
            const uint32_t e = static_cast<uint32_t>((static_cast<int64_t>(g) * R + row) * D + (i * 32 + lane) * 4);
            km[i] = (drop_keep(key, e, drop_thresh) ? 1u : 0u) | (drop_keep(key, e + 1, drop_thresh) ? 2u : 0u) |
                    (drop_keep(key, e + 2, drop_thresh) ? 4u : 0u) | (drop_keep(key, e + 3, drop_thresh) ? 8u : 0u);
            t.x = (km[i] & 1u) ? t.x * drop_scale : 0.f; t.y = (km[i] & 2u) ? t.y * drop_scale : 0.f;
            t.z = (km[i] & 4u) ? t.z * drop_scale : 0.f; t.w = (km[i] & 8u) ? t.w * drop_scale : 0.f;
          }
          v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        x[i] = v;
        go[i] = __ldg(reinterpret_cast<const float4*>(d_out + base) + i * 32 + lane);
        go[i].x *= wrow; go[i].y *= wrow; go[i].z *= wrow; go[i].w *= wrow;
        sum += (v.x + v.y) + (v.z + v.w);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum * inv_d;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        x[i].x -= mean; x[i].y -= mean; x[i].z -= mean; x[i].w -= mean;
        sq += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = rsqrtf(sq * inv_d + eps);
      float m1 = 0.f, m2 = 0.f;   // mean(dxhat), mean(dxhat * xhat)
      float4 dxh[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        x[i].x *= rstd; x[i].y *= rstd; x[i].z *= rstd; x[i].w *= rstd;   // xhat
        dxh[i] = make_float4(go[i].x * ga[i].x, go[i].y * ga[i].y, go[i].z * ga[i].z, go[i].w * ga[i].w);
        m1 += (dxh[i].x + dxh[i].y) + (dxh[i].z + dxh[i].w);
        m2 += (dxh[i].x * x[i].x + dxh[i].y * x[i].y) + (dxh[i].z * x[i].z + dxh[i].w * x[i].w);
        pg[i].x += go[i].x * x[i].x; pg[i].y += go[i].y * x[i].y; pg[i].z += go[i].z * x[i].z; pg[i].w += go[i].w * x[i].w;
        pb[i].x += go[i].x; pb[i].y += go[i].y; pb[i].z += go[i].z; pb[i].w += go[i].w;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        m1 += __shfl_xor_sync(0xffffffffu, m1, o);
        m2 += __shfl_xor_sync(0xffffffffu, m2, o);
      }
      m1 *= inv_d;
      m2 *= inv_d;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float4 dx;
        dx.x = rstd * (dxh[i].x - m1 - x[i].x * m2);
        dx.y = rstd * (dxh[i].y - m1 - x[i].y * m2);
        dx.z = rstd * (dxh[i].z - m1 - x[i].z * m2);
        dx.w = rstd * (dxh[i].w - m1 - x[i].w * m2);
        const int64_t off = base + (i * 32 + lane) * 4;
        float4 dy = dx;                 // gradient of the branch input y: through the dropout mask
        if (drop_thresh != 0) {
          dy.x = (km[i] & 1u) ? dx.x * drop_scale : 0.f; dy.y = (km[i] & 2u) ? dx.y * drop_scale : 0.f;
          dy.z = (km[i] & 4u) ? dx.z * drop_scale : 0.f; dy.w = (km[i] & 8u) ? dx.w * drop_scale : 0.f;
        }
        if (d_x != nullptr) *reinterpret_cast<float4*>(d_x + g * dx_group_stride + off) = dy;
        if (d_x16 != nullptr) {
          uint2 o;
          o.x = pack_bf16x2(dy.x, dy.y); o.y = pack_bf16x2(dy.z, dy.w);
          *reinterpret_cast<uint2*>(d_x16 + g * dx_group_stride + off) = o;
        }
        if (d_res != nullptr) {
          float4* dr = reinterpret_cast<float4*>(d_res + off);
          if (g == 0) {
            *dr = dx;
          } else {
            float4 o = *dr;
            o.x += dx.x; o.y += dx.y; o.z += dx.z; o.w += dx.w;
            *dr = o;
          }
        }
      }
    }
    if (d_gamma != nullptr || d_beta != nullptr) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        *reinterpret_cast<float4*>(&s_dg[wid][(i * 32 + lane) * 4]) = pg[i];
        *reinterpret_cast<float4*>(&s_db[wid][(i * 32 + lane) * 4]) = pb[i];
      }
      __syncthreads();
      for (int i = threadIdx.x; i < D; i += blockDim.x) {
        float tg = 0.f, tb = 0.f;
#pragma unroll
        for (int w = 0; w < kLnbWarps; ++w) { tg += s_dg[w][i]; tb += s_db[w][i]; }
        if (d_gamma != nullptr) atomicAdd(&d_gamma[g * D + i], tg);
        if (d_beta != nullptr) atomicAdd(&d_beta[g * D + i], tb);
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// attn_delta: delta[b][h][n] = sum_d dO[b*N+n][h*64+d] * O[b*N+n][h*64+d]   (bf16 inputs, fp32 out)
// ------------------------------------------------------------------------------------------------
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ dO, const __nv_bfloat16* __restrict__ O, int64_t ld,
                                  float* __restrict__ delta, int B, int H, int N) {
  pdl_sync();
  const int64_t w = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;   // (b, n, h)
  const int lane = threadIdx.x & 31;
  if (w >= static_cast<int64_t>(B) * N * H) return;
  const int h = static_cast<int>(w % H);
  const int64_t row = w / H;
  const __nv_bfloat162 a = reinterpret_cast<const __nv_bfloat162*>(dO + row * ld + h * 64)[lane];
  const __nv_bfloat162 o = reinterpret_cast<const __nv_bfloat162*>(O + row * ld + h * 64)[lane];
  float s = __bfloat162float(a.x) * __bfloat162float(o.x) + __bfloat162float(a.y) * __bfloat162float(o.y);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) {
    const int b = static_cast<int>(row / N), n = static_cast<int>(row % N);
    delta[(static_cast<int64_t>(b) * H + h) * N + n] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// softmax_bwd: scores S2 (log2 domain, recomputed by a batched GEMM) and dP = dO V^T arrive as fp32 [B][H][N][ld];
//   P  = ex2(S2 + bias - m) / l   (0 where masked)          dS2 = ln2 * P * (dP - delta)
// written in both orientations as bf16: P, dS2 [B][H][N][ld] and P^T, dS2^T [B][H][ld][Np] (columns n >= N zero),
// the K-major operands of the dV / dQ / dK products.
// ------------------------------------------------------------------------------------------------
struct SoftmaxBwdParams {
  const float *S2, *dP, *delta, *m, *l, *bias;
  const uint32_t* mask_bits;
  int64_t mask_b_stride, mask_h_stride, mask_q_stride, bias_ld;
  __nv_bfloat16 *P, *dS, *Pt, *dSt;
  int B, H, N, S, ld, Np;
};

__global__ void __launch_bounds__(256) softmax_bwd_kernel(const SoftmaxBwdParams p) {
  pdl_sync();
  // 64 queries x 64 keys per block; thread (row, chunk) owns 8 consecutive keys of one query row (two rows per thread)
  __shared__ __align__(16) __nv_bfloat16 tP[64 * kTcPitch], tD[64 * kTcPitch];     // [s][n]
  const int bh = blockIdx.z, b = bh / p.H, h = bh % p.H;
  const int n0 = blockIdx.y * 64, s0 = blockIdx.x * 64;
  const int t = threadIdx.x, chunk = t & 7;
  const int64_t base = static_cast<int64_t>(bh) * p.N * p.ld;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int nl = (t >> 3) + it * 32;
    const int n = n0 + nl, s = s0 + chunk * 8;
    float pv[8], dv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { pv[j] = 0.f; dv[j] = 0.f; }
    if (n < p.N && s < p.ld) {
      const int64_t row = static_cast<int64_t>(bh) * p.N + n;
      const int64_t off = base + static_cast<int64_t>(n) * p.ld + s;
      if (s < p.S) {
        uint32_t mbits = 0;
        if (p.mask_bits != nullptr)
          mbits = (p.mask_bits[b * p.mask_b_stride + h * p.mask_h_stride + n * p.mask_q_stride + (s >> 5)] >> (s & 31)) & 0xffu;
        const float m = p.m[row], inv_l = 1.f / p.l[row], dl = p.delta[row];
        const float4 sa = *reinterpret_cast<const float4*>(p.S2 + off), sb = *reinterpret_cast<const float4*>(p.S2 + off + 4);
        const float4 da = *reinterpret_cast<const float4*>(p.dP + off), db = *reinterpret_cast<const float4*>(p.dP + off + 4);
        float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
        const float dp[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
        if (p.bias != nullptr) {
          const float4 ba = *reinterpret_cast<const float4*>(p.bias + row * p.bias_ld + s);
          const float4 bb = *reinterpret_cast<const float4*>(p.bias + row * p.bias_ld + s + 4);
          sc[0] += ba.x; sc[1] += ba.y; sc[2] += ba.z; sc[3] += ba.w; sc[4] += bb.x; sc[5] += bb.y; sc[6] += bb.z; sc[7] += bb.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (s + j < p.S && !((mbits >> j) & 1u)) {
            pv[j] = exp2f(sc[j] - m) * inv_l;
            dv[j] = 0.6931471805599453f * pv[j] * (dp[j] - dl);
          }
        }
      }
      uint4 o;
      o.x = pack_bf16x2(dv[0], dv[1]); o.y = pack_bf16x2(dv[2], dv[3]); o.z = pack_bf16x2(dv[4], dv[5]); o.w = pack_bf16x2(dv[6], dv[7]);
      *reinterpret_cast<uint4*>(p.dS + off) = o;
      if (p.P != nullptr) {
        o.x = pack_bf16x2(pv[0], pv[1]); o.y = pack_bf16x2(pv[2], pv[3]); o.z = pack_bf16x2(pv[4], pv[5]); o.w = pack_bf16x2(pv[6], pv[7]);
        *reinterpret_cast<uint4*>(p.P + off) = o;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      tP[(chunk * 8 + j) * kTcPitch + nl] = __float2bfloat16_rn(pv[j]);
      tD[(chunk * 8 + j) * kTcPitch + nl] = __float2bfloat16_rn(dv[j]);
    }
  }
  __syncthreads();
  const int64_t tbase = static_cast<int64_t>(bh) * p.ld * p.Np;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int sl = (t >> 3) + it * 32;
    const int s = s0 + sl, n = n0 + chunk * 8;
    if (s < p.ld && n < p.Np) {
      const uint32_t* a = reinterpret_cast<const uint32_t*>(&tP[sl * kTcPitch + chunk * 8]);
      const uint32_t* d = reinterpret_cast<const uint32_t*>(&tD[sl * kTcPitch + chunk * 8]);
      *reinterpret_cast<uint4*>(p.Pt + tbase + static_cast<int64_t>(s) * p.Np + n) = make_uint4(a[0], a[1], a[2], a[3]);
      *reinterpret_cast<uint4*>(p.dSt + tbase + static_cast<int64_t>(s) * p.Np + n) = make_uint4(d[0], d[1], d[2], d[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// spatial_bias_bwd: bias[b,h,n,m] = log2(max(relu(v), 1e-6)), v = loc[b,n,m,:] . w[h,:] + bias[h];
// d_v = dS2[b,h,n,m] / (v ln2) where v > 1e-6, else 0;  d_w[h,:] += d_v * loc,  d_b[h] += d_v.
// ------------------------------------------------------------------------------------------------
__global__ void spatial_bias_bwd_kernel(const float* __restrict__ locs, const float* __restrict__ w,
                                        const float* __restrict__ bias, const __nv_bfloat16* __restrict__ dS, int64_t ld,
                                        float* __restrict__ d_w, float* __restrict__ d_b, int B, int H, int N) {
  pdl_sync();
  __shared__ float red[6][8];
  const int h = blockIdx.y;
  float acc[6] = {0, 0, 0, 0, 0, 0};
  const int64_t pairs = static_cast<int64_t>(B) * N * N;
  const float w0 = w[h * 5], w1 = w[h * 5 + 1], w2 = w[h * 5 + 2], w3 = w[h * 5 + 3], w4 = w[h * 5 + 4], bb = bias[h];
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < pairs;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i % N);
    const int n = static_cast<int>((i / N) % N);
    const int b = static_cast<int>(i / (static_cast<int64_t>(N) * N));
    const float* l5 = locs + i * 5;
    const float v = fmaf(l5[0], w0, fmaf(l5[1], w1, fmaf(l5[2], w2, fmaf(l5[3], w3, fmaf(l5[4], w4, bb)))));
    if (v > 1e-6f) {
      const float g = __bfloat162float(dS[((static_cast<int64_t>(b) * H + h) * N + n) * ld + m]) / (v * 0.6931471805599453f);
      acc[0] += g * l5[0]; acc[1] += g * l5[1]; acc[2] += g * l5[2]; acc[3] += g * l5[3]; acc[4] += g * l5[4];
      acc[5] += g;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (lane == 0) red[k][wid] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[threadIdx.x][i];
    if (threadIdx.x < 5) atomicAdd(&d_w[h * 5 + threadIdx.x], t);
    else atomicAdd(&d_b[h], t);
  }
}

// out = a + b (+ c), fp32, for accumulating gradient streams
__global__ void add3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                            float* __restrict__ out, int64_t n4) {
  pdl_sync();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 x = reinterpret_cast<const float4*>(a)[i];
    const float4 y = reinterpret_cast<const float4*>(b)[i];
    x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
    if (c != nullptr) {
      const float4 z = reinterpret_cast<const float4*>(c)[i];
      x.x += z.x; x.y += z.y; x.z += z.z; x.w += z.w;
    }
    reinterpret_cast<float4*>(out)[i] = x;
  }
}

}  // namespace pq3d

using namespace pq3d;

static inline int grid_for_b(int64_t work_items, int threads) {
  int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

extern "C" int pq3d_transpose_cast(const void* in, int in_fp32, int64_t ld_in, int64_t in_b1, int64_t in_b2,
                                   const void* gate, int64_t ld_gate, int64_t gate_b1, int64_t gate_b2, void* out_t,
                                   int64_t ld_t, int64_t t_b1, int64_t t_b2, void* out_c, int64_t ld_c, int64_t c_b1,
                                   int64_t c_b2, int R, int C, int Rp, int B1, int B2, float scale, void* stream) {
  PQ3D_CHECK_ARG(in && (out_t || out_c) && R > 0 && C > 0 && Rp >= R && B1 > 0 && B2 > 0,
                 "pq3d_transpose_cast: bad argument");
  TransposeParams p;
  p.in = in; p.gate = reinterpret_cast<const __nv_bfloat16*>(gate);
  p.out_t = reinterpret_cast<__nv_bfloat16*>(out_t); p.out_c = reinterpret_cast<__nv_bfloat16*>(out_c);
  p.ld_in = ld_in; p.in_b1 = in_b1; p.in_b2 = in_b2;
  p.ld_gate = ld_gate; p.gate_b1 = gate_b1; p.gate_b2 = gate_b2;
  p.ld_t = ld_t; p.t_b1 = t_b1; p.t_b2 = t_b2;
  p.ld_c = ld_c; p.c_b1 = c_b1; p.c_b2 = c_b2;
  p.R = R; p.C = C; p.Rp = Rp; p.B2 = B2; p.in_fp32 = in_fp32; p.scale = scale;
  auto m8 = [](int64_t v) { return v % 8 == 0; };
  auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool vec = m8(C) && m8(Rp) && m8(ld_in) && m8(in_b1) && m8(in_b2) && a16(in) &&
                   (!gate || (m8(ld_gate) && m8(gate_b1) && m8(gate_b2) && a16(gate))) &&
                   (!out_t || (m8(ld_t) && m8(t_b1) && m8(t_b2) && a16(out_t))) &&
                   (!out_c || (m8(ld_c) && m8(c_b1) && m8(c_b2) && a16(out_c)));
  if (vec) {
    dim3 grid((C + 63) / 64, (Rp + 63) / 64, B1 * B2);
    PQ3D_CUDA(launch_kernel(transpose_cast_vec_kernel, grid, dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), p));
    return PQ3D_OK;
  }
  dim3 grid((C + 31) / 32, (Rp + 31) / 32, B1 * B2), block(32, 8);
  PQ3D_CUDA(launch_kernel(transpose_cast_kernel, grid, block, 0, reinterpret_cast<cudaStream_t>(stream), p));
  return PQ3D_OK;
}

extern "C" int pq3d_colsum(const void* in, int in_fp32, int64_t ld, const void* gate, int64_t ld_gate, float* out, int R,
                           int C, int accumulate, float scale, void* stream) {
  PQ3D_CHECK_ARG(in && out && R > 0 && C > 0, "pq3d_colsum: bad argument");
  PQ3D_CHECK_ARG(C % 2 == 0 && ld % 2 == 0 && ld_gate % 2 == 0 && (reinterpret_cast<uintptr_t>(in) & 7) == 0,
                 "pq3d_colsum: C, ld must be even and the input 8-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!accumulate) PQ3D_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * C, st));
  const int col_blocks = (C + 63) / 64;
  int chunks = (sm_count() * 4 + col_blocks - 1) / col_blocks;        // ~4 blocks per SM in total
  const int max_chunks = (R + 63) / 64;                               // at least 64 rows per block
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const int rows_per_block = (R + chunks - 1) / chunks;
  PQ3D_CUDA(launch_kernel(colsum_kernel, dim3(col_blocks, (R + rows_per_block - 1) / rows_per_block), dim3(256), 0, st, in,
                          in_fp32, ld, reinterpret_cast<const __nv_bfloat16*>(gate), ld_gate, out, R, C, rows_per_block, scale));
  return PQ3D_OK;
}

extern "C" int pq3d_layernorm_bwd(const float* y, int64_t y_group_stride, const float* residual, const float* gamma,
                                  const float* d_out, int G, float eps, int R, int D, float* d_x,
                                  int64_t dx_group_stride, void* d_x_bf16, float* d_res, float* d_gamma, float* d_beta,
                                  float drop_p, const uint32_t* seed_dev, uint32_t site, const float* row_w,
                                  int rows_per_scene, void* stream) {
  PQ3D_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || seed_dev != nullptr),
                 "pq3d_layernorm_bwd: dropout needs p in [0,1) and a device seed");
  PQ3D_CHECK_ARG(row_w == nullptr || (rows_per_scene > 0 && R % rows_per_scene == 0),
                 "pq3d_layernorm_bwd: row weights need rows_per_scene dividing R");
  const uint32_t thresh = drop_p > 0.f ? drop_threshold(drop_p) : 0u;
  const float dscale = 1.f / (1.f - drop_p);
  PQ3D_CHECK_ARG((y || residual) && gamma && d_out && G >= 1 && R > 0 && D % 128 == 0 && D <= 1024,
                 "pq3d_layernorm_bwd: bad argument (D must be a multiple of 128, at most 1024)");
  int blocks = (R + kLnbWarps - 1) / kLnbWarps;
  if (blocks > 2 * sm_count()) blocks = 2 * sm_count();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t err = cudaSuccess;
#define PQ3D_LNB_CASE(NV)                                                                                            \
  case NV:                                                                                                           \
    err = launch_kernel(layernorm_bwd_kernel<NV>, dim3(blocks), dim3(kLnbWarps * 32), 0, st, y, y_group_stride,      \
                        residual, gamma, d_out, G, eps, R, d_x, dx_group_stride,                                     \
                        reinterpret_cast<__nv_bfloat16*>(d_x_bf16), d_res, d_gamma, d_beta, thresh, dscale,         \
                        seed_dev, site, row_w, rows_per_scene);                                                      \
    break;
  switch (D / 128) {
    PQ3D_LNB_CASE(1) PQ3D_LNB_CASE(2) PQ3D_LNB_CASE(3) PQ3D_LNB_CASE(4) PQ3D_LNB_CASE(5) PQ3D_LNB_CASE(6)
    PQ3D_LNB_CASE(7) PQ3D_LNB_CASE(8)
  }
#undef PQ3D_LNB_CASE
  PQ3D_CUDA(err);
  return PQ3D_OK;
}

extern "C" int pq3d_attn_delta(const void* dO, const void* O, int64_t ld, float* delta, int B, int H, int N,
                               void* stream) {
  PQ3D_CHECK_ARG(dO && O && delta && B > 0 && H > 0 && N > 0, "pq3d_attn_delta: bad argument");
  const int64_t warps = static_cast<int64_t>(B) * N * H;
  PQ3D_CUDA(launch_kernel(attn_delta_kernel, dim3(static_cast<unsigned>((warps * 32 + 255) / 256)), dim3(256), 0,
                          reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const __nv_bfloat16*>(dO),
                          reinterpret_cast<const __nv_bfloat16*>(O), ld, delta, B, H, N));
  return PQ3D_OK;
}

extern "C" int pq3d_softmax_bwd(const float* S2, const float* dP, const float* delta, const float* m, const float* l,
                                const float* bias, int64_t bias_ld, const uint32_t* mask_bits, int64_t mask_b_stride,
                                int64_t mask_h_stride, int64_t mask_q_stride, void* P, void* dS, void* Pt, void* dSt,
                                int B, int H, int N, int S, int ld, int Np, void* stream) {
  PQ3D_CHECK_ARG(S2 && dP && delta && m && l && dS && Pt && dSt, "pq3d_softmax_bwd: null argument");
  PQ3D_CHECK_ARG(B > 0 && H > 0 && N > 0 && S > 0 && ld >= S && Np >= N, "pq3d_softmax_bwd: bad shape");
  PQ3D_CHECK_ARG(ld % 32 == 0 && Np % 8 == 0 && bias_ld % 4 == 0,
                 "pq3d_softmax_bwd: ld must be a multiple of 32, Np of 8, bias_ld of 4 (16-byte vector accesses)");
  SoftmaxBwdParams p;
  p.S2 = S2; p.dP = dP; p.delta = delta; p.m = m; p.l = l; p.bias = bias; p.bias_ld = bias_ld;
  p.mask_bits = mask_bits; p.mask_b_stride = mask_b_stride; p.mask_h_stride = mask_h_stride;
  p.mask_q_stride = mask_q_stride;
  p.P = reinterpret_cast<__nv_bfloat16*>(P); p.dS = reinterpret_cast<__nv_bfloat16*>(dS);
  p.Pt = reinterpret_cast<__nv_bfloat16*>(Pt); p.dSt = reinterpret_cast<__nv_bfloat16*>(dSt);
  p.B = B; p.H = H; p.N = N; p.S = S; p.ld = ld; p.Np = Np;
  const int rows = Np > N ? Np : N;
  dim3 grid((ld + 63) / 64, (rows + 63) / 64, B * H);
  PQ3D_CUDA(launch_kernel(softmax_bwd_kernel, grid, dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), p));
  return PQ3D_OK;
}

extern "C" int pq3d_spatial_bias_bwd(const float* pairwise_locs, const float* loc_w, const float* loc_b, const void* dS,
                                     int64_t ld, float* d_w, float* d_b, int B, int H, int N, void* stream) {
  PQ3D_CHECK_ARG(pairwise_locs && loc_w && loc_b && dS && d_w && d_b, "pq3d_spatial_bias_bwd: null argument");
  const int64_t pairs = static_cast<int64_t>(B) * N * N;
  int bx = static_cast<int>((pairs + 255) / 256);
  if (bx > 64) bx = 64;
  PQ3D_CUDA(launch_kernel(spatial_bias_bwd_kernel, dim3(bx, H), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                          pairwise_locs, loc_w, loc_b, reinterpret_cast<const __nv_bfloat16*>(dS), ld, d_w, d_b, B, H, N));
  return PQ3D_OK;
}

extern "C" int pq3d_add3(const float* a, const float* b, const float* c, float* out, int64_t n, void* stream) {
  PQ3D_CHECK_ARG(a && b && out && n > 0 && n % 4 == 0, "pq3d_add3: n must be a positive multiple of 4");
  PQ3D_CUDA(launch_kernel(add3_kernel, dim3(grid_for_b(n / 4, 256)), dim3(256), 0,
                          reinterpret_cast<cudaStream_t>(stream), a, b, c, out, n / 4));
  return PQ3D_OK;
}
