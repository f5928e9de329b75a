// pq3d_attention_bwd: backward of the masked softmax attention core (pq3d_attention_fwd) in ONE kernel on tcgen05.
//
// What the reference runs here is autograd through torch/nn/functional.py:6630-6647 (baddbmm, softmax, bmm) for
// CrossAttentionLayer / SelfAttentionLayer (modules/grounding/query_encoder.py:288-307, 217-223) and through
// modules/layers/transformers.py:224-236 for the spatial self-attention: five GEMM-shaped products with the score
// matrices materialised in HBM.  Here scores never leave the SM: per 128-key tile
//
//   S^T  = K Q2^T                     (keys on TMEM lanes, queries on columns; Q2 = log2(e)/sqrt(dh) * q)
//   dP^T = V dO^T
//   P^T  = ex2(S^T + bias^T - m) / l  (m, l saved by the forward; 0 where masked)
//   dS^T = ln2 * P^T o (dP^T - delta) (delta = rowsum(dO o O))
//   dV   = P^T dO      dK = dS^T Q2      dQ += dS K           (dQ leaves scaled by log2(e)/sqrt(dh))
//
// All five products are tcgen05.mma with fp32 accumulators in TMEM.  S^T / dP^T take every operand K-major straight
// from the TMA tiles; dV / dK / dQ re-use THE SAME shared-memory tiles (dO, Q2, K, and the P^T / dS^T tiles the
// softmax warps write) as MN-major operands (instruction-descriptor transpose bits), so nothing is transposed in
// memory and no operand is staged twice.
//
// One CTA = (head, scene, chunk of consecutive key tiles).  320 threads: warp 0 TMA producer (Q2 + dO once, then a
// 2-stage ring of K / V tiles), warp 1 MMA issuer, warps 2..9 softmax + drains (two per TMEM lane quadrant, 64 query
// columns each).  dK / dV rows are owned by exactly one CTA (plain stores); dQ is accumulated in TMEM over the CTA's
// tiles and leaves through fp32 atomics when the key range is split over several CTAs.  Requires Nq <= 128.
#include <cstring>

#include "host_common.h"
#include "ptx.cuh"

namespace pq3d {

constexpr int kBwdSoftmaxWarps = 8;
constexpr int kBwdThreads = 64 + 32 * kBwdSoftmaxWarps;
constexpr int kBwdStages = 2;
constexpr int kTileBytes = 128 * 64 * 2;              // one [128 rows x 64 bf16] swizzled tile: 16 KB
constexpr int kSqBytes = 128 * 128 * 2;               // P^T / dS^T tile: two atom columns of 16 KB
constexpr int kBwdSmem = 2 * kTileBytes + kBwdStages * 2 * kTileBytes + 2 * kSqBytes + 256 + 1024;

struct AttnBwdMaps {
  CUtensorMap q, d_o, k, v;
};

struct AttnBwdParams {
  const uint32_t* mask_bits;
  int64_t mask_b_stride, mask_h_stride, mask_q_stride;
  const float* bias;
  int64_t bias_ld;
  const float *stat_m, *stat_l, *delta;
  __nv_bfloat16 *dK, *dV;
  int64_t ld_dk, ld_dv;
  float* dQ;
  int64_t ld_dq;
  __nv_bfloat16* dS_out;
  int64_t ds_ld;
  int32_t B, H, Nq, S, S_pitch;
  int32_t q_col0, do_col0, k_col0, v_col0, dk_col0, dv_col0, dq_col0;
  int32_t tiles_per_chunk, atomic_dq;
  float q_scale;
  uint32_t drop_thresh;        // attention-probability dropout of the forward (0 = none): same (seed, site, element)
  float drop_scale;
  const uint32_t* seed;
  uint32_t site;
};

__device__ __forceinline__ float bwd_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MN-major operand: rows of 64 elements (128 B) along M/N, one row per k, 8-row groups every 1024 B (SBO);
// a second 64-element block of the M/N extent sits `lbo_bytes` further (canonical ((8,n),(8,k)):((1,LBO),(8,SBO))).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// QMASK: per-query mask bits (mask_q_stride != 0); BIAS: additive score bias; DROP: probability dropout.  The common
// cross-attention case (key padding only) compiles to ~7 instructions per score.
template <bool QMASK, bool BIAS, bool DROP>
__global__ void __launch_bounds__(kBwdThreads, 1)
attention_bwd_kernel(const __grid_constant__ AttnBwdMaps maps, const AttnBwdParams p) {
  __shared__ float2 s_stat[128];     // per query: {m + log2(l), ln2 * delta}; +inf reference for absent rows
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + kTileBytes;
  uint8_t* sKV = sdO + kTileBytes;                       // stage s: K at s*32K, V at s*32K + 16K
  uint8_t* sPt = sKV + kBwdStages * 2 * kTileBytes;
  uint8_t* sdSt = sPt + kSqBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdSt + kSqBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + kBwdStages;
  uint64_t* s_full = kv_empty + kBwdStages;
  uint64_t* s_empty = s_full + 1;
  uint64_t* p_full = s_empty + 1;
  uint64_t* p_empty = p_full + 1;
  uint64_t* dkv_full = p_empty + 1;
  uint64_t* dkv_empty = dkv_full + 1;
  uint64_t* dq_full = dkv_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_full + 1);

  const int warp = threadIdx.x >> 5;
  const int h = blockIdx.x, b = blockIdx.y, chunk = blockIdx.z;
  const int tiles_total = (p.S + 127) / 128;
  const int t0 = chunk * p.tiles_per_chunk;
  const int Tn = min(p.tiles_per_chunk, tiles_total - t0);

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.d_o);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
    mbar_init(q_full, 1);
    for (int s = 0; s < kBwdStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, kBwdSoftmaxWarps);
    mbar_init(p_full, kBwdSoftmaxWarps);
    mbar_init(p_empty, 1);
    mbar_init(dkv_full, 1);
    mbar_init(dkv_empty, kBwdSoftmaxWarps);
    mbar_init(dq_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 128, tmem_dV = tmem_base + 256, tmem_dK = tmem_base + 320,
                 tmem_dQ = tmem_base + 384;
  pdl_sync();

  if (Tn > 0) {
    if (warp == 0) {
      if (elect_one()) {
        // ---------------- TMA producer
        mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
        tma_load_3d(sQ, &maps.q, q_full, p.q_col0 + h * 64, 0, b);
        tma_load_3d(sdO, &maps.d_o, q_full, p.do_col0 + h * 64, 0, b);
        for (int i = 0; i < Tn; ++i) {
          const int s = i % kBwdStages;
          mbar_wait(&kv_empty[s], ((i / kBwdStages) & 1) ^ 1, 500 + s);
          uint8_t* sk = sKV + s * 2 * kTileBytes;
          mbar_arrive_expect_tx(&kv_full[s], 2 * kTileBytes);
          tma_load_3d(sk, &maps.k, &kv_full[s], p.k_col0 + h * 64, (t0 + i) * 128, b);
          tma_load_3d(sk + kTileBytes, &maps.v, &kv_full[s], p.v_col0 + h * 64, (t0 + i) * 128, b);
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        // ---------------- MMA issuer
        constexpr uint32_t idesc_sq = umma_idesc_bf16(128, 128);                              // A, B K-major
        constexpr uint32_t idesc_bmn = umma_idesc_bf16(128, 64) | (1u << 16);                 // B MN-major
        constexpr uint32_t idesc_abmn = umma_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);   // A and B MN-major
        const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sdO), pt_addr = smem_u32(sPt), dst_addr = smem_u32(sdSt);
        mbar_wait(q_full, 0, 510);
        for (int i = 0; i < Tn; ++i) {
          const int s = i % kBwdStages;
          const uint32_t k_addr = smem_u32(sKV + s * 2 * kTileBytes), v_addr = k_addr + kTileBytes;
          mbar_wait(&kv_full[s], (i / kBwdStages) & 1, 520 + s);
          mbar_wait(s_empty, (i & 1) ^ 1, 530);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)          // S^T = K Q2^T   (contraction over the 64 head features)
            umma_ss(tmem_S, umma_desc_k_sw128(k_addr + k * 32), umma_desc_k_sw128(q_addr + k * 32), idesc_sq, k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)          // dP^T = V dO^T
            umma_ss(tmem_dP, umma_desc_k_sw128(v_addr + k * 32), umma_desc_k_sw128(do_addr + k * 32), idesc_sq, k != 0);
          tc_commit(s_full);
          mbar_wait(p_full, i & 1, 540);
          mbar_wait(dkv_empty, (i & 1) ^ 1, 550);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 8; ++k) {        // contraction over the 128 queries, 16 per instruction
            const uint32_t a_off = (k >> 2) * (kSqBytes / 2) + (k & 3) * 32;
            umma_ss(tmem_dV, umma_desc_k_sw128(pt_addr + a_off), umma_desc_mn_sw128(do_addr + k * 2048, 0), idesc_bmn,
                    k != 0);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t a_off = (k >> 2) * (kSqBytes / 2) + (k & 3) * 32;
            umma_ss(tmem_dK, umma_desc_k_sw128(dst_addr + a_off), umma_desc_mn_sw128(q_addr + k * 2048, 0), idesc_bmn,
                    k != 0);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k)          // dQ += dS K   (contraction over the tile's 128 keys)
            umma_ss(tmem_dQ, umma_desc_mn_sw128(dst_addr + k * 2048, kSqBytes / 2), umma_desc_mn_sw128(k_addr + k * 2048, 0),
                    idesc_abmn, (i | k) != 0);
          tc_commit(&kv_empty[s]);
          tc_commit(p_empty);
          tc_commit(dkv_full);
        }
        tc_commit(dq_full);
      }
    } else {
      // ---------------- softmax / drain warps
      const int quad = warp & 3;
      const int half = (warp - 2) >> 2;
      const int r = quad * 32 + static_cast<int>(lane_id());          // key row in the tile == TMEM lane; later query row
      const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
      const int64_t bh = static_cast<int64_t>(b) * p.H + h;
      if (half == 0) {
        // P = ex2(S2 - m) / l = ex2(S2 - (m + log2 l));  dS2 = P * (ln2 * dP - ln2 * delta)
        float m2 = __int_as_float(0x7f800000), d2 = 0.f;
        if (r < p.Nq) {
          const float l = p.stat_l[bh * p.Nq + r];
          if (l > 0.f) m2 = p.stat_m[bh * p.Nq + r] + __log2f(l);
          d2 = 0.6931471805599453f * p.delta[bh * p.Nq + r];
        }
        s_stat[r] = make_float2(m2, d2);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const uint32_t dkey = DROP ? drop_key(__ldg(p.seed), p.site) : 0u;
      const uint32_t s_pad = static_cast<uint32_t>(tiles_total * 128);
      const uint32_t* mbase =
          p.mask_bits == nullptr ? nullptr : p.mask_bits + b * p.mask_b_stride + h * p.mask_h_stride;
#pragma unroll 1
      for (int i = 0; i < Tn; ++i) {
        const int t = t0 + i;
        const int kk = t * 128 + r;
        const bool key_ok = kk < p.S;
        bool key_masked = !key_ok;
        if (!QMASK && mbase != nullptr) key_masked = key_masked || ((__ldg(mbase + t * 4 + quad) >> lane_id()) & 1u);
        mbar_wait(s_full, i & 1, 600);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int col0 = half * 64 + c * 32;
          uint32_t sv[32], dv[32];
          tmem_ld_32x32(tmem_S + lane_off + col0, sv);
          tmem_ld_32x32(tmem_dP + lane_off + col0, dv);
          uint32_t qword = 0;             // QMASK: lane j holds the mask word of query col0 + j for this warp's 32 keys
          if (QMASK) {
            const int nq = col0 + static_cast<int>(lane_id());
            qword = nq < p.Nq ? __ldg(mbase + nq * p.mask_q_stride + t * 4 + quad) : 0xffffffffu;
          }
          tmem_ld_wait();
          float* pr = reinterpret_cast<float*>(sv);      // results overwrite their inputs
          float* ds = reinterpret_cast<float*>(dv);
          if (!QMASK && key_masked) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { pr[j] = 0.f; ds[j] = 0.f; }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = col0 + j;
              const float2 st = s_stat[n];
              float sc = __uint_as_float(sv[j]);
              if (BIAS) {
                if (n < p.Nq && key_ok) sc += __ldg(p.bias + (bh * p.Nq + n) * p.bias_ld + kk);
              }
              float pv = bwd_ex2(sc - st.x);
              if (QMASK) {
                const uint32_t w = __shfl_sync(0xffffffffu, qword, j);
                if (key_masked || ((w >> lane_id()) & 1u)) pv = 0.f;
              }
              float dpv = __uint_as_float(dv[j]);
              if (DROP) {
                const int nc = n < p.Nq ? n : p.Nq - 1;
                const uint32_t e = static_cast<uint32_t>(bh * p.Nq + nc) * s_pad + static_cast<uint32_t>(kk);
                const float ks = drop_keep(dkey, e, p.drop_thresh) ? p.drop_scale : 0.f;
                pr[j] = pv * ks;
                dpv *= ks;
              } else {
                pr[j] = pv;
              }
              ds[j] = pv * fmaf(dpv, 0.6931471805599453f, -st.y);
            }
          }
          if (p.dS_out != nullptr && kk < p.ds_ld) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = col0 + j;
              if (n < p.Nq) p.dS_out[(bh * p.Nq + n) * p.ds_ld + kk] = __float2bfloat16_rn(ds[j]);
            }
          }
          if (c == 0) mbar_wait(p_empty, (i & 1) ^ 1, 610);     // the previous tile's MMAs have read P^T / dS^T
          uint8_t* prow = sPt + half * (kSqBytes / 2) + r * 128;
          uint8_t* drow = sdSt + half * (kSqBytes / 2) + r * 128;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint4 u, w;
            u.x = pack_bf16x2(pr[q4 * 8 + 0], pr[q4 * 8 + 1]); u.y = pack_bf16x2(pr[q4 * 8 + 2], pr[q4 * 8 + 3]);
            u.z = pack_bf16x2(pr[q4 * 8 + 4], pr[q4 * 8 + 5]); u.w = pack_bf16x2(pr[q4 * 8 + 6], pr[q4 * 8 + 7]);
            w.x = pack_bf16x2(ds[q4 * 8 + 0], ds[q4 * 8 + 1]); w.y = pack_bf16x2(ds[q4 * 8 + 2], ds[q4 * 8 + 3]);
            w.z = pack_bf16x2(ds[q4 * 8 + 4], ds[q4 * 8 + 5]); w.w = pack_bf16x2(ds[q4 * 8 + 6], ds[q4 * 8 + 7]);
            const int ci = c * 4 + q4;                    // 16-byte chunk inside this atom column's 128-B row
            *reinterpret_cast<uint4*>(prow + ((ci ^ (r & 7)) << 4)) = u;
            *reinterpret_cast<uint4*>(drow + ((ci ^ (r & 7)) << 4)) = w;
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane_id() == 0) {
          mbar_arrive(s_empty);
          mbar_arrive(p_full);
        }
        // drain this tile's dV (half 0) / dK (half 1): one 128-byte row per thread
        mbar_wait(dkv_full, i & 1, 620);
        tc_fence_after();
        {
          uint32_t a0[32], a1[32];
          const uint32_t src = (half == 0 ? tmem_dV : tmem_dK) + lane_off;
          tmem_ld_32x32(src, a0);
          tmem_ld_32x32(src + 32, a1);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane_id() == 0) mbar_arrive(dkv_empty);
          if (key_ok) {
            __nv_bfloat16* dst = half == 0
                ? p.dV + (static_cast<int64_t>(b) * p.S_pitch + kk) * p.ld_dv + p.dv_col0 + h * 64
                : p.dK + (static_cast<int64_t>(b) * p.S_pitch + kk) * p.ld_dk + p.dk_col0 + h * 64;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(a0[j]), __uint_as_float(a0[j + 1]));
              u.y = pack_bf16x2(__uint_as_float(a0[j + 2]), __uint_as_float(a0[j + 3]));
              u.z = pack_bf16x2(__uint_as_float(a0[j + 4]), __uint_as_float(a0[j + 5]));
              u.w = pack_bf16x2(__uint_as_float(a0[j + 6]), __uint_as_float(a0[j + 7]));
              *reinterpret_cast<uint4*>(dst + j) = u;
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(a1[j]), __uint_as_float(a1[j + 1]));
              w.y = pack_bf16x2(__uint_as_float(a1[j + 2]), __uint_as_float(a1[j + 3]));
              w.z = pack_bf16x2(__uint_as_float(a1[j + 4]), __uint_as_float(a1[j + 5]));
              w.w = pack_bf16x2(__uint_as_float(a1[j + 6]), __uint_as_float(a1[j + 7]));
              *reinterpret_cast<uint4*>(dst + 32 + j) = w;
            }
          }
        }
      }
      // ---- dQ: rows are queries now; this warp's 32 of the 64 feature columns
      mbar_wait(dq_full, 0, 630);
      tc_fence_after();
      {
        uint32_t a[32];
        tmem_ld_32x32(tmem_dQ + lane_off + half * 32, a);
        tmem_ld_wait();
        if (r < p.Nq) {
          float* dst = p.dQ + (static_cast<int64_t>(b) * p.Nq + r) * p.ld_dq + p.dq_col0 + h * 64 + half * 32;
          if (p.atomic_dq) {
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + j, __uint_as_float(a[j]) * p.q_scale);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) =
                  make_float4(__uint_as_float(a[j]) * p.q_scale, __uint_as_float(a[j + 1]) * p.q_scale,
                              __uint_as_float(a[j + 2]) * p.q_scale, __uint_as_float(a[j + 3]) * p.q_scale);
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace pq3d

using namespace pq3d;

extern "C" int pq3d_attention_bwd(const void* Q, int64_t ldq, int q_col0, const void* dO, int64_t lddo, int do_col0,
                                  const void* K, int64_t ldk, int k_col0, const void* V, int64_t ldv, int v_col0, int S,
                                  int S_pitch, const uint32_t* mask_bits, int64_t mask_b_stride, int64_t mask_h_stride,
                                  int64_t mask_q_stride, const float* bias, int64_t bias_ld, const float* stat_m,
                                  const float* stat_l, const float* delta, void* dK, int64_t ld_dk, int dk_col0, void* dV,
                                  int64_t ld_dv, int dv_col0, float* dQ, int64_t ld_dq, int dq_col0, void* dS_out,
                                  int64_t ds_ld, int B, int H, int Nq, float q_scale, float drop_p,
                                  const uint32_t* seed_dev, uint32_t site, void* stream) {
  PQ3D_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || seed_dev != nullptr),
                 "pq3d_attention_bwd: dropout needs p in [0,1) and a device seed");
  PQ3D_CHECK_ARG(Q && dO && K && V && stat_m && stat_l && delta && dK && dV && dQ, "pq3d_attention_bwd: null argument");
  PQ3D_CHECK_ARG(B > 0 && H > 0 && Nq > 0 && Nq <= 128 && S > 0 && S_pitch >= S,
                 "pq3d_attention_bwd: bad shape B=%d H=%d Nq=%d (<= 128) S=%d S_pitch=%d", B, H, Nq, S, S_pitch);
  auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  PQ3D_CHECK_ARG(ldq % 8 == 0 && lddo % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ld_dk % 8 == 0 && ld_dv % 8 == 0 &&
                     ld_dq % 4 == 0 && q_col0 % 8 == 0 && do_col0 % 8 == 0 && k_col0 % 8 == 0 && v_col0 % 8 == 0 &&
                     dk_col0 % 8 == 0 && dv_col0 % 8 == 0 && dq_col0 % 4 == 0 && a16(Q) && a16(dO) && a16(K) && a16(V) &&
                     a16(dK) && a16(dV) && a16(dQ),
                 "pq3d_attention_bwd: operands must be 16-byte aligned with 16-byte granular pitches / column offsets");
  AttnBwdMaps maps;
  auto rows_map = [&](CUtensorMap* m, const void* base, int64_t ld, int rows, int64_t pitch_rows) -> int {
    uint64_t dims[3] = {(uint64_t)ld, (uint64_t)rows, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)pitch_rows * ld * 2};
    uint32_t box[3] = {64u, 128u, 1u};
    return make_tmap_bf16(m, base, 3, dims, strides, box);
  };
  int rc;
  if ((rc = rows_map(&maps.q, Q, ldq, Nq, Nq)) != PQ3D_OK) return rc;
  if ((rc = rows_map(&maps.d_o, dO, lddo, Nq, Nq)) != PQ3D_OK) return rc;
  if ((rc = rows_map(&maps.k, K, ldk, S, S_pitch)) != PQ3D_OK) return rc;
  if ((rc = rows_map(&maps.v, V, ldv, S, S_pitch)) != PQ3D_OK) return rc;
  AttnBwdParams p;
  memset(&p, 0, sizeof(p));
  p.mask_bits = mask_bits;
  p.mask_b_stride = mask_b_stride; p.mask_h_stride = mask_h_stride; p.mask_q_stride = mask_q_stride;
  p.bias = bias; p.bias_ld = bias_ld;
  p.stat_m = stat_m; p.stat_l = stat_l; p.delta = delta;
  p.dK = reinterpret_cast<__nv_bfloat16*>(dK); p.dV = reinterpret_cast<__nv_bfloat16*>(dV);
  p.ld_dk = ld_dk; p.ld_dv = ld_dv;
  p.dQ = dQ; p.ld_dq = ld_dq;
  p.dS_out = reinterpret_cast<__nv_bfloat16*>(dS_out); p.ds_ld = ds_ld;
  p.B = B; p.H = H; p.Nq = Nq; p.S = S; p.S_pitch = S_pitch;
  p.q_col0 = q_col0; p.do_col0 = do_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.dk_col0 = dk_col0; p.dv_col0 = dv_col0; p.dq_col0 = dq_col0;
  p.q_scale = q_scale;
  if (drop_p > 0.f) {
    p.drop_thresh = drop_threshold(drop_p);
    p.drop_scale = 1.f / (1.f - drop_p);
    p.seed = seed_dev;
    p.site = site;
  }
  // split the key range over enough CTAs to fill the device; dQ then meets in fp32 atomics (caller zero-fills dQ's
  // [*, dq_col0 + H*64) block in that case — see the return value)
  const int tiles = (S + 127) / 128;
  int chunks = (sm_count() + B * H - 1) / (B * H);
  if (chunks > tiles) chunks = tiles;
  if (chunks < 1) chunks = 1;
  p.tiles_per_chunk = (tiles + chunks - 1) / chunks;
  chunks = (tiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
  p.atomic_dq = 1;      // dQ is always accumulated: the caller zero-fills it (one memset for all memories of a group)
  const bool qmask = mask_bits != nullptr && mask_q_stride != 0, has_bias = bias != nullptr, drop = drop_p > 0.f;
  const dim3 grid(H, B, chunks), block(kBwdThreads);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t err = cudaSuccess;
#define PQ3D_BWD_CASE(Q_, B_, D_)                                                                                     \
  if (qmask == Q_ && has_bias == B_ && drop == D_) {                                                                  \
    static bool configured = false;                                                                                   \
    if (!configured) {                                                                                                \
      PQ3D_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<Q_, B_, D_>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                     kBwdSmem));                                                                      \
      configured = true;                                                                                              \
    }                                                                                                                 \
    err = launch_kernel(attention_bwd_kernel<Q_, B_, D_>, grid, block, kBwdSmem, st, maps, p);                        \
  }
  PQ3D_BWD_CASE(false, false, false) PQ3D_BWD_CASE(false, false, true) PQ3D_BWD_CASE(false, true, false)
  PQ3D_BWD_CASE(false, true, true) PQ3D_BWD_CASE(true, false, false) PQ3D_BWD_CASE(true, false, true)
  PQ3D_BWD_CASE(true, true, false) PQ3D_BWD_CASE(true, true, true)
#undef PQ3D_BWD_CASE
  PQ3D_CUDA(err);
  return PQ3D_OK;
}
