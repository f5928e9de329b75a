// pq3d_attention_fwd: masked softmax attention of N queries over a scene memory, on tcgen05.
//
// Replaces the core of nn.MultiheadAttention(add_zero_attn=True) inside CrossAttentionLayer
// (modules/grounding/query_encoder.py:288-307 -> torch/nn/functional.py:6585-6647: zero key/value
// appended after projection, float -inf mask, baddbmm, softmax, bmm) and, with an additive score
// bias, MultiHeadAttentionSpatial fusion 'mul' (modules/layers/transformers.py:189-240:
// softmax(log(clamp(relu(W5 . pairwise_locs + b), 1e-6)) + q.k/sqrt(dh)); the bias tensor comes from
// pq3d_spatial_bias).
//
// Inputs are already projected: Q pre-scaled by log2(e)/sqrt(dh) — scores live in the log2 domain so
// a probability is one ex2 —, K as [token, feature] and V TRANSPOSED as [feature, token] (the
// projection GEMM writes V^T directly by swapping its operands), so every MMA operand is K-major and
// TMA-loadable with the 128-byte swizzle.
//
// One CTA = (head, scene, memory, 128-query tile).  576 threads:
//   warp 0      TMA producer: Q tile once, then K / K + V^T tiles of 128 keys into a 5-stage ring
//   warp 1      tcgen05.mma issuer: S = Q K^T (128x128, fp32 in TMEM, double-buffered),
//               O += P V (128x64 in TMEM) in the TS form: the A operand P is read from TENSOR MEMORY
//               (bf16 pairs, 64 columns per 128-key tile, double-buffered), V^T from shared memory.
//               The softmax warps write P with tcgen05.st, so probabilities never touch shared memory.
//   warps 2..17 softmax: 16 warps, four per TMEM lane quadrant, each owning one 32-key column chunk
//               of every 128-key tile for its 32 query rows (TMEM lane = row); per-row partials are
//               combined through shared memory once per pass.
// Three schedules:
//   ONE_PASS   (add_zero_attn, > 256 keys — the scene memories): the zero-attn key pins a score of 0
//              into every row, so 0 is a natural softmax reference: p = ex2(s) with NO running max, no
//              rescale of O and a single sweep over K/V.  Scores far below 0 underflow to the weight
//              they deserve next to the zero-attn key; only very large scores could overflow, which is
//              detected on the row sums (l > 2^100) and then the CTA redoes the tile loop as TWO_PASS.
//   TWO_PASS   row maxima first (QK^T only), then p = ex2(s - m), row sums and PV.  N <= 128 queries
//              per CTA make the second QK^T cheap on the tensor pipe and O never needs a rescale.
//   RESIDENT   memories of <= 256 keys (language prompt, query self-attention): both score tiles stay
//              in TMEM, QK^T runs once and the second sweep re-reads them.
// add_zero_attn is analytic everywhere: the extra key has score 0 and value 0, so it only adds
// ex2(0 - m) to the denominator; fully masked rows give exactly 0, never NaN.
// Masks arrive bit-packed (1 = ignore), one word per 32 keys; key-padding masks broadcast over rows
// with a zero row stride.  Bits past S are set by the packer.
#include <cstring>

#include "host_common.h"
#include "ptx.cuh"

namespace pq3d {

constexpr int kMaxMem = 4;
constexpr int kSoftmaxWarps = 16;                       // 4 per TMEM lane quadrant, one 32-key chunk each
constexpr int kAttnThreads = 64 + 32 * kSoftmaxWarps;
constexpr int kKvTile = 128;
constexpr int kKvStages = 5;      // no P tiles in shared memory any more: the ring gets their space
constexpr int kHeadDim = 64;

enum AttnMode : int { kResident = 0, kTwoPass = 1, kOnePass = 2 };

struct AttnMaps {
  CUtensorMap q;
  CUtensorMap k[kMaxMem];
  CUtensorMap vt[kMaxMem];
};

struct AttnMem {
  const uint32_t* mask_bits;  // may be null (nothing masked; tail handled by S)
  const int32_t* kv_tiles;    // optional [B]: number of leading 128-key tiles that hold a visible key
  int64_t mask_b_stride, mask_h_stride, mask_q_stride;  // in 32-bit words
  int32_t S;
  int32_t k_col0;
  int32_t vt_row0;
  int32_t pad_;
};

struct AttnParams {
  AttnMem mem[kMaxMem];
  __nv_bfloat16* O;
  int64_t ldo, o_mem_stride;
  int32_t q_mem_stride;
  int32_t B, H, Nq, q_tiles;
  int32_t zero_attn;
  int32_t force_two_pass;      // testing hook: never take the ONE_PASS schedule
  const float* score_bias;     // [B, H, Nq, bias_ld] fp32 (log2 domain) added to the scores, or null
  int64_t bias_ld;
  float *stat_m, *stat_l;      // optional [n_mem][B][H][Nq]: softmax reference and denominator, saved for the backward
  int64_t stat_mem_stride;
  // training only: dropout on the attention probabilities (nn.MultiheadAttention(dropout=p), applied after the
  // softmax normalisation); element ((b*H + h)*Nq + n)*ceil128(S) + key of stream site[mem] (csrc/ptx.cuh RNG)
  uint32_t drop_thresh;
  float drop_scale;
  const uint32_t* seed;
  uint32_t site[kMaxMem];
  unsigned long long* dbg;
};

constexpr int kQBytes = 128 * kHeadDim * 2;           // 16 KB
constexpr int kKBytes = kKvTile * kHeadDim * 2;       // 16 KB
constexpr int kVBytes = kHeadDim * kKvTile * 2;       // 16 KB (two [64 x 64] boxes)

constexpr int kNumBars = 1 + 2 * kKvStages + 8 + 1;   // q_full, kv_full/empty, s_full/empty, p_full/empty, o_full
constexpr int kAttnSmem = kQBytes + kKvStages * (kKBytes + kVBytes) + 256 + 2048 /*row partials*/;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnCtx {
  uint8_t *sQ, *sKV;
  uint64_t *q_full, *kv_full, *kv_empty, *s_full, *s_empty, *p_full, *p_empty, *o_full;
  float* s_part;
  uint32_t tmem_S0, tmem_O, tmem_P;
  int h, b, mi, qt, T;
};

__device__ __forceinline__ void init_barriers(const AttnCtx& c) {
  mbar_init(c.q_full, 1);
  for (int s = 0; s < kKvStages; ++s) {
    mbar_init(&c.kv_full[s], 1);
    mbar_init(&c.kv_empty[s], 1);
  }
  for (int i = 0; i < 2; ++i) {
    mbar_init(&c.s_full[i], 1);
    mbar_init(&c.s_empty[i], kSoftmaxWarps);
    mbar_init(&c.p_full[i], kSoftmaxWarps);
    mbar_init(&c.p_empty[i], 1);
  }
  mbar_init(c.o_full, 1);
  fence_mbar_init();
}

// ---------------------------------------------------------------------------------------------------
// TMA producer (one elected lane)
// ---------------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void attn_producer(const AttnMaps& maps, const AttnParams& p, const AttnMem& mem,
                                              const AttnCtx& c) {
  mbar_arrive_expect_tx(c.q_full, kQBytes);
  tma_load_3d(c.sQ, &maps.q, c.q_full, c.mi * p.q_mem_stride + c.h * kHeadDim, c.qt * 128, c.b);
  int it = 0;
  for (int pass = (MODE == kTwoPass ? 0 : 1); pass < 2; ++pass) {
    for (int t = 0; t < c.T; ++t, ++it) {
      const int s = it % kKvStages;
      mbar_wait(&c.kv_empty[s], ((it / kKvStages) & 1) ^ 1, 100 + s);
      uint8_t* sk = c.sKV + s * (kKBytes + kVBytes);
      mbar_arrive_expect_tx(&c.kv_full[s], pass == 0 ? kKBytes : kKBytes + kVBytes);
      tma_load_3d(sk, &maps.k[c.mi], &c.kv_full[s], mem.k_col0 + c.h * kHeadDim, t * kKvTile, c.b);
      if (pass == 1) {
        tma_load_3d(sk + kKBytes, &maps.vt[c.mi], &c.kv_full[s], t * kKvTile, c.b, mem.vt_row0 + c.h * kHeadDim);
        tma_load_3d(sk + kKBytes + kVBytes / 2, &maps.vt[c.mi], &c.kv_full[s], t * kKvTile + 64, c.b,
                    mem.vt_row0 + c.h * kHeadDim);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// MMA issuer (one elected lane)
// ---------------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void attn_mma(const AttnCtx& c) {
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, kKvTile);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim);
  const uint32_t q_addr = smem_u32(c.sQ);
  int it = 0, g = 0;
  auto issue_s = [&](bool release_kv) {
    const int s = it % kKvStages;
    mbar_wait(&c.kv_full[s], (it / kKvStages) & 1, 200 + s);
    const int sb = g & 1;
    if (MODE != kResident) mbar_wait(&c.s_empty[sb], ((g >> 1) & 1) ^ 1, 210 + sb);
    tc_fence_after();
    const uint32_t k_addr = smem_u32(c.sKV + s * (kKBytes + kVBytes));
#pragma unroll
    for (int k = 0; k < kHeadDim / 16; ++k)
      umma_ss(c.tmem_S0 + sb * kKvTile, umma_desc_k_sw128(q_addr + k * 32), umma_desc_k_sw128(k_addr + k * 32),
              idesc_s, k != 0 ? 1u : 0u);
    if (release_kv) tc_commit(&c.kv_empty[s]);
    tc_commit(&c.s_full[sb]);
    ++it;
    ++g;
  };
  auto issue_pv = [&](int t, int ring_pos) {
    const int pb = t & 1;
    mbar_wait(&c.p_full[pb], (t >> 1) & 1, 230 + pb);
    tc_fence_after();
    const int s = ring_pos % kKvStages;
    const uint32_t v_addr = smem_u32(c.sKV + s * (kKBytes + kVBytes) + kKBytes);
    const uint32_t p_tmem = c.tmem_P + pb * (kKvTile / 2);   // P tile: 128 lanes x 128 bf16 = 64 columns, A operand in TMEM
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_ts(c.tmem_O, p_tmem + (j * 4 + k) * 8, umma_desc_k_sw128(v_addr + j * (kVBytes / 2) + k * 32), idesc_o,
                (t | j | k) != 0 ? 1u : 0u);
    }
    tc_commit(&c.kv_empty[s]);
    tc_commit(&c.p_empty[pb]);
  };
  mbar_wait(c.q_full, 0, 220);
  if (MODE == kResident) {
    for (int t = 0; t < c.T; ++t) issue_s(false);           // scores once, both tiles stay in TMEM
    for (int t = 0; t < c.T; ++t) issue_pv(t, t);
  } else {
    if (MODE == kTwoPass)
      for (int t = 0; t < c.T; ++t) issue_s(true);          // pass 1: scores only
    const int it2 = it;                                     // ring position of the PV sweep's tile 0
    issue_s(false);
    for (int t = 0; t < c.T; ++t) {
      if (t + 1 < c.T) issue_s(false);                      // keep one QK^T ahead of the softmax
      issue_pv(t, it2 + t);
    }
  }
  tc_commit(c.o_full);
}

// ---------------------------------------------------------------------------------------------------
// softmax warps.  Returns true when the ONE_PASS row sums overflowed their safe range.
// ---------------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ bool attn_softmax(const AttnParams& p, const AttnMem& mem, const AttnCtx& c, int warp,
                                             unsigned long long* dbg) {
  const int quad = warp & 3;
  const int cc = (warp - 2) >> 2;               // which 32-key chunk of every tile this warp owns
  const int r = quad * 32 + lane_id();          // row in the query tile == TMEM lane
  const int n = c.qt * 128 + r;                 // query index in the scene
  const int n_c = n < p.Nq ? n : p.Nq - 1;      // clamped for reads
  const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
  const uint32_t* mrow = mem.mask_bits == nullptr ? nullptr
                                                  : mem.mask_bits + c.b * mem.mask_b_stride +
                                                        c.h * mem.mask_h_stride + n_c * mem.mask_q_stride;
  const float* brow = p.score_bias == nullptr
                          ? nullptr
                          : p.score_bias + ((static_cast<int64_t>(c.b) * p.H + c.h) * p.Nq + n_c) * p.bias_ld;
  const float kNegInf = __int_as_float(0xff800000);

  auto mask_word = [&](int t) -> uint32_t {
    if (mrow != nullptr) return __ldg(mrow + t * 4 + cc);
    const int rem = mem.S - (t * kKvTile + cc * 32);       // no mask tensor: only the tail past S is ignored
    return rem >= 32 ? 0u : (rem <= 0 ? 0xffffffffu : (0xffffffffu << rem));
  };
  // this warp's 32 scores of tile t (buffer sb): TMEM -> registers (+ bias); masked keys -> -inf.
  // The select is skipped when no lane of the warp has a masked key in this chunk (the common case).
  auto load_scores = [&](int t, int sb, uint32_t word, float (&sc)[32]) {
    uint32_t acc[32];
    tmem_ld_32x32(c.tmem_S0 + sb * kKvTile + lane_off + cc * 32, acc);
    const bool any_masked = __any_sync(0xffffffffu, word != 0u);
    if (brow != nullptr) {
      const float4* b4 = reinterpret_cast<const float4*>(brow + t * kKvTile + cc * 32);
      float bias[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 v = __ldg(b4 + j);
        bias[4 * j] = v.x; bias[4 * j + 1] = v.y; bias[4 * j + 2] = v.z; bias[4 * j + 3] = v.w;
      }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) sc[j] = __uint_as_float(acc[j]) + bias[j];
    } else {
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) sc[j] = __uint_as_float(acc[j]);
    }
    if (any_masked) {
#pragma unroll
      for (int j = 0; j < 32; ++j) sc[j] = ((word >> j) & 1u) ? kNegInf : sc[j];
    }
  };
  auto release_scores = [&](int sb) {
    tc_fence_before();
    __syncwarp();
    if (lane_id() == 0) mbar_arrive(&c.s_empty[sb]);
  };
  int g = 0;
  float m = 0.f;                                 // ONE_PASS: the zero-attn score is the reference
  if (MODE != kOnePass) {
    float m_run = kNegInf;
    for (int t = 0; t < c.T; ++t, ++g) {         // ---- sweep 1: row maxima (partial: this warp's columns)
      const int sb = g & 1;
      const uint32_t word = mask_word(t);
      mbar_wait(&c.s_full[sb], (g >> 1) & 1, 300 + sb);
      tc_fence_after();
      float sc[32];
      load_scores(t, sb, word, sc);
      if (MODE == kTwoPass) release_scores(sb);  // values are in registers: free the buffer early
#pragma unroll
      for (int j = 0; j < 32; ++j) m_run = fmaxf(m_run, sc[j]);
    }
    c.s_part[cc * 128 + r] = m_run;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    m_run = fmaxf(fmaxf(c.s_part[r], c.s_part[128 + r]), fmaxf(c.s_part[256 + r], c.s_part[384 + r]));
    asm volatile("bar.sync 1, 512;" ::: "memory");        // s_part is reused for the row sums
    m = p.zero_attn ? fmaxf(m_run, 0.f) : m_run;
    if (m == kNegInf) m = 0.f;
    if (MODE == kResident) g = 0;                // sweep 2 re-reads the resident score tiles
  }
  if (dbg != nullptr && threadIdx.x == 64) dbg[2] = clock64();
  const uint32_t dkey = p.drop_thresh != 0 ? drop_key(__ldg(p.seed), p.site[c.mi]) : 0u;
  const uint32_t drow = static_cast<uint32_t>(((static_cast<int64_t>(c.b) * p.H + c.h) * p.Nq + n_c) *
                                              (((mem.S + kKvTile - 1) / kKvTile) * kKvTile));
  // ---- probabilities, row sums, P tiles.  Each warp's 32-key chunk is processed as two HALVES of 16 keys: the TMEM load
  // of the next half is in flight while ex2 / packing work on the current one, and only 2 x 16 score registers are live
  // (the kernel is capped at 96 registers per thread: 18 warps).  Scores land in float registers directly, the row sum
  // runs as four interleaved partial sums, mask words are fetched one tile ahead.
  float l4[4] = {0.f, 0.f, 0.f, 0.f};
  {
    float bufA[16], bufB[16];
    const int g0 = g;
    auto s_buffer = [&](int t) { return (g0 + t) & 1; };
    auto wait_scores = [&](int t) {
      if (MODE != kResident) {
        mbar_wait(&c.s_full[s_buffer(t)], ((g0 + t) >> 1) & 1, 310 + s_buffer(t));
        tc_fence_after();
      }
    };
    auto issue = [&](int t, int h, float (&buf)[16]) {
      tmem_ld_32x16f(c.tmem_S0 + s_buffer(t) * kKvTile + lane_off + cc * 32 + h * 16, buf);
    };
    // 16 scores -> 16 probabilities (bias, mask, ex2, row sum, dropout) -> two 16-byte chunks of the swizzled P tile
    auto process_half = [&](int t, int h, float (&pr)[16], uint32_t bits, bool any_masked, int pb) {
      if (brow != nullptr) {
        const float4* b4 = reinterpret_cast<const float4*>(brow + t * kKvTile + cc * 32 + h * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = __ldg(b4 + j);
          pr[4 * j] += v.x; pr[4 * j + 1] += v.y; pr[4 * j + 2] += v.z; pr[4 * j + 3] += v.w;
        }
      }
      if (any_masked) {
#pragma unroll
        for (int j = 0; j < 16; ++j) pr[j] = ((bits >> j) & 1u) ? kNegInf : pr[j];
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        pr[j] = ex2_approx(MODE == kOnePass ? pr[j] : pr[j] - m);    // ex2(-inf) = 0 for masked keys
        l4[j & 3] += pr[j];
      }
      if (p.drop_thresh != 0) {                  // the denominator keeps every key; P V uses the dropped, rescaled weights
        const uint32_t e0 = drow + t * kKvTile + cc * 32 + h * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) pr[j] = drop_keep(dkey, e0 + j, p.drop_thresh) ? pr[j] * p.drop_scale : 0.f;
      }
      // 16 probabilities -> 8 packed bf16 pairs -> 8 TMEM columns of this row (keys cc*32 + h*16 ..): P never touches
      // shared memory; the P.V product reads it as its tensor-memory A operand
      uint32_t u[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = pack_bf16x2(pr[2 * j], pr[2 * j + 1]);
      tmem_st_32x8(c.tmem_P + pb * (kKvTile / 2) + lane_off + cc * 16 + h * 8, u);
    };
    uint32_t word_next = mask_word(0);
    wait_scores(0);
    issue(0, 0, bufA);
    for (int t = 0; t < c.T; ++t, ++g) {
      const int pb = t & 1;
      const uint32_t word = word_next;
      if (t + 1 < c.T) word_next = mask_word(t + 1);
      const bool any_masked = __any_sync(0xffffffffu, word != 0u);
      tmem_ld_wait();                            // half 0 of tile t is in bufA
      issue(t, 1, bufB);
      mbar_wait(&c.p_empty[pb], ((t >> 1) & 1) ^ 1, 320 + pb);
      process_half(t, 0, bufA, word & 0xffffu, any_masked, pb);
      tmem_ld_wait();                            // half 1 in bufB: the score buffer can go back to the MMA warp
      if (MODE != kResident) release_scores(s_buffer(t));
      if (t + 1 < c.T) {
        wait_scores(t + 1);
        issue(t + 1, 0, bufA);
      }
      process_half(t, 1, bufB, word >> 16, any_masked, pb);
      tmem_st_wait();                            // this warp's part of the P tile is in tensor memory
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&c.p_full[pb]);
    }
  }
  float l = (l4[0] + l4[1]) + (l4[2] + l4[3]);
  if (dbg != nullptr && threadIdx.x == 64) dbg[3] = clock64();
  // ---- finalize: combine the partial row sums, zero-attn column, normalise, store
  c.s_part[cc * 128 + r] = l;
  asm volatile("bar.sync 1, 512;" ::: "memory");
  l = (c.s_part[r] + c.s_part[128 + r]) + (c.s_part[256 + r] + c.s_part[384 + r]);
  const bool overflow = (MODE == kOnePass) && !(l < 1.2676506e30f);   // 2^100; also catches inf / NaN
  if (p.zero_attn) l += ex2_approx(-m);
  const float inv = 1.f / l;
  if (p.stat_m != nullptr && cc == 0 && n < p.Nq) {
    const int64_t si = c.mi * p.stat_mem_stride + (static_cast<int64_t>(c.b) * p.H + c.h) * p.Nq + n;
    p.stat_m[si] = m;
    p.stat_l[si] = l;
  }
  mbar_wait(c.o_full, 0, 330);
  tc_fence_after();
  if (cc < 2) {                                  // 64 output columns: two of the four warps per quadrant
    __nv_bfloat16* orow =
        p.O + c.mi * p.o_mem_stride + (static_cast<int64_t>(c.b) * p.Nq + n_c) * p.ldo + c.h * kHeadDim;
    uint32_t acc[32];
    tmem_ld_32x32(c.tmem_O + lane_off + cc * 32, acc);
    tmem_ld_wait();
    if (n < p.Nq) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(acc[j]) * inv, __uint_as_float(acc[j + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(acc[j + 2]) * inv, __uint_as_float(acc[j + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(acc[j + 4]) * inv, __uint_as_float(acc[j + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(acc[j + 6]) * inv, __uint_as_float(acc[j + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + cc * 32 + j) = u;
      }
    }
  }
  tc_fence_before();
  if (dbg != nullptr && threadIdx.x == 64) dbg[4] = clock64();
  return overflow && n < p.Nq;
}

template <int MODE>
__device__ __forceinline__ bool attn_run(const AttnMaps& maps, const AttnParams& p, const AttnMem& mem,
                                         const AttnCtx& c, int warp, unsigned long long* dbg) {
  bool overflow = false;
  if (warp == 0) {
    if (elect_one()) attn_producer<MODE>(maps, p, mem, c);
  } else if (warp == 1) {
    if (elect_one()) attn_mma<MODE>(c);
  } else {
    overflow = attn_softmax<MODE>(p, mem, c, warp, dbg);
  }
  return overflow;
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attention_fwd_kernel(const __grid_constant__ AttnMaps maps, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  AttnCtx c;
  c.sQ = smem;
  c.sKV = c.sQ + kQBytes;
  uint8_t* after_ring = c.sKV + kKvStages * (kKBytes + kVBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(after_ring);
  c.q_full = bars;
  c.kv_full = bars + 1;
  c.kv_empty = c.kv_full + kKvStages;
  c.s_full = c.kv_empty + kKvStages;
  c.s_empty = c.s_full + 2;
  c.p_full = c.s_empty + 2;
  c.p_empty = c.p_full + 2;
  c.o_full = c.p_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c.o_full + 1);
  uint32_t* redo_flag = tmem_slot + 1;
  c.s_part = reinterpret_cast<float*>(after_ring + 256);   // [4 column chunks][128 rows]

  const int warp = threadIdx.x >> 5;
  c.h = blockIdx.x;
  c.b = blockIdx.y;
  c.mi = blockIdx.z / p.q_tiles;
  c.qt = blockIdx.z % p.q_tiles;
  const AttnMem& mem = p.mem[c.mi];
  c.T = (mem.S + kKvTile - 1) / kKvTile;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("pq3d: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k[c.mi]);
    tma_prefetch_desc(&maps.vt[c.mi]);
    init_barriers(c);
    *redo_flag = 0;
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  c.tmem_S0 = tmem_base;        // columns [0,128) and [128,256): S double buffer
  c.tmem_O = tmem_base + 256;   // columns [256,320): O
  c.tmem_P = tmem_base + 320;   // columns [320,384) and [384,448): P double buffer (bf16 pairs), A operand of P.V
  pdl_sync();
  if (mem.kv_tiles != nullptr) {          // ragged scenes: skip the trailing tiles that are padding for every query
    const int act = __ldg(mem.kv_tiles + c.b);
    c.T = act < 1 ? 1 : (act < c.T ? act : c.T);
  }
  unsigned long long* dbg =
      p.dbg == nullptr ? nullptr : p.dbg + 8ull * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z));
  if (dbg != nullptr && threadIdx.x == 64) {
    dbg[0] = global_timer_ns();
    dbg[1] = clock64();
  }

  if (c.T <= 2) {
    attn_run<kResident>(maps, p, mem, c, warp, dbg);
  } else if (p.zero_attn && !p.force_two_pass) {
    if (attn_run<kOnePass>(maps, p, mem, c, warp, dbg)) *redo_flag = 1;   // benign race: all writers store 1
    __syncthreads();                       // every pipeline of the first attempt has drained
    if (*redo_flag != 0) {                 // some row sum left the safe range: redo with running maxima
      if (threadIdx.x == 0) {
        for (int i = 0; i < kNumBars; ++i)
          asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars[i])) : "memory");
        init_barriers(c);
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      attn_run<kTwoPass>(maps, p, mem, c, warp, dbg);
      if (dbg != nullptr && threadIdx.x == 64) dbg[6] = 1;
    }
  } else {
    attn_run<kTwoPass>(maps, p, mem, c, warp, dbg);
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
  if (dbg != nullptr && threadIdx.x == 64) dbg[5] = clock64();
}

// score bias of the spatial self-attention for every layer at once, in the log2 domain:
// out[l][b][h][n][m] = log2(max(relu(loc[b,n,m,:] . w[l,h,:] + bias[l,h]), 1e-6))
// (modules/layers/transformers.py:196-199,231-232)
__global__ void spatial_bias_kernel(const float* __restrict__ locs, const float* __restrict__ w,
                                    const float* __restrict__ bias, float* __restrict__ out, int L, int B, int H, int N,
                                    int ld) {
  pdl_sync();
  const int64_t pairs = static_cast<int64_t>(B) * N * N;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < pairs;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i % N);
    const int n = static_cast<int>((i / N) % N);
    const int b = static_cast<int>(i / (static_cast<int64_t>(N) * N));
    const float* l5 = locs + i * 5;
    const float l0 = l5[0], l1 = l5[1], l2 = l5[2], l3 = l5[3], l4 = l5[4];
    for (int lh = 0; lh < L * H; ++lh) {
      const float* wv = w + lh * 5;
      const float v = fmaf(l0, wv[0], fmaf(l1, wv[1], fmaf(l2, wv[2], fmaf(l3, wv[3], fmaf(l4, wv[4], bias[lh])))));
      const int l = lh / H, hh = lh % H;
      out[(((static_cast<int64_t>(l) * B + b) * H + hh) * N + n) * ld + m] = __log2f(fmaxf(fmaxf(v, 0.f), 1e-6f));
    }
  }
}

}  // namespace pq3d

using namespace pq3d;

static unsigned long long* g_attn_timeline = nullptr;
static int g_force_two_pass = 0;
extern "C" int pq3d_debug_set_attention_timeline(void* buf) {
  g_attn_timeline = reinterpret_cast<unsigned long long*>(buf);
  return PQ3D_OK;
}
// Testing hook: 1 = always use the running-max (two-pass) schedule for long memories.
extern "C" int pq3d_debug_force_two_pass(int on) {
  g_force_two_pass = on;
  return PQ3D_OK;
}

extern "C" int pq3d_spatial_bias(const float* pairwise_locs, const float* loc_w, const float* loc_b, float* out, int L,
                                 int B, int H, int N, int64_t ld, void* stream) {
  PQ3D_CHECK_ARG(pairwise_locs && loc_w && loc_b && out, "pq3d_spatial_bias: null argument");
  PQ3D_CHECK_ARG(L > 0 && B > 0 && H > 0 && N > 0 && ld >= N && ld % 4 == 0, "pq3d_spatial_bias: bad shape");
  const int64_t pairs = static_cast<int64_t>(B) * N * N;
  int blocks = static_cast<int>((pairs + 255) / 256);
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  PQ3D_CUDA(launch_kernel(spatial_bias_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                          pairwise_locs, loc_w, loc_b, out, L, B, H, N, static_cast<int>(ld)));
  return PQ3D_OK;
}

static int attention_fwd_impl(int n_mem, const void* Q, int64_t ldq, int64_t q_mem_stride, const void* const* K,
                                  const int64_t* ldk, const int64_t* k_col0, const void* const* Vt,
                                  const int64_t* ldvt, const int64_t* vt_row0, const int64_t* vt_rows,
                                  const int32_t* S, const int32_t* S_pitch, const int32_t* Vt_pitch,
                                  const uint32_t* const* mask_bits, const int64_t* mask_b_stride,
                                  const int64_t* mask_h_stride, const int64_t* mask_q_stride,
                                  const int32_t* const* kv_tiles, void* O, int64_t ldo,
                                  int64_t o_mem_stride, int B, int H, int Nq, int zero_attn, const float* score_bias,
                                  int64_t bias_ld, float* stat_m, float* stat_l, int64_t stat_mem_stride,
                                  float drop_p, const uint32_t* seed_dev, const uint32_t* sites, void* stream) {
  PQ3D_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || (seed_dev != nullptr && sites != nullptr)),
                 "pq3d_attention_fwd: dropout needs p in [0,1), a device seed and one site id per memory");
  PQ3D_CHECK_ARG(n_mem >= 1 && n_mem <= kMaxMem, "pq3d_attention_fwd: n_mem=%d not in [1,%d]", n_mem, kMaxMem);
  PQ3D_CHECK_ARG(Q && O && K && Vt && S && S_pitch && Vt_pitch, "pq3d_attention_fwd: null argument");
  PQ3D_CHECK_ARG(B > 0 && H > 0 && Nq > 0, "pq3d_attention_fwd: bad shape B=%d H=%d Nq=%d", B, H, Nq);
  PQ3D_CHECK_ARG(ldq % 8 == 0 && ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(Q) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(O) & 15) == 0 && (o_mem_stride % 8) == 0,
                 "pq3d_attention_fwd: Q / O must be 16-byte aligned with leading dimensions multiple of 8");
  AttnMaps maps;          // filled per call; passed by value to the launch
  AttnParams p;
  memset(&p, 0, sizeof(p));
  {
    uint64_t dims[3] = {(uint64_t)ldq, (uint64_t)Nq, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)ldq * 2, (uint64_t)Nq * ldq * 2};
    uint32_t box[3] = {(uint32_t)kHeadDim, 128u, 1u};
    int rc = make_tmap_bf16(&maps.q, Q, 3, dims, strides, box);
    if (rc != PQ3D_OK) return rc;
  }
  for (int i = 0; i < n_mem; ++i) {
    PQ3D_CHECK_ARG(S[i] > 0 && S_pitch[i] >= S[i] && Vt_pitch[i] >= S[i] && Vt_pitch[i] % 8 == 0,
                   "pq3d_attention_fwd: memory %d: S=%d S_pitch=%d Vt_pitch=%d (pitches must be >= S, Vt_pitch a "
                   "multiple of 8)", i, S[i], S_pitch[i], Vt_pitch[i]);
    PQ3D_CHECK_ARG(ldk[i] % 8 == 0 && ldvt[i] % 8 == 0 && ldvt[i] >= (int64_t)B * Vt_pitch[i] &&
                       (reinterpret_cast<uintptr_t>(K[i]) & 15) == 0 && (reinterpret_cast<uintptr_t>(Vt[i]) & 15) == 0,
                   "pq3d_attention_fwd: memory %d: K / V^T alignment or leading dimension", i);
    PQ3D_CHECK_ARG(score_bias == nullptr ||
                       (bias_ld % 4 == 0 && bias_ld >= ((S[i] + kKvTile - 1) / kKvTile) * kKvTile &&
                        (reinterpret_cast<uintptr_t>(score_bias) & 15) == 0),
                   "pq3d_attention_fwd: score_bias rows must be 16-byte aligned and padded to a multiple of %d keys",
                   kKvTile);
    {
      uint64_t dims[3] = {(uint64_t)ldk[i], (uint64_t)S[i], (uint64_t)B};
      uint64_t strides[2] = {(uint64_t)ldk[i] * 2, (uint64_t)S_pitch[i] * ldk[i] * 2};
      uint32_t box[3] = {(uint32_t)kHeadDim, (uint32_t)kKvTile, 1u};
      int rc = make_tmap_bf16(&maps.k[i], K[i], 3, dims, strides, box);
      if (rc != PQ3D_OK) return rc;
    }
    {
      uint64_t dims[3] = {(uint64_t)S[i], (uint64_t)B, (uint64_t)vt_rows[i]};
      uint64_t strides[2] = {(uint64_t)Vt_pitch[i] * 2, (uint64_t)ldvt[i] * 2};
      uint32_t box[3] = {64u, 1u, (uint32_t)kHeadDim};
      int rc = make_tmap_bf16(&maps.vt[i], Vt[i], 3, dims, strides, box);
      if (rc != PQ3D_OK) return rc;
    }
    AttnMem& m = p.mem[i];
    m.mask_bits = mask_bits ? mask_bits[i] : nullptr;
    m.mask_b_stride = mask_bits ? mask_b_stride[i] : 0;
    m.mask_h_stride = mask_bits ? mask_h_stride[i] : 0;
    m.mask_q_stride = mask_bits ? mask_q_stride[i] : 0;
    m.kv_tiles = kv_tiles ? kv_tiles[i] : nullptr;
    m.S = S[i];
    m.k_col0 = (int32_t)k_col0[i];
    m.vt_row0 = (int32_t)vt_row0[i];
  }
  for (int i = n_mem; i < kMaxMem; ++i) {
    maps.k[i] = maps.k[0];
    maps.vt[i] = maps.vt[0];
  }
  p.O = reinterpret_cast<__nv_bfloat16*>(O);
  p.ldo = ldo;
  p.o_mem_stride = o_mem_stride;
  p.q_mem_stride = (int32_t)q_mem_stride;
  p.B = B;
  p.H = H;
  p.Nq = Nq;
  p.q_tiles = (Nq + 127) / 128;
  p.zero_attn = zero_attn;
  p.force_two_pass = g_force_two_pass;
  p.score_bias = score_bias;
  p.bias_ld = bias_ld;
  p.stat_m = stat_m;
  p.stat_l = stat_l;
  p.stat_mem_stride = stat_mem_stride;
  PQ3D_CHECK_ARG((stat_m == nullptr) == (stat_l == nullptr), "pq3d_attention_fwd: stat_m and stat_l go together");
  p.dbg = g_attn_timeline;
  if (drop_p > 0.f) {
    p.drop_thresh = drop_threshold(drop_p);
    p.drop_scale = 1.f / (1.f - drop_p);
    p.seed = seed_dev;
    for (int i = 0; i < n_mem; ++i) {
      p.site[i] = sites[i];
      PQ3D_CHECK_ARG(static_cast<int64_t>(B) * H * Nq * (((S[i] + kKvTile - 1) / kKvTile) * kKvTile) < (int64_t(1) << 32),
                     "pq3d_attention_fwd: memory %d has too many score elements for the 32-bit dropout counter", i);
    }
  }
  static bool configured = false;
  if (!configured) {
    PQ3D_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    configured = true;
  }
  dim3 grid(H, B, n_mem * p.q_tiles);
  PQ3D_CUDA(launch_kernel(attention_fwd_kernel, grid, dim3(kAttnThreads), kAttnSmem,
                          reinterpret_cast<cudaStream_t>(stream), maps, p));
  return PQ3D_OK;
}

extern "C" int pq3d_attention_fwd(int n_mem, const void* Q, int64_t ldq, int64_t q_mem_stride, const void* const* K,
                                  const int64_t* ldk, const int64_t* k_col0, const void* const* Vt,
                                  const int64_t* ldvt, const int64_t* vt_row0, const int64_t* vt_rows,
                                  const int32_t* S, const int32_t* S_pitch, const int32_t* Vt_pitch,
                                  const uint32_t* const* mask_bits, const int64_t* mask_b_stride,
                                  const int64_t* mask_h_stride, const int64_t* mask_q_stride,
                                  const int32_t* const* kv_tiles, void* O, int64_t ldo,
                                  int64_t o_mem_stride, int B, int H, int Nq, int zero_attn, const float* score_bias,
                                  int64_t bias_ld, float* stat_m, float* stat_l, int64_t stat_mem_stride,
                                  void* stream) {
  return attention_fwd_impl(n_mem, Q, ldq, q_mem_stride, K, ldk, k_col0, Vt, ldvt, vt_row0, vt_rows, S, S_pitch, Vt_pitch,
                            mask_bits, mask_b_stride, mask_h_stride, mask_q_stride, kv_tiles, O, ldo, o_mem_stride, B, H,
                            Nq, zero_attn, score_bias, bias_ld, stat_m, stat_l, stat_mem_stride, 0.f, nullptr, nullptr,
                            stream);
}

// Training-mode forward: same as pq3d_attention_fwd plus dropout on the attention probabilities.
extern "C" int pq3d_attention_fwd_train(int n_mem, const void* Q, int64_t ldq, int64_t q_mem_stride, const void* const* K,
                                        const int64_t* ldk, const int64_t* k_col0, const void* const* Vt,
                                        const int64_t* ldvt, const int64_t* vt_row0, const int64_t* vt_rows,
                                        const int32_t* S, const int32_t* S_pitch, const int32_t* Vt_pitch,
                                        const uint32_t* const* mask_bits, const int64_t* mask_b_stride,
                                        const int64_t* mask_h_stride, const int64_t* mask_q_stride,
                                        const int32_t* const* kv_tiles, void* O, int64_t ldo, int64_t o_mem_stride,
                                        int B, int H, int Nq, int zero_attn, const float* score_bias, int64_t bias_ld,
                                        float* stat_m, float* stat_l, int64_t stat_mem_stride, float drop_p,
                                        const uint32_t* seed_dev, const uint32_t* sites, void* stream) {
  return attention_fwd_impl(n_mem, Q, ldq, q_mem_stride, K, ldk, k_col0, Vt, ldvt, vt_row0, vt_rows, S, S_pitch, Vt_pitch,
                            mask_bits, mask_b_stride, mask_h_stride, mask_q_stride, kv_tiles, O, ldo, o_mem_stride, B, H,
                            Nq, zero_attn, score_bias, bias_ld, stat_m, stat_l, stat_mem_stride, drop_p, seed_dev, sites,
                            stream);
}
