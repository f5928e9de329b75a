// Matcher cost matrices and matched mask losses, the step directly downstream of the decoder's
// `predictions_mask` / `predictions_class` in stage-1 training (13 prediction sets per step).
//
// Reference: modules/third_party/mask3d/matcher.py:12-64 (batch_dice_loss, batch_sigmoid_ce_loss: three
// "nc,mc->nm" einsums per scene over all S points, computed in fp32 under autocast(enabled=False), :160-181) and
// :103-181 (class cost -softmax(logits)[:, tgt_ids], -1 for ignore labels; C = w_mask*mask + w_class*class +
// w_dice*dice); modules/third_party/mask3d/criterion.py:26-75,163-196 (dice_loss / sigmoid_ce_loss on the matched
// pairs).  The reference materialises pos / neg / sigmoid (N x S each), 1 - targets (M x S) and three cost matrices per
// scene and prediction set; here ONE kernel reads the mask logits once, forms softplus(+-x) and sigmoid(x) on the fly
// and contracts them against the (binary, uint8) target masks.
//
// Arithmetic is fp32 on the CUDA cores by design: the reference pins this block to fp32, the work is tiny
// (6*N*M*S = 37 MFLOP per scene at N=100, M=30, S=2048) and the kernel is bound by the transcendental pipe and the
// N x S logit read (0.8 MB per scene), not by FLOPs — reshaping it for tensor cores would buy nothing.
#include "host_common.h"
#include "ptx.cuh"

namespace pq3d {

constexpr int kTQ = 16;      // queries per block
constexpr int kTM = 16;      // targets per block
constexpr int kTS = 64;      // points per shared-memory chunk

struct MatchParams {
  const float* pred_masks;    // [B, S, N]
  const float* pred_logits;   // [B, N, C]
  const uint8_t* tgt_masks;   // [B, Mmax, S], 1 = the point belongs to the instance
  const int64_t* tgt_labels;  // [B, Mmax]
  const int32_t* tgt_count;   // [B]
  float* cost;                // [B, N, Mmax]
  int32_t B, N, S, C, Mmax;
  float w_class, w_mask, w_dice;
  int64_t ignore_label;
};

__device__ __forceinline__ float softplus_neg_abs(float x) { return log1pf(expf(-fabsf(x))); }

__global__ void __launch_bounds__(kTQ* kTM) match_cost_kernel(const MatchParams p) {
  __shared__ float s_pos[kTS][kTQ], s_neg[kTS][kTQ], s_sig[kTS][kTQ];
  __shared__ float s_t[kTM][kTS + 1];
  __shared__ float s_max[kTQ], s_den[kTQ];
  pdl_sync();
  const int b = blockIdx.z, n0 = blockIdx.x * kTQ, m0 = blockIdx.y * kTM;
  const int M = p.tgt_count[b];
  if (m0 >= M && m0 > 0) {                        // nothing to match in this column tile: zero it and leave
    const int tn = threadIdx.x % kTQ, tm = threadIdx.x / kTQ;
    if (n0 + tn < p.N && m0 + tm < p.Mmax) p.cost[(static_cast<int64_t>(b) * p.N + n0 + tn) * p.Mmax + m0 + tm] = 0.f;
    return;
  }
  const int tn = threadIdx.x % kTQ, tm = threadIdx.x / kTQ;
  // ---- softmax statistics of this block's query rows (torch: exp(x - max) / sum)
  {
    const int row = threadIdx.x / 16, l16 = threadIdx.x % 16;     // 16 lanes per query row
    const int n = n0 + row;
    float mx = -INFINITY;
    const float* lg = p.pred_logits + (static_cast<int64_t>(b) * p.N + (n < p.N ? n : p.N - 1)) * p.C;
    for (int c = l16; c < p.C; c += 16) mx = fmaxf(mx, lg[c]);
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d, 16));
    float den = 0.f;
    for (int c = l16; c < p.C; c += 16) den += expf(lg[c] - mx);
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) den += __shfl_xor_sync(0xffffffffu, den, d, 16);
    if (l16 == 0) {
      s_max[row] = mx;
      s_den[row] = den;
    }
  }
  float a_pos = 0.f, a_neg = 0.f, a_sig = 0.f, sum_sig = 0.f, sum_t = 0.f;
  const float* pm = p.pred_masks + static_cast<int64_t>(b) * p.S * p.N;
  const uint8_t* tg = p.tgt_masks + static_cast<int64_t>(b) * p.Mmax * p.S;
  for (int c0 = 0; c0 < p.S; c0 += kTS) {
    __syncthreads();
    for (int e = threadIdx.x; e < kTS * kTQ; e += kTQ * kTM) {      // (point, query): logits contiguous over queries
      const int c = e / kTQ, q = e % kTQ;
      float ps = 0.f, ng = 0.f, sg = 0.f;
      if (c0 + c < p.S && n0 + q < p.N) {
        const float x = pm[static_cast<int64_t>(c0 + c) * p.N + n0 + q];
        const float sp = softplus_neg_abs(x);
        ps = fmaxf(-x, 0.f) + sp;                 // BCE-with-logits against 1
        ng = fmaxf(x, 0.f) + sp;                  // ... against 0
        sg = 1.f / (1.f + expf(-x));
      }
      s_pos[c][q] = ps; s_neg[c][q] = ng; s_sig[c][q] = sg;
    }
    for (int e = threadIdx.x; e < kTM * kTS; e += kTQ * kTM) {      // (target, point): masks contiguous over points
      const int m = e / kTS, c = e % kTS;
      s_t[m][c] = (m0 + m < M && c0 + c < p.S) ? static_cast<float>(tg[static_cast<int64_t>(m0 + m) * p.S + c0 + c] != 0)
                                               : 0.f;
    }
    __syncthreads();
    const int lim = min(kTS, p.S - c0);
    for (int c = 0; c < lim; ++c) {
      const float t = s_t[tm][c];
      a_pos = fmaf(s_pos[c][tn], t, a_pos);
      a_neg = fmaf(s_neg[c][tn], 1.f - t, a_neg);
      a_sig = fmaf(s_sig[c][tn], t, a_sig);
      sum_sig += s_sig[c][tn];
      sum_t += t;
    }
  }
  const int n = n0 + tn, m = m0 + tm;
  if (n >= p.N || m >= p.Mmax) return;
  float out = 0.f;
  if (m < M) {
    const float cost_mask = (a_pos + a_neg) / static_cast<float>(p.S);
    const float cost_dice = 1.f - (2.f * a_sig + 1.f) / (sum_sig + sum_t + 1.f);
    const int64_t lab = p.tgt_labels[static_cast<int64_t>(b) * p.Mmax + m];
    float cost_class = -1.f;                                        // ignore labels pretend a perfect match (:128-132)
    if (lab != p.ignore_label) {
      const int64_t l = lab < 0 ? 0 : (lab >= p.C ? p.C - 1 : lab);
      cost_class = -(expf(p.pred_logits[(static_cast<int64_t>(b) * p.N + n) * p.C + l] - s_max[tn]) / s_den[tn]);
    }
    out = p.w_mask * cost_mask + p.w_class * cost_class + p.w_dice * cost_dice;
  }
  p.cost[(static_cast<int64_t>(b) * p.N + n) * p.Mmax + m] = out;
}

// Matched mask losses, forward: for pair k = (scene b_k, query q_k, target t_k):
//   ce[k]   = mean_c BCE(x[c], t[c])                                  (criterion.py:51-71, loss.mean(1))
//   dice[k] = 1 - (2 * sum sig*t + 1) / (sum sig + sum t + 1)          (criterion.py:26-46)
// and the per-pair sums the backward needs.  One warp per pair.
__global__ void __launch_bounds__(256) matched_loss_fwd_kernel(const float* __restrict__ pred_masks,
                                                               const uint8_t* __restrict__ tgt_masks,
                                                               const int32_t* __restrict__ pairs, int n_pairs, int N,
                                                               int S, int Mmax, float* __restrict__ ce,
                                                               float* __restrict__ dice, float* __restrict__ sums) {
  pdl_sync();
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (k >= n_pairs) return;
  const int lane = lane_id();
  const int b = pairs[3 * k], q = pairs[3 * k + 1], t = pairs[3 * k + 2];
  const float* pm = pred_masks + static_cast<int64_t>(b) * S * N + q;
  const uint8_t* tg = tgt_masks + (static_cast<int64_t>(b) * Mmax + t) * S;
  float s_ce = 0.f, s_st = 0.f, s_s = 0.f, s_t = 0.f;
  for (int c = lane; c < S; c += 32) {
    const float x = pm[static_cast<int64_t>(c) * N];
    const float tt = static_cast<float>(tg[c] != 0);
    const float sg = 1.f / (1.f + expf(-x));
    s_ce += fmaxf(x, 0.f) - x * tt + softplus_neg_abs(x);
    s_st += sg * tt;
    s_s += sg;
    s_t += tt;
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    s_ce += __shfl_xor_sync(0xffffffffu, s_ce, d);
    s_st += __shfl_xor_sync(0xffffffffu, s_st, d);
    s_s += __shfl_xor_sync(0xffffffffu, s_s, d);
    s_t += __shfl_xor_sync(0xffffffffu, s_t, d);
  }
  if (lane == 0) {
    ce[k] = s_ce / static_cast<float>(S);
    dice[k] = 1.f - (2.f * s_st + 1.f) / (s_s + s_t + 1.f);
    sums[2 * k] = s_st;
    sums[2 * k + 1] = s_s + s_t;
  }
}

// backward: d_pred[b, c, q] += g_ce[k]/S * (sig - t) + g_dice[k] * d dice / d x,
//   d dice / d x = -(2 t (D + 1) - (2 I + 1)) / (D + 1)^2 * sig (1 - sig),  I = sums[2k], D = sums[2k+1].
// Pairs of one prediction set never share a (scene, query), so plain stores into a zero-filled gradient suffice.
__global__ void __launch_bounds__(256) matched_loss_bwd_kernel(const float* __restrict__ pred_masks,
                                                               const uint8_t* __restrict__ tgt_masks,
                                                               const int32_t* __restrict__ pairs, int n_pairs, int N,
                                                               int S, int Mmax, const float* __restrict__ g_ce,
                                                               const float* __restrict__ g_dice,
                                                               const float* __restrict__ sums,
                                                               float* __restrict__ d_pred) {
  pdl_sync();
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (k >= n_pairs) return;
  const int lane = lane_id();
  const int b = pairs[3 * k], q = pairs[3 * k + 1], t = pairs[3 * k + 2];
  const float* pm = pred_masks + static_cast<int64_t>(b) * S * N + q;
  float* dp = d_pred + static_cast<int64_t>(b) * S * N + q;
  const uint8_t* tg = tgt_masks + (static_cast<int64_t>(b) * Mmax + t) * S;
  const float gc = g_ce[k] / static_cast<float>(S), gd = g_dice[k];
  const float I2 = 2.f * sums[2 * k] + 1.f, D1 = sums[2 * k + 1] + 1.f;
  const float inv = 1.f / (D1 * D1);
  for (int c = lane; c < S; c += 32) {
    const float x = pm[static_cast<int64_t>(c) * N];
    const float tt = static_cast<float>(tg[c] != 0);
    const float sg = 1.f / (1.f + expf(-x));
    dp[static_cast<int64_t>(c) * N] = gc * (sg - tt) - gd * (2.f * tt * D1 - I2) * inv * sg * (1.f - sg);
  }
}

}  // namespace pq3d

using namespace pq3d;

// cost[b, n, m] for m < tgt_count[b] (0 beyond).  pred_masks fp32 [B, S, N] contiguous (the decoder's
// predictions_mask entry), pred_logits fp32 [B, N, C] contiguous, tgt_masks uint8 [B, Mmax, S], tgt_labels int64
// [B, Mmax], tgt_count int32 [B].  Replaces HungarianMatcher.memory_efficient_forward's per-scene cost assembly
// (matcher.py:110-181) with num_points = -1 (all points, the shipped configuration).
extern "C" int pq3d_match_cost(const float* pred_masks, const float* pred_logits, const uint8_t* tgt_masks,
                               const int64_t* tgt_labels, const int32_t* tgt_count, float* cost, int B, int N, int S, int C,
                               int Mmax, float w_class, float w_mask, float w_dice, int64_t ignore_label, void* stream) {
  PQ3D_CHECK_ARG(pred_masks && pred_logits && tgt_masks && tgt_labels && tgt_count && cost, "pq3d_match_cost: null argument");
  PQ3D_CHECK_ARG(B > 0 && N > 0 && S > 0 && C > 0 && Mmax > 0 && B <= 65535, "pq3d_match_cost: bad shape");
  MatchParams p;
  p.pred_masks = pred_masks; p.pred_logits = pred_logits; p.tgt_masks = tgt_masks; p.tgt_labels = tgt_labels;
  p.tgt_count = tgt_count; p.cost = cost;
  p.B = B; p.N = N; p.S = S; p.C = C; p.Mmax = Mmax;
  p.w_class = w_class; p.w_mask = w_mask; p.w_dice = w_dice; p.ignore_label = ignore_label;
  dim3 grid((N + kTQ - 1) / kTQ, (Mmax + kTM - 1) / kTM, B);
  PQ3D_CUDA(launch_kernel(match_cost_kernel, grid, dim3(kTQ * kTM), 0, reinterpret_cast<cudaStream_t>(stream), p));
  return PQ3D_OK;
}

// pairs: int32 [n_pairs, 3] = (scene, query, target) on the device.  Outputs fp32 [n_pairs]: ce, dice; sums fp32
// [n_pairs, 2] is kept for pq3d_matched_mask_loss_bwd.
extern "C" int pq3d_matched_mask_loss_fwd(const float* pred_masks, const uint8_t* tgt_masks, const int32_t* pairs,
                                          int n_pairs, int B, int N, int S, int Mmax, float* ce, float* dice, float* sums,
                                          void* stream) {
  PQ3D_CHECK_ARG(pred_masks && tgt_masks && pairs && ce && dice && sums, "pq3d_matched_mask_loss_fwd: null argument");
  PQ3D_CHECK_ARG(n_pairs > 0 && B > 0 && N > 0 && S > 0 && Mmax > 0, "pq3d_matched_mask_loss_fwd: bad shape");
  PQ3D_CUDA(launch_kernel(matched_loss_fwd_kernel, dim3((n_pairs + 7) / 8), dim3(256), 0,
                          reinterpret_cast<cudaStream_t>(stream), pred_masks, tgt_masks, pairs, n_pairs, N, S, Mmax, ce,
                          dice, sums));
  return PQ3D_OK;
}

// d_pred fp32 [B, S, N] must be zero-filled by the caller; entries of matched (scene, query) columns are written.
extern "C" int pq3d_matched_mask_loss_bwd(const float* pred_masks, const uint8_t* tgt_masks, const int32_t* pairs,
                                          int n_pairs, int B, int N, int S, int Mmax, const float* g_ce,
                                          const float* g_dice, const float* sums, float* d_pred, void* stream) {
  PQ3D_CHECK_ARG(pred_masks && tgt_masks && pairs && g_ce && g_dice && sums && d_pred,
                 "pq3d_matched_mask_loss_bwd: null argument");
  PQ3D_CHECK_ARG(n_pairs > 0 && B > 0 && N > 0 && S > 0 && Mmax > 0, "pq3d_matched_mask_loss_bwd: bad shape");
  PQ3D_CUDA(launch_kernel(matched_loss_bwd_kernel, dim3((n_pairs + 7) / 8), dim3(256), 0,
                          reinterpret_cast<cudaStream_t>(stream), pred_masks, tgt_masks, pairs, n_pairs, N, S, Mmax, g_ce,
                          g_dice, sums, d_pred));
  return PQ3D_OK;
}
