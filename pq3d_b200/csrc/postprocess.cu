// Inference post-processing of the decoder's instance predictions, per scene:
// evaluator/instseg_eval.py:85-150 (eval_instance_step) and :272-305 (get_full_res_mask, get_mask_and_scores).
//
//   probs  = softmax(pred_logits)[:, :-1]                                  (Q, C)        drop the no-object class
//   top-K over the flattened (Q x C) scores, sorted                        -> score_k, query q_k, class c_k
//   masks  = pred_masks[voxel2segment][:, q_k]                              (V, K)        segment -> voxel gather
//   mask_score_k = sum_v sig(m) [m > 0] / (sum_v [m > 0] + 1e-6);  score_k *= mask_score_k
//   full resolution: (m > 0)[voxel_to_full] -> scatter_mean over segment_to_full -> > 0.5 -> [segment_to_full]   (P, K)
//   heatmap = sig(masks)[voxel_to_full]                                     (P, K)
//   columns sorted by score, descending.
//
// The reference materialises the (V, Q) voxel-level logits, a (V, K) gather, (P, K) float masks, a float scatter_mean and
// a second (P, K) gather on the HOST.  Here nothing voxel-sized is ever built:
//   * a voxel's logit is its segment's logit, so the voxel sums are segment sums weighted by the voxel count of the
//     segment (an integer histogram of voxel2segment);
//   * the full-resolution vote "mean over the points of a full-res segment > 0.5" is an INTEGER majority vote:
//     2 * #points with m > 0 > #points — bit-exact (a mean of 0/1 floats over < 2^24 points is exact in fp32);
//   * the final order by score only needs segment-level quantities, so it is known BEFORE the (P, K) pass, which then
//     writes its columns already sorted — one HBM-bound pass: P*(8 + 8 + 8) index bytes in, P*K*(1 + 4) bytes out.
#include "host_common.h"
#include "ptx.cuh"

namespace pq3d {

// ---- probs = softmax(logits)[:, :C] for logits [Q, C + 1]: one warp per query
__global__ void class_probs_kernel(const float* __restrict__ logits, float* __restrict__ probs, int Q, int C1) {
  pdl_sync();
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
  if (q >= Q) return;
  const float* row = logits + static_cast<int64_t>(q) * C1;
  float mx = -INFINITY;
  for (int c = lane; c < C1; c += 32) mx = fmaxf(mx, row[c]);
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  float den = 0.f;
  for (int c = lane; c < C1; c += 32) den += expf(row[c] - mx);
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) den += __shfl_xor_sync(0xffffffffu, den, d);
  for (int c = lane; c < C1 - 1; c += 32) probs[static_cast<int64_t>(q) * (C1 - 1) + c] = expf(row[c] - mx) / den;
}

// ---- exact top-K of n non-negative floats, sorted descending (ties: lower index first).  One block of 1024 threads:
// 4-pass radix select on the float bits finds the K-th largest key, the survivors are compacted in index order and
// bitonic-sorted in shared memory.  n <= 2^20, K <= 1024.
constexpr int kTopThreads = 1024;
__device__ __forceinline__ uint32_t key_of(float v) {                 // total order on floats, larger float = larger key
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void __launch_bounds__(kTopThreads) topk_kernel(const float* __restrict__ vals, int n, int K,
                                                           float* __restrict__ out_val, int32_t* __restrict__ out_idx) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_prefix, s_need;
  __shared__ uint64_t sel[1024];                 // (key << 32) | (0xffffffff - idx): one 64-bit descending sort key
  __shared__ int s_count, s_eq_taken;
  pdl_sync();
  const int tid = threadIdx.x;
  uint32_t prefix = 0, mask = 0;
  int need = K;                                  // how many of the elements matching `prefix` on `mask` we still need
  for (int pass = 3; pass >= 0; --pass) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kTopThreads) {
      const uint32_t k = key_of(vals[i]);
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> (8 * pass)) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t acc = 0;
      int b = 255;
      for (; b > 0; --b) {                       // walk the buckets from the largest digit down
        if (acc + hist[b] >= static_cast<uint32_t>(need)) break;
        acc += hist[b];
      }
      s_prefix = prefix | (static_cast<uint32_t>(b) << (8 * pass));
      s_need = need - acc;
    }
    __syncthreads();
    prefix = s_prefix;
    need = static_cast<int>(s_need);
    mask |= 255u << (8 * pass);
    __syncthreads();
  }
  // prefix = key of the K-th largest element; `need` = how many elements EQUAL to it belong to the top K
  if (tid == 0) { s_count = 0; s_eq_taken = 0; }
  __syncthreads();
  const uint32_t thr = prefix;
  // strictly larger: any order (sorted afterwards); equal: lowest indices first -> chunked, ordered pass
  for (int i = tid; i < n; i += kTopThreads) {
    const uint32_t k = key_of(vals[i]);
    if (k > thr) {
      const int pos = atomicAdd(&s_count, 1);
      sel[pos] = (static_cast<uint64_t>(k) << 32) | (0xffffffffu - static_cast<uint32_t>(i));
    }
  }
  __syncthreads();
  const int n_gt = s_count;
  for (int base = 0; base < n && s_eq_taken < need; base += kTopThreads) {   // uniform loop: s_eq_taken read after sync
    const int i = base + tid;
    const bool eq = i < n && key_of(vals[i]) == thr;
    // rank of this thread among the equal elements of the chunk (ballot + warp prefix through shared memory)
    const unsigned bal = __ballot_sync(0xffffffffu, eq);
    __shared__ int warp_cnt[32];
    if (lane_id() == 0) warp_cnt[tid >> 5] = __popc(bal);
    __syncthreads();
    int before = s_eq_taken;
    for (int w = 0; w < (tid >> 5); ++w) before += warp_cnt[w];
    const int rank = before + __popc(bal & ((1u << lane_id()) - 1u));
    if (eq && rank < need) sel[n_gt + rank] = (static_cast<uint64_t>(thr) << 32) | (0xffffffffu - static_cast<uint32_t>(i));
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += warp_cnt[w];
      s_eq_taken += tot;
    }
    __syncthreads();
  }
  // pad to a power of two and bitonic-sort descending
  int P2 = 1;
  while (P2 < K) P2 <<= 1;
  for (int i = K + tid; i < P2; i += kTopThreads) sel[i] = 0ull;
  __syncthreads();
  for (int size = 2; size <= P2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < P2; i += kTopThreads) {
        const int j = i ^ stride;
        if (j > i) {
          const bool desc = (i & size) == 0;
          const uint64_t a = sel[i], b = sel[j];
          if ((a < b) == desc) { sel[i] = b; sel[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < K; i += kTopThreads) {
    const uint32_t idx = 0xffffffffu - static_cast<uint32_t>(sel[i] & 0xffffffffull);
    out_idx[i] = static_cast<int32_t>(idx);
    out_val[i] = vals[idx];
  }
}

// ---- integer histogram (voxels per segment; points per full-resolution segment)
__global__ void bincount_kernel(const int64_t* __restrict__ idx, int64_t n, int32_t* __restrict__ out, int bins) {
  pdl_sync();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = idx[i];
    if (b >= 0 && b < bins) atomicAdd(&out[b], 1);
  }
}

// ---- mask score of the K selected (query, class) pairs from SEGMENT-level logits: one warp per k
//   num = sum_s cnt[s] * sig(x) * [x > 0], den = sum_s cnt[s] * [x > 0], x = pred_masks[s, q_k]
__global__ void mask_score_kernel(const float* __restrict__ pred_masks, const int32_t* __restrict__ seg_count,
                                  const int32_t* __restrict__ sel_flat, int C, int S, int Q, int K,
                                  const float* __restrict__ cls_score, float* __restrict__ score) {
  pdl_sync();
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
  if (k >= K) return;
  const int q = sel_flat[k] / C;
  float num = 0.f, den = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float x = pred_masks[static_cast<int64_t>(s) * Q + q];
    if (x > 0.f) {
      const float c = static_cast<float>(seg_count[s]);
      num += c * (1.f / (1.f + expf(-x)));
      den += c;
    }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    num += __shfl_xor_sync(0xffffffffu, num, d);
    den += __shfl_xor_sync(0xffffffffu, den, d);
  }
  if (lane == 0) score[k] = cls_score[k] * (num / (den + 1e-6f));
}

// ---- full-resolution pass 1: votes[s', k] += [pred_masks[seg(p), q_k] > 0] for every point p of full-res segment s'
__global__ void fullres_vote_kernel(const float* __restrict__ pred_masks, const int32_t* __restrict__ q_of,
                                    const int64_t* __restrict__ voxel2segment, const int64_t* __restrict__ voxel_to_full,
                                    const int64_t* __restrict__ segment_to_full, int64_t P, int Q, int K,
                                    int32_t* __restrict__ votes) {
  pdl_sync();
  const int64_t total = P * K;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t p = e / K;
    const int k = static_cast<int>(e % K);
    const int64_t s = voxel2segment[voxel_to_full[p]];
    if (pred_masks[s * Q + q_of[k]] > 0.f) atomicAdd(&votes[segment_to_full[p] * K + k], 1);
  }
}
// ---- pass 2: mask[p, k] = 2 * votes[s'(p), k] > points[s'(p)];  heat[p, k] = sig(pred_masks[seg(p), q_k])
__global__ void fullres_emit_kernel(const float* __restrict__ pred_masks, const int32_t* __restrict__ q_of,
                                    const int64_t* __restrict__ voxel2segment, const int64_t* __restrict__ voxel_to_full,
                                    const int64_t* __restrict__ segment_to_full, const int32_t* __restrict__ votes,
                                    const int32_t* __restrict__ points, int64_t P, int Q, int K,
                                    float* __restrict__ mask, float* __restrict__ heat) {
  pdl_sync();
  const int64_t total = P * K;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t p = e / K;
    const int k = static_cast<int>(e % K);
    const int64_t sf = segment_to_full[p];
    mask[e] = (2 * votes[sf * K + k] > points[sf]) ? 1.f : 0.f;
    if (heat != nullptr) {
      const float x = pred_masks[voxel2segment[voxel_to_full[p]] * Q + q_of[k]];
      heat[e] = 1.f / (1.f + expf(-x));
    }
  }
}

__global__ void split_index_kernel(const int32_t* __restrict__ flat, const int32_t* __restrict__ order, int K, int C,
                                   int32_t* __restrict__ q_of, int32_t* __restrict__ cls_of) {
  pdl_sync();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int f = flat[order == nullptr ? k : order[k]];
  q_of[k] = f / C;
  cls_of[k] = f % C;
}

}  // namespace pq3d

using namespace pq3d;

static int grid_of(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" int pq3d_class_probs(const float* logits, float* probs, int Q, int C1, void* stream) {
  PQ3D_CHECK_ARG(logits && probs && Q > 0 && C1 > 1, "pq3d_class_probs: bad argument");
  PQ3D_CUDA(launch_kernel(class_probs_kernel, dim3((Q * 32 + 255) / 256), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                          logits, probs, Q, C1));
  return PQ3D_OK;
}

// Exact top-K (sorted descending; ties: lower index first) of n floats; K <= 1024, n <= 2^20.
extern "C" int pq3d_topk(const float* vals, int n, int K, float* out_val, int32_t* out_idx, void* stream) {
  PQ3D_CHECK_ARG(vals && out_val && out_idx, "pq3d_topk: null argument");
  PQ3D_CHECK_ARG(n > 0 && K > 0 && K <= 1024 && K <= n && n <= (1 << 20), "pq3d_topk: n=%d K=%d (K <= min(n, 1024), n <= 2^20)", n, K);
  PQ3D_CUDA(launch_kernel(topk_kernel, dim3(1), dim3(kTopThreads), 0, reinterpret_cast<cudaStream_t>(stream), vals, n, K,
                          out_val, out_idx));
  return PQ3D_OK;
}

extern "C" int pq3d_bincount(const int64_t* idx, int64_t n, int32_t* out, int bins, void* stream) {
  PQ3D_CHECK_ARG(idx && out && n >= 0 && bins > 0, "pq3d_bincount: bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  PQ3D_CUDA(cudaMemsetAsync(out, 0, sizeof(int32_t) * bins, st));
  if (n > 0) PQ3D_CUDA(launch_kernel(bincount_kernel, dim3(grid_of(n, 256)), dim3(256), 0, st, idx, n, out, bins));
  return PQ3D_OK;
}

// score[k] = cls_score[k] * mask_score(q_k), q_k = sel_flat[k] / C (get_mask_and_scores, instseg_eval.py:285-303).
extern "C" int pq3d_instseg_scores(const float* pred_masks, const int32_t* seg_count, const int32_t* sel_flat, int C, int S,
                                   int Q, int K, const float* cls_score, float* score, void* stream) {
  PQ3D_CHECK_ARG(pred_masks && seg_count && sel_flat && cls_score && score && C > 0 && S > 0 && Q > 0 && K > 0,
                 "pq3d_instseg_scores: bad argument");
  PQ3D_CUDA(launch_kernel(mask_score_kernel, dim3((K * 32 + 255) / 256), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                          pred_masks, seg_count, sel_flat, C, S, Q, K, cls_score, score));
  return PQ3D_OK;
}

// q_of[k] = flat[order[k]] / C, cls_of[k] = flat[order[k]] % C (order may be NULL).
extern "C" int pq3d_split_index(const int32_t* flat, const int32_t* order, int K, int C, int32_t* q_of, int32_t* cls_of,
                                void* stream) {
  PQ3D_CHECK_ARG(flat && q_of && cls_of && K > 0 && C > 0, "pq3d_split_index: bad argument");
  PQ3D_CUDA(launch_kernel(split_index_kernel, dim3((K + 255) / 256), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), flat,
                          order, K, C, q_of, cls_of));
  return PQ3D_OK;
}

// Full-resolution masks (P, K) float 0/1 and heatmaps (P, K) for the K selected queries q_of (already in output order).
// votes: int32 [n_fullseg, K] scratch, points: int32 [n_fullseg] = bincount(segment_to_full) — both caller-owned.
extern "C" int pq3d_instseg_fullres(const float* pred_masks, const int32_t* q_of, const int64_t* voxel2segment,
                                    const int64_t* voxel_to_full, const int64_t* segment_to_full, int64_t P, int S, int Q,
                                    int K, int n_fullseg, int32_t* votes, const int32_t* points, float* mask, float* heat,
                                    void* stream) {
  PQ3D_CHECK_ARG(pred_masks && q_of && voxel2segment && voxel_to_full && segment_to_full && votes && points && mask,
                 "pq3d_instseg_fullres: null argument");
  PQ3D_CHECK_ARG(P > 0 && S > 0 && Q > 0 && K > 0 && n_fullseg > 0, "pq3d_instseg_fullres: bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  PQ3D_CUDA(cudaMemsetAsync(votes, 0, sizeof(int32_t) * static_cast<size_t>(n_fullseg) * K, st));
  const int grid = grid_of(P * K, 256);
  PQ3D_CUDA(launch_kernel(fullres_vote_kernel, dim3(grid), dim3(256), 0, st, pred_masks, q_of, voxel2segment, voxel_to_full,
                          segment_to_full, P, Q, K, votes));
  PQ3D_CUDA(launch_kernel(fullres_emit_kernel, dim3(grid), dim3(256), 0, st, pred_masks, q_of, voxel2segment, voxel_to_full,
                          segment_to_full, static_cast<const int32_t*>(votes), points, P, Q, K, mask, heat));
  return PQ3D_OK;
}
