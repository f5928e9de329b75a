// pq3d_linear_bf16: C = epilogue(A · Wᵀ) on tcgen05 tensor cores, operands staged by TMA.
//
// This is every nn.Linear on the decoder hot path — the cross-attention K/V/Q in-projections and
// out-projection (torch/nn/functional.py:5867-5873, :6653), the spatial self-attention
// w_qs/w_ks/w_vs/fc (modules/layers/transformers.py:180-185), the FFN (query_encoder.py:384) and the
// mask-head projections (modules/heads/mask_head.py:46-57).  The K/V projection is 85 % of the
// decoder's FLOPs, so this is the dominant kernel.
//
// One CTA = one 128 x BN output tile.  Warp roles (192 threads):
//   warp 0      TMA producer: A tile [128 x 64] and W tile [BN x 64] per k-block, 128B-swizzled,
//               into a kStages-deep shared-memory ring guarded by full/empty mbarriers
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue (M=128, N=BN, K=16, bf16 -> fp32
//               accumulators in TMEM); tcgen05.commit releases ring slots / signals the epilogue
//   warps 2..5  epilogue: tcgen05.ld the accumulator (lane = row, 32 columns per load), apply
//               (+bias) * alpha, optional ReLU / row zeroing, convert, 16-byte global stores
// Groups (blockIdx.z) shift the A / W / C / bias bases: per-memory out-projections, per-layer
// multi-scale voxel K/V projections and per-scene mask-logit products run as one launch.
#include "host_common.h"
#include "ptx.cuh"

namespace pq3d {

struct LinearParams {
  void* C;
  const float* bias;
  const uint8_t* row_zero;  // optional: rows with a non-zero byte are written as 0 (mask-head validity)
  int64_t ldc;
  int64_t c_group_stride;   // elements
  int64_t bias_group_stride;
  int64_t row_zero_group_stride;
  int32_t a_group_rows;     // row offset of group g in the A tensor map = g * a_group_rows
  int32_t w_group_rows;
  int32_t M, N, K;
  int32_t out_fp32;
  int32_t bias_along_m;
  int32_t relu;
  int32_t alpha_ncols;      // columns n < alpha_ncols are scaled by alpha (n >= : unscaled)
  float alpha;
  int32_t vec_ok;           // ldc / base alignment allow 16-byte stores
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle atom row
constexpr int kGemmThreads = 192;

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
linear_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                   const LinearParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* accum_bar = empty_bar + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int m0 = blockIdx.y * kBlockM;
  const int n0 = blockIdx.x * BN;
  const int g = blockIdx.z;
  const int num_kb = p.K / kBlockK;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<(BN < 32 ? 32 : BN)>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      const int a_row = g * p.a_group_rows + m0;
      const int w_row = g * p.w_group_rows + n0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % Cfg::kStages;
        const uint32_t ph = (kb / Cfg::kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1, 100 + s);
        uint8_t* sa = smem + s * Cfg::kStageBytes;
        mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
        tma_load_2d(sa, &tmap_a, &full_bar[s], kb * kBlockK, a_row);
        tma_load_2d(sa + Cfg::kABytes, &tmap_w, &full_bar[s], kb * kBlockK, w_row);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % Cfg::kStages;
        const uint32_t ph = (kb / Cfg::kStages) & 1;
        mbar_wait(&full_bar[s], ph, 200 + s);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
        const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          // advancing 16 bf16 (32 B) along K inside the 128-byte swizzle atom = +32 B on the start address
          umma_ss(tmem_base, umma_desc_k_sw128(sa + k * 32), umma_desc_k_sw128(sb + k * 32), idesc,
                  (kb | k) != 0 ? 1u : 0u);
        }
        tc_commit(&empty_bar[s]);
      }
      tc_commit(accum_bar);
    }
  } else {
    // epilogue warps 2..5 -> TMEM lane quadrant (warp % 4)
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane_id();
    const int m = m0 + row_in_tile;
    mbar_wait(accum_bar, 0, 300);
    tc_fence_after();
    const bool row_ok = m < p.M;
    const bool zero_row = row_ok && p.row_zero != nullptr && p.row_zero[g * p.row_zero_group_stride + m] != 0;
    const float bias_m = (p.bias != nullptr && p.bias_along_m && row_ok) ? p.bias[g * p.bias_group_stride + m] : 0.f;
    const float* bias_n = (p.bias != nullptr && !p.bias_along_m) ? p.bias + g * p.bias_group_stride : nullptr;
    uint8_t* c_row = reinterpret_cast<uint8_t*>(p.C) +
                     (static_cast<int64_t>(g) * p.c_group_stride + static_cast<int64_t>(m) * p.ldc) *
                         (p.out_fp32 ? 4 : 2);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t acc[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c0, acc);
      tmem_ld_wait();
      const int n_base = n0 + c0;
      if (!row_ok || n_base >= p.N) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int n = n_base + j;
        float x = __uint_as_float(acc[j]);
        if (bias_n != nullptr && n < p.N) x += __ldg(bias_n + n);
        x += bias_m;
        if (n < p.alpha_ncols) x *= p.alpha;
        if (p.relu) x = fmaxf(x, 0.f);
        v[j] = zero_row ? 0.f : x;
      }
      const bool full = n_base + 32 <= p.N;
      if (p.out_fp32) {
        float* dst = reinterpret_cast<float*>(c_row) + n_base;
        if (full && p.vec_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          for (int j = 0; j < 32; ++j)
            if (n_base + j < p.N) dst[j] = v[j];
        }
      } else {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(c_row) + n_base;
        if (full && p.vec_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 u;
            u.x = pack_bf16x2(v[j], v[j + 1]);
            u.y = pack_bf16x2(v[j + 2], v[j + 3]);
            u.z = pack_bf16x2(v[j + 4], v[j + 5]);
            u.w = pack_bf16x2(v[j + 6], v[j + 7]);
            *reinterpret_cast<uint4*>(dst + j) = u;
          }
        } else {
          for (int j = 0; j < 32; ++j)
            if (n_base + j < p.N) dst[j] = __float2bfloat16_rn(v[j]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem_base);
  }
}

template <int BN>
static int launch_linear(const CUtensorMap& ta, const CUtensorMap& tw, const LinearParams& p, int groups,
                         cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    PQ3D_CUDA(cudaFuncSetAttribute(linear_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg::kSmemBytes));
    configured = true;
  }
  dim3 grid((p.N + BN - 1) / BN, (p.M + kBlockM - 1) / kBlockM, groups);
  linear_bf16_kernel<BN><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tw, p);
  PQ3D_CUDA(cudaGetLastError());
  return PQ3D_OK;
}

}  // namespace pq3d

using namespace pq3d;

extern "C" int pq3d_linear_bf16(const void* A, int64_t lda, int64_t a_rows_total, int64_t a_group_rows,
                                const void* W, int64_t ldw, int64_t w_rows_total, int64_t w_group_rows, void* C,
                                int64_t ldc, int64_t c_group_stride, int out_fp32, const float* bias,
                                int64_t bias_group_stride, int bias_along_m, const uint8_t* row_zero,
                                int64_t row_zero_group_stride, int M, int N, int K, int groups, float alpha,
                                int alpha_ncols, int relu, int block_n, void* stream) {
  PQ3D_CHECK_ARG(A && W && C, "pq3d_linear_bf16: null operand");
  PQ3D_CHECK_ARG(M > 0 && N > 0 && K > 0 && groups > 0, "pq3d_linear_bf16: bad shape M=%d N=%d K=%d groups=%d", M, N,
                 K, groups);
  PQ3D_CHECK_ARG(K % kBlockK == 0, "pq3d_linear_bf16: K=%d must be a multiple of %d", K, kBlockK);
  PQ3D_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K,
                 "pq3d_linear_bf16: lda=%lld / ldw=%lld must be >= K and multiples of 8 (16-byte TMA strides)",
                 (long long)lda, (long long)ldw);
  PQ3D_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
                 "pq3d_linear_bf16: A / W must be 16-byte aligned");
  PQ3D_CHECK_ARG(a_rows_total >= (int64_t)(groups - 1) * a_group_rows + M &&
                     w_rows_total >= (int64_t)(groups - 1) * w_group_rows + N,
                 "pq3d_linear_bf16: group offsets exceed the operand extents");
  if (block_n == 0) {
    // skinny problems (few CTAs): narrow tiles for parallelism; big ones: 128x256 tiles
    const int64_t tiles256 = (int64_t)((M + 127) / 128) * ((N + 255) / 256) * groups;
    const int64_t tiles128 = (int64_t)((M + 127) / 128) * ((N + 127) / 128) * groups;
    block_n = tiles256 >= 2 * sm_count() ? 256 : (tiles128 >= sm_count() ? 128 : 64);
  }
  PQ3D_CHECK_ARG(block_n == 64 || block_n == 128 || block_n == 256, "pq3d_linear_bf16: block_n=%d not in {64,128,256}",
                 block_n);

  CUtensorMap ta, tw;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)a_rows_total};
    uint64_t strides[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)kBlockM};
    int rc = make_tmap_bf16(&ta, A, 2, dims, strides, box);
    if (rc != PQ3D_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)w_rows_total};
    uint64_t strides[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)block_n};
    int rc = make_tmap_bf16(&tw, W, 2, dims, strides, box);
    if (rc != PQ3D_OK) return rc;
  }
  LinearParams p;
  p.C = C;
  p.bias = bias;
  p.row_zero = row_zero;
  p.ldc = ldc;
  p.c_group_stride = c_group_stride;
  p.bias_group_stride = bias_group_stride;
  p.row_zero_group_stride = row_zero_group_stride;
  p.a_group_rows = (int32_t)a_group_rows;
  p.w_group_rows = (int32_t)w_group_rows;
  p.M = M;
  p.N = N;
  p.K = K;
  p.out_fp32 = out_fp32;
  p.bias_along_m = bias_along_m;
  p.relu = relu;
  p.alpha_ncols = alpha_ncols;
  p.alpha = alpha;
  const int esz = out_fp32 ? 4 : 2;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && ((ldc * esz) % 16 == 0) &&
             ((c_group_stride * esz) % 16 == 0);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (block_n) {
    case 64: return launch_linear<64>(ta, tw, p, groups, st);
    case 128: return launch_linear<128>(ta, tw, p, groups, st);
    default: return launch_linear<256>(ta, tw, p, groups, st);
  }
}
