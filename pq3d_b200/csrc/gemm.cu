// pq3d_linear_bf16: C = epilogue(A · Wᵀ) on tcgen05 tensor cores, operands staged by TMA.
//
// This is every nn.Linear on the decoder hot path — the cross-attention K/V/Q in-projections and
// out-projection (torch/nn/functional.py:5867-5873, :6653), the spatial self-attention
// w_qs/w_ks/w_vs/fc (modules/layers/transformers.py:180-185), the FFN (query_encoder.py:384) and the
// mask-head projections (modules/heads/mask_head.py:46-57).  The K/V projection is 85 % of the
// decoder's FLOPs, so this is the dominant kernel.
//
// Persistent: grid = min(#tiles, #SMs); each CTA walks 128 x BN output tiles (n fastest, so CTAs
// running at the same time share A rows in L2).  Warp roles (192 threads):
//   warp 0      TMA producer: A tile [128 x 64] and W tile [BN x 64] per k-block, 128B-swizzled,
//               into a 192 KB shared-memory ring guarded by full/empty mbarriers; runs ahead across
//               tile boundaries
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue (M=128, N=BN, K=16, bf16 -> fp32)
//               into one of TWO accumulator buffers in TMEM; tcgen05.commit releases ring slots and
//               publishes the finished accumulator
//   warps 2..5  epilogue, overlapped with the next tile's main loop: tcgen05.ld (lane = row),
//               (acc + bias) * alpha, ReLU / row zeroing, convert, stage one [32 rows x 128 B] box
//               per warp in swizzled shared memory (double-buffered) and TMA-store it; tails in M / N
//               are clipped by the tensor map.  C whose leading dimension is not 16-byte granular
//               (the 201-class logits) takes a direct-store path.
// Groups shift the A / W / C / bias bases: per-memory out-projections, per-layer multi-scale voxel
// K/V projections, per-scene V^T and mask-logit products run as one launch.
#include <cstdlib>
#include <cstring>

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
  int32_t num_m, num_n, num_tiles;
  int32_t out_fp32;
  int32_t bias_along_m;
  int32_t relu;
  int32_t alpha_ncols;      // columns n < alpha_ncols are scaled by alpha (n >= : unscaled)
  float alpha;
  int32_t tma_store;        // C is 16-byte granular: epilogue goes through shared memory + TMA store
  unsigned long long* dbg;  // optional per-CTA timeline (pq3d_debug_set_timeline), 8 slots per CTA
  int32_t batched;          // operands / result addressed through 4-D tensor maps {inner, rows, g2, g1} (pq3d_bgemm_bf16)
  int32_t G2;               // size of the inner group level when batched (group g -> g1 = g / G2, g2 = g % G2)
  int32_t dbg_flags;        // PQ3D_GEMM_DEBUG: 1 = skip the TMA stores, 2 = skip staging + stores (WRONG RESULTS; timing only)
  int32_t w_prefetch;       // W is constant (weights): its first ring pass may be fetched before griddepcontrol.wait
  int32_t use_a_off;        // group g's A rows start at a_off[g] instead of g * a_group_rows (groups <= kMaxAOff)
  int32_t a_off[8];
};
constexpr int kMaxAOff = 8;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle atom row
constexpr int kGemmThreads = 192;      // EW = 1: TMA warp + MMA warp + 4 epilogue warps; EW = 2: 8 epilogue warps (320 threads)
constexpr int kBoxBytes = 32 * 128;  // one epilogue box: 32 rows x 128 B

constexpr int kFullRing = 196608;   // 192 KB of operands in flight: one CTA per SM
constexpr int kHalfRing = 98304;    // 96 KB: two CTAs (of different streams' kernels) fit one SM

template <int BN, int CL, int RING = kFullRing>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = (BN / CL) * kBlockK * 2;                    // a CTA pair splits the W tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = RING / kStageBytes;
  static constexpr int kRingBytes = kStages * kStageBytes;
  static constexpr int kStagingBytes = 4 * 2 * kBoxBytes;                    // EW = 1: 4 warps x double buffer; EW = 2: 8 warps x 1
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kRingBytes + kStagingBytes + kBarBytes + 2 * BN * 4;
  static constexpr uint32_t kTmemCols = 2 * BN;                              // two accumulator buffers
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB shared-memory limit");
};

// CL = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) computes a 256 x BN tile.  Each CTA stages its own 128
// rows of A and HALF of the W tile; the leader CTA's single MMA thread issues M = 256 instructions that read
// both shared memories and write both tensor memories.  With one CTA per tile the shared-memory port is the
// limit (MMA operand reads 96 B/clk + TMA fill 96 B/clk against 128 B/clk: measured 66 % of the MMA floor, and
// multicasting W into both CTAs changed nothing); the pair halves the W traffic on both sides of that port.
// MC = 4 (with CL = 1, opt-in): a cluster of four CTAs owns four NEIGHBOURING column tiles of one row tile.  They need
// the same A tile, so each fetches a quarter of its rows and multicasts it into all four shared memories: per k-block a
// CTA pulls 4 KB of A + its 8 KB of W instead of 16 + 8 KB.  A ring slot is reusable once all four CTAs consumed it
// (multicast commit).  Correct (the kernel checks run it) but measured slower for the query-side GEMMs, see the host side.
template <int BN, int CL, int MC = 1, int RING = kFullRing, int EW = 1>
__global__ void __launch_bounds__(64 + 128 * EW, 1)
linear_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                   const __grid_constant__ CUtensorMap tmap_c, const LinearParams p) {
  using Cfg = GemmCfg<BN, CL, RING>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* staging = smem + Cfg::kRingBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;   // accumulator buffer ready   (MMA -> epilogue)
  uint64_t* tempty_bar = tfull_bar + 2;             // accumulator buffer drained (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(staging + Cfg::kStagingBytes + Cfg::kBarBytes);

  const int warp = threadIdx.x >> 5;
  const int num_kb = p.K / kBlockK;
  static_assert(MC == 1 || CL == 1, "multicast clusters use cta_group::1");
  constexpr int kClusterSize = CL * MC;
  const int crank = kClusterSize > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int rank = CL > 1 ? crank : 0;                                  // position inside a cta_group::2 pair
  // p.num_tiles counts the units a CLUSTER walks: tiles (CL = MC = 1), tile pairs (CL = 2) or groups of MC column tiles
  const int first_tile = blockIdx.x / kClusterSize, tile_step = gridDim.x / kClusterSize;
  const int n_units = p.num_n / MC;                                     // column units per row tile
  const int tiles_per_group = ((p.num_m + CL - 1) / CL) * n_units;
  auto m_tile_of = [&](int rem) { return (rem / n_units) * CL + rank; };
  auto n_tile_of = [&](int rem) { return (rem % n_units) * MC + (MC > 1 ? crank : 0); };

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("pq3d: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    if (p.tma_store) tma_prefetch_desc(&tmap_c);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], MC);      // multicast clusters: every CTA that received the slot's A rows frees it
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4 * EW * CL);   // pair: the leader's barrier collects the epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CL > 1) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot); else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  if (kClusterSize > 1) cluster_sync_all(); else __syncthreads();   // peers' barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above overlapped the previous kernel's tail.  When the caller declares W
  // a constant operand (weights: written before the stream's current chain of kernels began), the producer also
  // issues the W tiles of the first ring pass BEFORE waiting on the previous kernel — only A (activations) depends on
  // it — so the weight fetch (HBM latency + ~100 KB per CTA) hides behind the predecessor's tail as well.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  int w_prefetched = 0;
  if (p.w_prefetch && CL == 1 && !p.batched && warp == 0 && elect_one()) {
    const int t = first_tile;
    if (t < p.num_tiles) {
      const int g = t / tiles_per_group, rem = t % tiles_per_group;
      const int w_row = g * p.w_group_rows + n_tile_of(rem) * BN;
      const int n_pre = num_kb < Cfg::kStages ? num_kb : Cfg::kStages;
      for (int kb = 0; kb < n_pre; ++kb) {
        mbar_arrive_expect_tx(&full_bar[kb], Cfg::kStageBytes);      // A's bytes complete the phase later
        tma_load_2d(smem + kb * Cfg::kStageBytes + Cfg::kABytes, &tmap_w, &full_bar[kb], kb * kBlockK, w_row);
      }
      w_prefetched = n_pre;
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  unsigned long long* dbg = p.dbg == nullptr ? nullptr : p.dbg + 8ull * blockIdx.x;
  if (dbg != nullptr && threadIdx.x == 0) {
    dbg[0] = global_timer_ns();
    dbg[1] = clock64();
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    dbg[7] = smid;
  }

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int it = 0;
      for (int t = first_tile; t < p.num_tiles; t += tile_step) {
        const int g = t / tiles_per_group, rem = t % tiles_per_group;
        const int a_row = (p.use_a_off ? p.a_off[g] : g * p.a_group_rows) + m_tile_of(rem) * kBlockM;
        const int w_row = g * p.w_group_rows + n_tile_of(rem) * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          mbar_wait(&empty_bar[s], ((it / Cfg::kStages) & 1) ^ 1, 100 + s);
          uint8_t* sa = smem + s * Cfg::kStageBytes;
          if (CL > 1) {
            // both CTAs' bytes are counted on the LEADER's barrier (its MMA thread is the only consumer)
            const uint32_t lead_bar = mapa_u32(smem_u32(&full_bar[s]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], CL * Cfg::kStageBytes);
            tma_load_2d_pair(sa, &tmap_a, lead_bar, kb * kBlockK, a_row);
            tma_load_2d_pair(sa + Cfg::kABytes, &tmap_w, lead_bar, kb * kBlockK, w_row + rank * (BN / CL));
          } else if (MC > 1) {
            // this CTA's quarter of the A rows goes to every CTA of the cluster; W is private
            constexpr uint16_t kMcMask = static_cast<uint16_t>((1u << MC) - 1);
            if (it >= w_prefetched) {
              mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
              tma_load_2d(sa + Cfg::kABytes, &tmap_w, &full_bar[s], kb * kBlockK, w_row);
            }
            tma_load_2d_multicast(sa + crank * (Cfg::kABytes / MC), &tmap_a, &full_bar[s], kb * kBlockK,
                                  a_row + crank * (kBlockM / MC), kMcMask);
          } else if (it < w_prefetched) {
            tma_load_2d(sa, &tmap_a, &full_bar[s], kb * kBlockK, a_row);      // W of this slot is already in flight
          } else {
            mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
            if (p.batched) {      // strided batches: per-group offsets live in the tensor maps' outer dimensions
              const int m_tile = (rem / p.num_n) * kBlockM, n_tile = (rem % p.num_n) * BN;
              tma_load_4d(sa, &tmap_a, &full_bar[s], kb * kBlockK, m_tile, g % p.G2, g / p.G2);
              tma_load_4d(sa + Cfg::kABytes, &tmap_w, &full_bar[s], kb * kBlockK, n_tile, g % p.G2, g / p.G2);
            } else {
              tma_load_2d(sa, &tmap_a, &full_bar[s], kb * kBlockK, a_row);
              tma_load_2d(sa + Cfg::kABytes, &tmap_w, &full_bar[s], kb * kBlockK, w_row);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (pair: leader CTA only)
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM * CL, BN);
      constexpr uint16_t kPairMask = static_cast<uint16_t>((1u << CL) - 1);
      int it = 0, lt = 0;
      for (int t = first_tile; t < p.num_tiles; t += tile_step, ++lt) {
        const int buf = lt & 1;
        mbar_wait(&tempty_bar[buf], ((lt >> 1) & 1) ^ 1, 210 + buf);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + buf * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          mbar_wait(&full_bar[s], (it / Cfg::kStages) & 1, 200 + s);
          tc_fence_after();
          if (dbg != nullptr && it == 0) dbg[2] = clock64();
          const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advancing 16 bf16 (32 B) along K inside the 128-byte swizzle atom = +32 B on the start address
            if (CL > 1)
              umma_ss_pair(tmem_acc, umma_desc_k_sw128(sa + k * 32), umma_desc_k_sw128(sb + k * 32), idesc,
                           (kb | k) != 0 ? 1u : 0u);
            else
              umma_ss(tmem_acc, umma_desc_k_sw128(sa + k * 32), umma_desc_k_sw128(sb + k * 32), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          }
          if (CL > 1) tc_commit_pair(&empty_bar[s], kPairMask);   // frees the slot in BOTH CTAs
          else if (MC > 1) tc_commit_multicast(&empty_bar[s], static_cast<uint16_t>((1u << MC) - 1));
          else tc_commit(&empty_bar[s]);
        }
        if (CL > 1) tc_commit_pair(&tfull_bar[buf], kPairMask);   // each CTA's epilogue reads its own TMEM half
        else tc_commit(&tfull_bar[buf]);
      }
      if (dbg != nullptr) {
        dbg[3] = clock64();
        dbg[4] = lt;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps 2..5 (EW = 2: 2..9)
    // EW = 2: two warps per TMEM lane quadrant, each taking every other box of the tile's columns — the epilogue of a
    // 128 x 256 tile (TMEM loads, bias, bf16 pack, staging stores, proxy fence, TMA store per box) is a dependent chain
    // of ~6 k cycles for one warp, the same as the tile's MMAs, so with four warps it paces the persistent loop
    const int quad = warp & 3;   // TMEM lane quadrant this warp may access
    const int eh = (warp - 2) >> 2;            // which of the quadrant's EW warps
    const int r = lane_id();
    constexpr int kBoxBufs = EW == 1 ? 2 : 1;  // staging boxes per warp
    uint8_t* stage = staging + (warp - 2) * kBoxBufs * kBoxBytes;
    int lt = 0, bx = 0;
    for (int t = first_tile; t < p.num_tiles; t += tile_step, ++lt) {
      const int g = t / tiles_per_group, rem = t % tiles_per_group;
      const int m0 = m_tile_of(rem) * kBlockM, n0 = n_tile_of(rem) * BN;
      const int buf = lt & 1;
      float* bias_s = s_bias + buf * BN;
      for (int c = threadIdx.x - 64; c < BN; c += 128 * EW) {
        const int n = n0 + c;
        bias_s[c] = (p.bias != nullptr && !p.bias_along_m && n < p.N) ? p.bias[g * p.bias_group_stride + n] : 0.f;
      }
      if (EW == 1) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 1, 256;" ::: "memory");
      const int m = m0 + quad * 32 + r;
      const bool row_ok = m < p.M;
      const bool zero_row = row_ok && p.row_zero != nullptr && p.row_zero[g * p.row_zero_group_stride + m] != 0;
      const float bias_m = (p.bias != nullptr && p.bias_along_m && row_ok) ? p.bias[g * p.bias_group_stride + m] : 0.f;
      const float lo = p.relu ? 0.f : __int_as_float(0xff800000);
      const bool plain = !p.relu && p.row_zero == nullptr;      // warp-uniform: (acc + bias) only
      mbar_wait(&tfull_bar[buf], (lt >> 1) & 1, 300 + buf);
      tc_fence_after();
      const uint32_t tmem_row = tmem_base + buf * BN + (static_cast<uint32_t>(quad * 32) << 16);

      // 32 accumulator columns [c0, c0+32) of this thread's row -> finished fp32 values
      auto load_chunk = [&](int c0, float (&v)[32]) {
        uint32_t acc[32];
        tmem_ld_32x32(tmem_row + c0, acc);
        tmem_ld_wait();
        const bool uniform = (n0 + c0 + 32 <= p.alpha_ncols) || (n0 + c0 >= p.alpha_ncols);
        const float sc = (n0 + c0 < p.alpha_ncols) ? p.alpha : 1.f;
        // the epilogue warps are the throughput limit of the persistent schedule (one warp per SMSP):
        // the plain bias-add case (K / V projections, out-projections) must cost one FADD per element
        if (plain && sc == 1.f) {
          if (p.bias_along_m) {                                  // column bias is all zero
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) + bias_m;
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + j);
              v[j] = __uint_as_float(acc[j]) + b4.x;
              v[j + 1] = __uint_as_float(acc[j + 1]) + b4.y;
              v[j + 2] = __uint_as_float(acc[j + 2]) + b4.z;
              v[j + 3] = __uint_as_float(acc[j + 3]) + b4.w;
            }
          }
          return;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + j);
          const float bj[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float x = __uint_as_float(acc[j + q]) + bj[q] + bias_m;
            x *= uniform ? sc : ((n0 + c0 + j + q < p.alpha_ncols) ? p.alpha : 1.f);
            x = fmaxf(x, lo);
            v[j + q] = zero_row ? 0.f : x;
          }
        }
      };
      auto release_accumulator = [&]() {
        tc_fence_before();
        __syncwarp();
        if (r == 0) {
          if (CL > 1) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[buf]), 0));   // the leader's barrier
          else mbar_arrive(&tempty_bar[buf]);
        }
      };
      auto claim_box = [&]() -> uint8_t* {   // staging: wait until the box used kBoxBufs stores ago was read
        if (kBoxBufs == 2) {
          if (bx >= 2 && r == 0) tma_store_wait_read<1>();
        } else {
          if (bx >= 1 && r == 0) tma_store_wait_read<0>();
        }
        __syncwarp();
        return stage + (bx++ & (kBoxBufs - 1)) * kBoxBytes;
      };
      auto store_box = [&](uint8_t* box, int c0) {
        fence_proxy_async_smem();
        __syncwarp();
        if (r == 0 && !(p.dbg_flags & 3)) {
          if (p.batched) tma_store_4d(&tmap_c, box, n0 + c0, m0 + quad * 32, g % p.G2, g / p.G2);
          else tma_store_3d(&tmap_c, box, n0 + c0, m0 + quad * 32, g);
          tma_store_commit();
        }
      };

      if (p.tma_store && p.out_fp32) {
        const int n_boxes = min(BN, p.N - n0 + 31) / 32;            // one box = 32 fp32 columns
        if (eh >= n_boxes) release_accumulator();                   // nothing to read for this warp
        for (int b = eh; b < n_boxes; b += EW) {
          float v[32];
          load_chunk(b * 32, v);
          if (b + EW >= n_boxes) release_accumulator();
          uint8_t* box = claim_box() + r * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(box + ((j ^ (r & 7)) << 4)) =
                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          store_box(box - r * 128, b * 32);
        }
      } else if (p.tma_store) {
        const int n_boxes = min(BN, p.N - n0 + 63) / 64;            // one box = 64 bf16 columns = two TMEM loads
        if (eh >= n_boxes) release_accumulator();
        for (int b = eh; b < n_boxes; b += EW) {
          float v0[32], v1[32];
          load_chunk(b * 64, v0);
          load_chunk(b * 64 + 32, v1);
          if (b + EW >= n_boxes) release_accumulator();
          uint8_t* box = claim_box() + r * 128;
          if (p.dbg_flags & 2) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack_bf16x2(v0[8 * j + 0], v0[8 * j + 1]);
            u.y = pack_bf16x2(v0[8 * j + 2], v0[8 * j + 3]);
            u.z = pack_bf16x2(v0[8 * j + 4], v0[8 * j + 5]);
            u.w = pack_bf16x2(v0[8 * j + 6], v0[8 * j + 7]);
            *reinterpret_cast<uint4*>(box + ((j ^ (r & 7)) << 4)) = u;
            u.x = pack_bf16x2(v1[8 * j + 0], v1[8 * j + 1]);
            u.y = pack_bf16x2(v1[8 * j + 2], v1[8 * j + 3]);
            u.z = pack_bf16x2(v1[8 * j + 4], v1[8 * j + 5]);
            u.w = pack_bf16x2(v1[8 * j + 6], v1[8 * j + 7]);
            *reinterpret_cast<uint4*>(box + (((4 + j) ^ (r & 7)) << 4)) = u;
          }
          store_box(box - r * 128, b * 64);
        }
      } else {
        // Unaligned C (leading dimension not 16-byte granular): direct stores.
        uint8_t* c_row = reinterpret_cast<uint8_t*>(p.C) +
                         (static_cast<int64_t>(g) * p.c_group_stride + static_cast<int64_t>(m) * p.ldc) *
                             (p.out_fp32 ? 4 : 2);
        for (int c0 = eh * 32; c0 < BN; c0 += 32 * EW) {
          if (n0 + c0 >= p.N) break;
          float v[32];
          load_chunk(c0, v);
          if (!row_ok) continue;
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + c0 + j;
            if (n >= p.N) break;
            if (p.out_fp32) reinterpret_cast<float*>(c_row)[n] = v[j];
            else reinterpret_cast<__nv_bfloat16*>(c_row)[n] = __float2bfloat16_rn(v[j]);
          }
        }
        release_accumulator();
      }
    }
    if (r == 0) tma_store_wait_read<0>();   // bulk stores must finish READING smem before exit (writes complete at grid end)
    __syncwarp();
    tc_fence_before();
    if (dbg != nullptr && threadIdx.x == 64) dbg[5] = clock64();
  }
  if (kClusterSize > 1) cluster_sync_all(); else __syncthreads();   // no CTA leaves while its peer may still signal it
  if (warp == 1) {
    tc_fence_after();
    if (CL > 1) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base); else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }

  if (dbg != nullptr && threadIdx.x == 0) dbg[6] = clock64();
}

template <int BN, int CL, int MC = 1, int RING = kFullRing, int EW = 1>
static int launch_linear(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tc, const LinearParams& p,
                         cudaStream_t stream, int max_ctas = 0) {
  using Cfg = GemmCfg<BN, CL, RING>;
  constexpr int kCluster = CL * MC;
  static bool configured = false;
  if (!configured) {
    PQ3D_CUDA(cudaFuncSetAttribute(linear_bf16_kernel<BN, CL, MC, RING, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg::kSmemBytes));
    configured = true;
  }
  int sms = sm_count();
  if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;      // leave SMs to kernels running next to this one
  // max_ctas < 0: -max_ctas WAVES of CTAs instead of one persistent wave — each CTA walks fewer tiles and gives its SM
  // back earlier, so urgent kernels of other streams get in at that granularity (pq3d_set_launch_priority)
  if (max_ctas < 0) sms = sms * (-max_ctas);
  const int max_clusters = sms / kCluster > 0 ? sms / kCluster : 1;
  const int grid = kCluster * (p.num_tiles < max_clusters ? p.num_tiles : max_clusters);
  PQ3D_CUDA(launch_kernel_cluster(linear_bf16_kernel<BN, CL, MC, RING, EW>, dim3(grid), dim3(64 + 128 * EW),
                                  Cfg::kSmemBytes, stream, kCluster, ta, tw, tc, p));
  return PQ3D_OK;
}

}  // namespace pq3d

using namespace pq3d;

static unsigned long long* g_timeline = nullptr;
// Debug only: when set, every linear_bf16 CTA records {globaltimer at start, clock at start, first operands
// landed, last MMA issued, #tiles, epilogue done, exit, smid} into buf[8 * blockIdx.x].
extern "C" int pq3d_debug_set_timeline(void* buf) {
  g_timeline = reinterpret_cast<unsigned long long*>(buf);
  return PQ3D_OK;
}

static int linear_impl(const void* A, int64_t lda, int64_t a_rows_total, int64_t a_group_rows,
                                const void* W, int64_t ldw, int64_t w_rows_total, int64_t w_group_rows, void* C,
                                int64_t ldc, int64_t c_group_stride, int out_fp32, const float* bias,
                                int64_t bias_group_stride, int bias_along_m, const uint8_t* row_zero,
                                int64_t row_zero_group_stride, int M, int N, int K, int groups, float alpha,
                                int alpha_ncols, int relu, int block_n, void* stream, int max_ctas,
                                const int32_t* a_row_offsets, int flags) {
  const int w_is_constant = flags & 1;
  const bool no_pairs = (flags & 2) != 0;
  // bit 2: the caller runs several batches concurrently — the 64- / 128-wide tiles keep only 96 KB of operands in flight
  // so that two CTAs (of different streams' latency-bound kernels) share an SM
  static const int half_ring_mode = [] {      // PQ3D_GEMM_HALF_RING: 0 = never, 1 = when flags bit 2 is set (default), 2 = always
    const char* e = getenv("PQ3D_GEMM_HALF_RING");
    return e == nullptr ? 1 : atoi(e);
  }();
  // ... and only when the concurrent batches can actually fill the machine: with a handful of tiles per GEMM (small
  // scenes) SMs are free anyway and the deeper ring is faster (config 1: 0.152 vs 0.167 ms per step)
  const int64_t tiles_here = (int64_t)((M + kBlockM - 1) / kBlockM) * ((N + 63) / 64) * groups;
  const bool half_ring = half_ring_mode == 2 || (half_ring_mode == 1 && (flags & 4) != 0 && 4 * tiles_here > sm_count());
  PQ3D_CHECK_ARG(A && W && C, "pq3d_linear_bf16: null operand");
  PQ3D_CHECK_ARG(M > 0 && N > 0 && K > 0 && groups > 0, "pq3d_linear_bf16: bad shape M=%d N=%d K=%d groups=%d", M, N,
                 K, groups);
  PQ3D_CHECK_ARG(K % kBlockK == 0, "pq3d_linear_bf16: K=%d must be a multiple of %d", K, kBlockK);
  PQ3D_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K,
                 "pq3d_linear_bf16: lda=%lld / ldw=%lld must be >= K and multiples of 8 (16-byte TMA strides)",
                 (long long)lda, (long long)ldw);
  PQ3D_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
                 "pq3d_linear_bf16: A / W must be 16-byte aligned");
  PQ3D_CHECK_ARG((a_row_offsets != nullptr || a_rows_total >= (int64_t)(groups - 1) * a_group_rows + M) &&
                     w_rows_total >= (int64_t)(groups - 1) * w_group_rows + N,
                 "pq3d_linear_bf16: group offsets exceed the operand extents");
  PQ3D_CHECK_ARG(a_row_offsets == nullptr || groups <= kMaxAOff, "pq3d_linear_bf16_ex: a_row_offsets supports <= %d groups",
                 kMaxAOff);
  if (a_row_offsets != nullptr)
    for (int g = 0; g < groups; ++g)
      PQ3D_CHECK_ARG(a_row_offsets[g] >= 0 && a_rows_total >= (int64_t)a_row_offsets[g] + M,
                     "pq3d_linear_bf16_ex: a_row_offsets[%d]=%d exceeds the operand extent", g, a_row_offsets[g]);
  if (block_n == 0) {
    // big problems: 128x256 tiles; fewer than a wave of those: narrower tiles for parallelism
    const int64_t tiles256 = (int64_t)((M + 127) / 128) * ((N + 255) / 256) * groups;
    const int64_t tiles128 = (int64_t)((M + 127) / 128) * ((N + 127) / 128) * groups;
    block_n = tiles256 >= sm_count() ? 256 : (tiles128 >= sm_count() ? 128 : 64);
  }
  PQ3D_CHECK_ARG(block_n == 64 || block_n == 128 || block_n == 256, "pq3d_linear_bf16: block_n=%d not in {64,128,256}",
                 block_n);

  // CTA pairs (cta_group::2, 256 x 256 tiles) whenever there are at least two row tiles and more than a wave of work
  // (PQ3D_GEMM_CLUSTER=0 disables, for A/B measurements)
  static const bool cluster_ok = [] {
    const char* e = getenv("PQ3D_GEMM_CLUSTER");
    return e == nullptr || e[0] != '0';
  }();
  const int num_m_tiles = (M + kBlockM - 1) / kBlockM;
  const int cl = (cluster_ok && !no_pairs && a_row_offsets == nullptr && block_n == 256 && num_m_tiles >= 2 &&
                  (int64_t)num_m_tiles * ((N + 255) / 256) * groups >= 2 * sm_count()) ? 2 : 1;
  // Opt-in (PQ3D_GEMM_MULTICAST=1): clusters of 4 neighbouring column tiles share their A tile through TMA multicast.
  // MEASURED on B200 (round 2, M = 400 query rows, in-graph): it does NOT pay — N=768: 5.22 -> 5.44 us, N=2048 (128
  // CTAs): 5.27 -> 6.11 us, N=2304 (144 CTAs): 5.28 -> 10.8 us (clusters of 4 leave 16 of the 148 SMs unusable, so 144
  // CTAs need two waves).  These GEMMs are bound by a chain of latencies (launch, first TMA round trip, epilogue, store,
  // grid completion), not by the 288 KB each SM pulls; kept for shapes where A traffic does dominate.
  static const bool multicast_ok = [] {
    const char* e = getenv("PQ3D_GEMM_MULTICAST");
    return e != nullptr && e[0] == '1';
  }();
  constexpr int kMc = 4;
  const int n_tiles64 = (N + 63) / 64;
  const int mc = (multicast_ok && block_n == 64 && cl == 1 && n_tiles64 % kMc == 0 &&
                  (int64_t)num_m_tiles * n_tiles64 * groups <= 2 * sm_count()) ? kMc : 1;
  CUtensorMap ta, tw;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)a_rows_total};
    uint64_t strides[1] = {(uint64_t)lda * 2};
    uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)(kBlockM / mc)};   // multicast: each CTA fetches a quarter of the rows
    int rc = make_tmap_bf16(&ta, A, 2, dims, strides, box);
    if (rc != PQ3D_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)w_rows_total};
    uint64_t strides[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)(block_n / cl)};   // each CTA of a pair fetches its half
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
  p.num_m = (M + kBlockM - 1) / kBlockM;
  p.num_n = (N + block_n - 1) / block_n;
  p.num_tiles = ((p.num_m + cl - 1) / cl) * (p.num_n / mc) * groups;   // tile PAIRS when cl = 2, tile quads when mc = 4
  p.out_fp32 = out_fp32;
  p.bias_along_m = bias_along_m;
  p.relu = relu;
  p.alpha_ncols = alpha_ncols;
  p.alpha = alpha;
  p.dbg = g_timeline;
  static const int dbg_flags = [] {
    const char* e = getenv("PQ3D_GEMM_DEBUG");
    return e == nullptr ? 0 : atoi(e);
  }();
  p.dbg_flags = dbg_flags;
  p.batched = 0;
  p.G2 = 1;
  p.w_prefetch = (w_is_constant && pdl_enabled()) ? 1 : 0;
  p.use_a_off = a_row_offsets != nullptr;
  for (int g = 0; g < kMaxAOff; ++g) p.a_off[g] = (a_row_offsets != nullptr && g < groups) ? a_row_offsets[g] : 0;
  const int esz = out_fp32 ? 4 : 2;
  p.tma_store = ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && ((ldc * esz) % 16 == 0) &&
                ((c_group_stride * esz) % 16 == 0);
  // An epilogue that stores from registers without the shared-memory staging was built and MEASURED (round 2,
  // 3 x [8192 x 3072 x 768]): CTA pairs 1252 -> 984 TFLOP/s, single CTA 1180 -> 969 — a thread owns a row, so one store
  // instruction touches 32 different 128-byte lines and the epilogue warps become the limit.  Its mere presence as a
  // runtime branch also slowed the 64-wide tiles by 8 % (5.23 -> 5.64 us per chain GEMM: bigger kernel image), so the
  // code was removed again (git history: "register-layout (unstaged) epilogue").
  CUtensorMap tc = ta;   // placeholder when the direct-store path is taken
  if (p.tma_store) {
    // {N, M, groups}: tails in N and M are clipped by TMA, a tile never spills into the next group
    uint64_t dims[3] = {(uint64_t)N, (uint64_t)M, (uint64_t)groups};
    uint64_t gs = groups > 1 ? (uint64_t)c_group_stride * esz : (uint64_t)ldc * esz * (uint64_t)M;
    uint64_t strides[2] = {(uint64_t)ldc * esz, gs};
    uint32_t box[3] = {(uint32_t)(128 / esz), 32u, 1u};
    int rc = make_tmap(&tc, C, esz, 3, dims, strides, box);
    if (rc != PQ3D_OK) return rc;
  }
  static const int epi_warps = [] {      // PQ3D_GEMM_EPI_WARPS: epilogue warps per TMEM lane quadrant of the 256-wide tiles
    const char* e = getenv("PQ3D_GEMM_EPI_WARPS");
    return e == nullptr ? 2 : atoi(e);
  }();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (block_n) {
    case 64: return mc == kMc ? launch_linear<64, 1, kMc>(ta, tw, tc, p, st, max_ctas)
                   : half_ring ? launch_linear<64, 1, 1, kHalfRing>(ta, tw, tc, p, st, max_ctas)
                               : launch_linear<64, 1>(ta, tw, tc, p, st, max_ctas);
    case 128: return half_ring ? launch_linear<128, 1, 1, kHalfRing>(ta, tw, tc, p, st, max_ctas)
                               : launch_linear<128, 1>(ta, tw, tc, p, st, max_ctas);
    default:
      if (epi_warps == 2)
        return cl == 2 ? launch_linear<256, 2, 1, kFullRing, 2>(ta, tw, tc, p, st, max_ctas)
                       : launch_linear<256, 1, 1, kFullRing, 2>(ta, tw, tc, p, st, max_ctas);
      return cl == 2 ? launch_linear<256, 2>(ta, tw, tc, p, st, max_ctas)
                     : launch_linear<256, 1>(ta, tw, tc, p, st, max_ctas);
  }
}

extern "C" int pq3d_linear_bf16(const void* A, int64_t lda, int64_t a_rows_total, int64_t a_group_rows,
                                const void* W, int64_t ldw, int64_t w_rows_total, int64_t w_group_rows, void* C,
                                int64_t ldc, int64_t c_group_stride, int out_fp32, const float* bias,
                                int64_t bias_group_stride, int bias_along_m, const uint8_t* row_zero,
                                int64_t row_zero_group_stride, int M, int N, int K, int groups, float alpha,
                                int alpha_ncols, int relu, int block_n, void* stream) {
  return linear_impl(A, lda, a_rows_total, a_group_rows, W, ldw, w_rows_total, w_group_rows, C, ldc, c_group_stride,
                     out_fp32, bias, bias_group_stride, bias_along_m, row_zero, row_zero_group_stride, M, N, K, groups,
                     alpha, alpha_ncols, relu, block_n, stream, 0, nullptr, 0);
}

// pq3d_linear_bf16 with two scheduling / addressing extras:
//   max_ctas       > 0: the persistent grid uses at most this many CTAs (SMs), so a long GEMM can run NEXT TO a chain of
//                  small latency-bound kernels on another stream instead of holding every SM until it finishes
//   a_row_offsets  host array [groups] (<= 8) or NULL: group g's A rows start at a_row_offsets[g] (instead of
//                  g * a_group_rows) — groups that share or permute their A operand, e.g. the self-attention q / k / v
//                  projections reading (x + pos), (x + pos), x
//   flags          bit 0: W holds weights that no kernel of the current dependency chain writes — its first tiles are
//                  fetched before the programmatic-dependent-launch wait on the previous kernel (only A depends on it)
//                  bit 1: never use CTA pairs (cta_group::2 clusters).  Callers that keep SEVERAL graphs in flight on
//                  different streams set it: with pair GEMMs of different streams competing for SMs, long loops stalled
//                  on the device about once per 40 k decoder steps (one graph left spinning, no mbarrier timeout, GPU
//                  otherwise idle; 180 k steps clean with cta_group::1 only, 60 k clean with pairs on a single stream) —
//                  root cause not established, so concurrency and pairs are not combined.
extern "C" int pq3d_linear_bf16_ex(const void* A, int64_t lda, int64_t a_rows_total, int64_t a_group_rows,
                                   const void* W, int64_t ldw, int64_t w_rows_total, int64_t w_group_rows, void* C,
                                   int64_t ldc, int64_t c_group_stride, int out_fp32, const float* bias,
                                   int64_t bias_group_stride, int bias_along_m, const uint8_t* row_zero,
                                   int64_t row_zero_group_stride, int M, int N, int K, int groups, float alpha,
                                   int alpha_ncols, int relu, int block_n, int max_ctas, const int32_t* a_row_offsets,
                                   int flags, void* stream) {
  return linear_impl(A, lda, a_rows_total, a_group_rows, W, ldw, w_rows_total, w_group_rows, C, ldc, c_group_stride,
                     out_fp32, bias, bias_group_stride, bias_along_m, row_zero, row_zero_group_stride, M, N, K, groups,
                     alpha, alpha_ncols, relu, block_n, stream, max_ctas, a_row_offsets, flags);
}

// Strided batched GEMM: C[g1,g2] = alpha * A[g1,g2] · W[g1,g2]ᵀ for G1 x G2 independent problems whose operands are
// strided views (e.g. per-(scene, head) slices of [tokens, heads*64] tensors).  Same kernel; the per-group offsets are
// the outer two dimensions of 4-D tensor maps, so rows past M / N are zero-filled on load and clipped on store PER GROUP.
extern "C" int pq3d_bgemm_bf16(const void* A, int64_t a_row_stride, int64_t a_g2_stride, int64_t a_g1_stride,
                               const void* W, int64_t w_row_stride, int64_t w_g2_stride, int64_t w_g1_stride, void* C,
                               int64_t c_row_stride, int64_t c_g2_stride, int64_t c_g1_stride, int out_fp32, int M,
                               int N, int K, int G2, int G1, float alpha, int block_n, void* stream) {
  PQ3D_CHECK_ARG(A && W && C, "pq3d_bgemm_bf16: null operand");
  PQ3D_CHECK_ARG(M > 0 && N > 0 && K > 0 && G1 > 0 && G2 > 0 && K % kBlockK == 0,
                 "pq3d_bgemm_bf16: bad shape M=%d N=%d K=%d G1=%d G2=%d (K must be a multiple of %d)", M, N, K, G1, G2,
                 kBlockK);
  const int esz = out_fp32 ? 4 : 2;
  auto ok16 = [](int64_t elems, int sz) { return (elems * sz) % 16 == 0; };
  PQ3D_CHECK_ARG(ok16(a_row_stride, 2) && ok16(a_g2_stride, 2) && ok16(a_g1_stride, 2) && ok16(w_row_stride, 2) &&
                     ok16(w_g2_stride, 2) && ok16(w_g1_stride, 2) && ok16(c_row_stride, esz) &&
                     ok16(c_g2_stride, esz) && ok16(c_g1_stride, esz) &&
                     (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(C) & 15) == 0,
                 "pq3d_bgemm_bf16: every base pointer and stride must be 16-byte granular");
  if (block_n == 0) {
    const int64_t t256 = (int64_t)((M + 127) / 128) * ((N + 255) / 256) * G1 * G2;
    const int64_t t128 = (int64_t)((M + 127) / 128) * ((N + 127) / 128) * G1 * G2;
    block_n = (N > 128 && t256 >= sm_count()) ? 256 : ((N > 64 && t128 >= sm_count()) ? 128 : 64);
  }
  PQ3D_CHECK_ARG(block_n == 64 || block_n == 128 || block_n == 256, "pq3d_bgemm_bf16: block_n=%d", block_n);
  // a stride of 0 elements (broadcast operand) is not representable in a tensor map: use a dimension of extent 1
  auto dim_of = [](int64_t stride, int n) { return stride == 0 ? 1 : n; };
  auto str_of = [](int64_t stride, int64_t fallback) { return stride == 0 ? fallback : stride; };
  CUtensorMap ta, tw, tc;
  {
    uint64_t dims[4] = {(uint64_t)K, (uint64_t)M, (uint64_t)dim_of(a_g2_stride, G2), (uint64_t)dim_of(a_g1_stride, G1)};
    uint64_t strides[3] = {(uint64_t)a_row_stride * 2, (uint64_t)str_of(a_g2_stride, a_row_stride) * 2,
                           (uint64_t)str_of(a_g1_stride, a_row_stride) * 2};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)kBlockM, 1u, 1u};
    int rc = make_tmap_bf16(&ta, A, 4, dims, strides, box);
    if (rc != PQ3D_OK) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)K, (uint64_t)N, (uint64_t)dim_of(w_g2_stride, G2), (uint64_t)dim_of(w_g1_stride, G1)};
    uint64_t strides[3] = {(uint64_t)w_row_stride * 2, (uint64_t)str_of(w_g2_stride, w_row_stride) * 2,
                           (uint64_t)str_of(w_g1_stride, w_row_stride) * 2};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)block_n, 1u, 1u};
    int rc = make_tmap_bf16(&tw, W, 4, dims, strides, box);
    if (rc != PQ3D_OK) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)N, (uint64_t)M, (uint64_t)G2, (uint64_t)G1};
    uint64_t strides[3] = {(uint64_t)c_row_stride * esz, (uint64_t)c_g2_stride * esz, (uint64_t)c_g1_stride * esz};
    uint32_t box[4] = {(uint32_t)(128 / esz), 32u, 1u, 1u};
    int rc = make_tmap(&tc, C, esz, 4, dims, strides, box);
    if (rc != PQ3D_OK) return rc;
  }
  PQ3D_CHECK_ARG(a_g2_stride != 0 && a_g1_stride != 0 && w_g2_stride != 0 && w_g1_stride != 0,
                 "pq3d_bgemm_bf16: broadcast (zero-stride) operands are not supported; pass G=1 for that level");
  LinearParams p;
  memset(&p, 0, sizeof(p));
  p.C = C;
  p.M = M;
  p.N = N;
  p.K = K;
  p.num_m = (M + kBlockM - 1) / kBlockM;
  p.num_n = (N + block_n - 1) / block_n;
  p.num_tiles = p.num_m * p.num_n * G1 * G2;
  p.out_fp32 = out_fp32;
  p.alpha = alpha;
  p.alpha_ncols = alpha == 1.f ? 0 : N;
  p.tma_store = 1;
  p.batched = 1;
  p.G2 = G2;
  p.ldc = c_row_stride;
  p.dbg = nullptr;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (block_n) {
    case 64: return launch_linear<64, 1>(ta, tw, tc, p, st);
    case 128: return launch_linear<128, 1>(ta, tw, tc, p, st);
    default: return launch_linear<256, 1>(ta, tw, tc, p, st);
  }
}
