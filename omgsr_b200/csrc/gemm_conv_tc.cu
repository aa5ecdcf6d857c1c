// Implicit-GEMM convolution / linear layer / batched attention GEMM on the 5th-gen tensor cores
// (tcgen05 + TMEM), fed by TMA.
//
//   D[pixel, cout] = epilogue( sum_{tap, cin} A[pixel + tap, cin] * Wt[cout, tap*Cin + cin] )
//
// * Activations are NHWC bf16. The A operand of a 128-pixel tile for one (tap, 64-channel slice) is ONE
//   TMA box {64 ch, tw, th, tb} of the input tensor shifted by the tap offset; TMA's out-of-bounds zero
//   fill implements the convolution padding, so no im2col buffer is ever materialised. The box lands in
//   shared memory as 128 rows x 128 B with the 128-byte swizzle, which is exactly the K-major UMMA
//   operand layout. Stride-2 convolutions use a 5-D space-to-depth view of the same tensor
//   {pixel pair, W/2, row parity, H/2, B}; nearest-2x-upsample + 3x3 is algebraically collapsed on the host into four
//   2x2 convolutions (one per output parity) and uses the stride-1 path with strided output addressing.
// * Weights are [Cout, R*S*Cin] bf16 (K-major), one TMA box {64, BN} per k-block. For the batched
//   attention GEMMs the B operand is a per-(batch, head) matrix, either K-major (Q K^T) or MN-major
//   (P V, V read in place as [kv, d]).
// * Warp-specialised persistent kernel, one CTA per SM: warp 0 = TMA producer, warp 1 = MMA issuer
//   (single elected thread, tcgen05.mma, M=128 per CTA, N=BN, K=16), warps 2..5 = epilogue
//   (tcgen05.ld -> bias/activation/gate/residual -> bf16|f32 -> global, epilogue.cuh). smem ring of kStages, two TMEM
//   accumulator stages (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// * Two forms: conv_gemm_tc2_kernel (default) pairs two CTAs of a cluster on one cta_group::2 MMA (M = 256: each CTA
//   loads its own A tile and HALF of the weight tile; 92.6 % tensor-pipe activity at N = 256), conv_gemm_tc_kernel is
//   the single-CTA form for batched / MN-major B operands, ragged Cout and one-tile problems. The tile width BN is
//   chosen by a small cost model (pick_bn). A plain GEMM may have K that does not fill the last 64-deep k-block.
//
// Replaces the cuDNN / cuBLAS calls underneath diffusers' Conv2d / Linear modules that the reference
// reaches from infer/omgsr_s_infer_model.py:75-84,173 and infer/omgsr_f_infer_model.py:16,191-211.
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include "epilogue.cuh"
#include "../../include/omgsr_b200.h"

namespace omgsr {

struct ConvGemmParams {
  int tiles_x, tiles_y, tiles_b;  // tile grid over (W, H, B) of the GEMM-side pixel space
  int log_tw, log_th;             // tile extents are powers of two, tw*th*tb == 128
  int W, H, B;                    // GEMM-side pixel extents
  int n_tiles, Cout;
  int kc_per_tap, R, S, pad_t, pad_l;
  int stride2, pair_off;          // stride-2 view: channel offset of the odd pixel of a pair (= input ld)
  int b_batched;                  // B operand has (head, batch) coordinates taken from the tile (y, b)
  int b_mn_major;                 // B operand is [K, N] row-major (attention P*V)
  long long o_sx, o_sy, o_sb;     // output element strides per GEMM-side x / y / b
  long long r_sx, r_sy, r_sb;     // residual strides
  void* out;
  const float* bias;   // [Cout] or null
  const bf16* res;     // or null
  const float* gate;   // [B or M/gate_rows, Cout_out] or null
  int gate_rows;       // linear mode: rows per gate vector (tokens per image)
  long long gate_ld;   // elements between consecutive gate vectors
  int act_c0, act_c1;  // the activation is applied to output columns in [act_c0, act_c1) only
  int act;             // OMGSR_ACT_*
  int out_f32;
  int out_vec32, res_vec32;  // output / residual rows allow 256-bit accesses (see epilogue.cuh)
  int res_f32;               // the residual is fp32 (fp32 residual stream, together with out_f32)
  int a_bytes;         // bytes one A box delivers (tb may be clamped to B)
  RopeArgs rope;       // rope.cos_t != null: q / k head columns get RMSNorm + RoPE in the epilogue (FLUX projections)
  float* gn_sums;      // optional fused GroupNorm statistics of the output: [B][groups][2]
  int gn_cpg, gn_groups, gn_slots;  // gn_sums: scratch [image][gn_slots][groups * 2], see gn_flush (epilogue.cuh)
};

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kThreads = 192;
constexpr int kAccStride = 256;  // TMEM columns per accumulator stage

template <int BN>
struct Cfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 160 ? 5 : 6);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ +
                                    1024 /*fused GroupNorm partial sums: 4 warps x 64 floats*/ +
                                    16 * BN /*bias slices: 4 warps x BN floats*/;
};

struct TileCoord {
  int n_tile, tx, ty, tbi;
};
__device__ __forceinline__ TileCoord decode_tile(int tile, const ConvGemmParams& p) {
  TileCoord t;
  t.n_tile = tile % p.n_tiles;
  const int m = tile / p.n_tiles;
  t.tx = m % p.tiles_x;
  t.ty = (m / p.tiles_x) % p.tiles_y;
  t.tbi = m / (p.tiles_x * p.tiles_y);
  return t;
}

// accumulator row of a rope-mode GEMM: head by head (128 columns), q / k heads through epilogue_head_rope, the other
// columns (v, the single blocks' MLP half) through the ordinary specialised routine
template <int BN>
__device__ __forceinline__ void epilogue_row_rope(const ConvGemmParams& p, int epi_mode, uint32_t t_row, int n_tile, int x,
                                                  bool valid, EpiArgs& e, const float* bias_s, uint4 (&rr)[4], int lane) {
  static_assert(BN % 128 == 0 || BN < 128, "rope mode needs whole heads per tile");
  const long long pos = p.rope.pos_offset + x;
#pragma unroll 1
  for (int hh = 0; hh < BN / 128; ++hh) {
    const int cv = n_tile * BN + hh * 128;
    const bool is_q = cv >= p.rope.q0 && cv < p.rope.q1, is_k = cv >= p.rope.k0 && cv < p.rope.k1;
    if (is_q || is_k) {
      epilogue_head_rope(t_row + hh * 128, bias_s ? bias_s + hh * 128 : nullptr, is_q ? p.rope.wq : p.rope.wk,
                         p.rope.cos_t + pos * 128, p.rope.sin_t + pos * 128,
                         static_cast<bf16*>(e.out) + e.o_off + cv, valid, p.rope.eps);
    } else {
      epilogue_row_m<false>(epi_mode, t_row + hh * 128, 128, 0, cv, cv, valid, e, nullptr, 0, false,
                            bias_s ? bias_s + hh * 128 : nullptr, rr, nullptr, 0, lane);
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const ConvGemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled operand tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // offset, not a cast: keeps the shared address space (LDS / STS)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;                        // [kStages]
  uint64_t* empty_bar = bars + C::kStages;          // [kStages]
  uint64_t* acc_full = bars + 2 * C::kStages;       // [2]
  uint64_t* acc_empty = bars + 2 * C::kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 4);
  float* gn_acc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [4 warps][64]

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform to the compiler (see the MMA issuer)
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_launch_dependents();
  pdl_wait();  // nothing above touches global memory

  const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_b * p.n_tiles;
  const int num_kb = p.R * p.S * p.kc_per_tap;
  const int tw = 1 << p.log_tw, th = 1 << p.log_th;
  const int tb = kBM >> (p.log_tw + p.log_th);

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(tile, p);
        const int x0 = t.tx * tw, y0 = t.ty * th, b0 = t.tbi * tb;
        const int bh = p.b_batched ? t.ty : 0, bb = p.b_batched ? t.tbi : 0;
        int kb = 0;
        for (int r = 0; r < p.R; ++r) {
          for (int s = 0; s < p.S; ++s) {
            int c1, c2, c3, c4, cbase;
            if (!p.stride2) {
              cbase = 0;
              c1 = x0 + s - p.pad_l;
              c2 = y0 + r - p.pad_t;
              c3 = b0;
              c4 = 0;
            } else {
              // input (iy, ix) = (2*oy + r - pad_t, 2*ox + s - pad_l) in the {2C, W/2, 2, H/2, B} view
              const int dy = r - p.pad_t, dx = s - p.pad_l;
              const int py = dy & 1, px = dx & 1;
              cbase = px * p.pair_off;
              c1 = x0 + ((dx - px) >> 1);
              c2 = py;
              c3 = y0 + ((dy - py) >> 1);
              c4 = b0;
            }
            for (int kc = 0; kc < p.kc_per_tap; ++kc, ++kb) {
              mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = smem + stage * C::kStageBytes;
              uint8_t* sb = sa + C::kABytes;
              mbar_expect_tx(&full_bar[stage], p.a_bytes + C::kBBytes);
              tma_load_5d(sa, &tmA, &full_bar[stage], cbase + kc * kBK, c1, c2, c3, c4);
              if (!p.b_mn_major) {
                tma_load_4d(sb, &tmB, &full_bar[stage], kb * kBK, t.n_tile * BN, bh, bb);
              } else {
                // [K, N] row-major operand: one {64 n, 64 k} box per 64-wide N atom
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                  tma_load_4d(sb + j * (64 * kBK * 2), &tmB, &full_bar[stage], t.n_tile * BN + j * 64, kb * kBK, bh,
                              bb);
              }
              if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // whole warp on warp-uniform values, one elected lane issues (conv_gemm_tc2_kernel has the argument)
    {
      const bool leader = elect_one();
      const uint32_t idesc = umma_idesc_bf16(kBM, BN, 0, p.b_mn_major);
      const uint64_t da0 = umma_desc_sw128(0, 16, 1024);
      // MN-major SW128 B: atoms of {64 n x 8 k}; LBO = next 64-wide N atom, SBO = next 8 k-rows
      const uint64_t db0 = p.b_mn_major ? umma_desc_sw128(0, 64 * kBK * 2, 1024) : da0;
      const uint32_t a_hi = static_cast<uint32_t>(da0 >> 32), a_lo0 = static_cast<uint32_t>(da0);
      const uint32_t b_hi = static_cast<uint32_t>(db0 >> 32), b_lo0 = static_cast<uint32_t>(db0);
      // advance 16 bf16 along K: K-major +32 B inside the swizzle atom (+2 in the addr >> 4 field); MN-major two
      // 1024-byte row groups (+128)
      const uint32_t b_step = p.b_mn_major ? 128u : 2u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t a_lo = a_lo0 | (sa >> 4), b_lo = b_lo0 | ((sa + C::kABytes) >> 4);
          if (leader) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              tc_mma_ss(d_tmem, (static_cast<uint64_t>(a_hi) << 32) | (a_lo + 2 * k),
                        (static_cast<uint64_t>(b_hi) << 32) | (b_lo + b_step * k), idesc, (kb | k) != 0);
            tc_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (leader) tc_commit(&acc_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue (warps 2..5) =================
    const int quarter = warp & 3;         // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;  // accumulator row == pixel index inside the tile
    const int xl = row & (tw - 1);
    const int yl = (row >> p.log_tw) & (th - 1);
    const int bl = row >> (p.log_tw + p.log_th);
    const bool geglu = (p.act == OMGSR_ACT_GEGLU);
    const int out_bn = geglu ? BN / 2 : BN;
    const int cout_out = geglu ? p.Cout / 2 : p.Cout;
    int acc = 0;
    uint32_t acc_phase = 0;
    // fused GroupNorm statistics: this warp's shared-memory slice holds the sums of the groups of ONE (image,
    // n_tile) and is flushed to HBM (atomics) only when that pair changes: consecutive tiles of a persistent CTA
    // mostly stay within one image, so atomics drop by ~14x versus one flush per tile.
    float* my_gn = gn_acc + quarter * 64;
    float* my_bias = gn_acc + 256 + quarter * BN;
    int bias_tile = -1;
    const int epi_mode = epilogue_mode(p.act, p.gate != nullptr, p.res != nullptr, p.out_f32 != 0, p.gn_sums != nullptr,
                                       p.res_f32 != 0);
    const int gn_n = p.gn_groups * 2;  // <= 64: the slice holds every group of one image, whatever the n_tile
    int gn_b = -1;
    if (p.gn_sums) { my_gn[lane] = 0.f; my_gn[lane + 32] = 0.f; __syncwarp(); }
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(tile, p);
      if (p.bias && t.n_tile != bias_tile) {  // first tile only: later ones are prefetched below
        stage_bias<BN>(my_bias, p.bias, t.n_tile * BN, p.Cout, lane);
        bias_tile = t.n_tile;
      }
      float bias_nx[(BN + 31) / 32];
      const int next_tile = tile + gridDim.x;
      const int next_n = next_tile < total_tiles ? next_tile % p.n_tiles : bias_tile;
      if (p.bias && next_n != bias_tile) bias_prefetch<BN>(bias_nx, p.bias, next_n * BN, p.Cout, lane);
      const int x = t.tx * tw + xl, y = t.ty * th + yl, b = t.tbi * tb + bl;
      if (p.gn_sums && t.tbi != gn_b) {
        if (gn_b >= 0) gn_flush(my_gn, gn_n, gn_slot(p.gn_sums, gn_b, gn_n, blockIdx.x * 4 + quarter, p.gn_slots), lane);
        gn_b = t.tbi;
      }
      const bool valid = (x < p.W) && (y < p.H) && (b < p.B);
      EpiArgs e;
      e.bias = p.bias;
      e.out = p.out;
      e.o_off = b * p.o_sb + y * p.o_sy + x * p.o_sx;
      e.res_row = p.res ? res_row_ptr(p.res, b * p.r_sb + y * p.r_sy + x * p.r_sx, p.res_f32) : nullptr;
      e.res_f32 = p.res_f32;
      const int gate_b = p.gate ? (p.gate_rows > 0 ? x / p.gate_rows : b) : 0;
      e.gate_row = p.gate ? p.gate + static_cast<long long>(gate_b) * p.gate_ld : nullptr;
      e.Cout = p.Cout; e.cout_out = cout_out; e.act = p.act; e.act_c0 = p.act_c0; e.act_c1 = p.act_c1;
      e.out_f32 = p.out_f32; e.vec32 = p.out_vec32; e.res_vec32 = p.res_vec32;
      float gn_part[16];
      e.gn_part = p.gn_sums ? gn_part : nullptr;

      uint4 rr[4] = {};
      if (!p.res_f32) {
        prefetch_res_row(e, t.n_tile * out_bn, out_bn, valid);
        load_res32(rr, e, t.n_tile * out_bn, valid);
      }
      mbar_wait_relaxed(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * kAccStride + (static_cast<uint32_t>(quarter * 32) << 16);
      const float* bias_s = p.bias ? my_bias : nullptr;
      if (p.rope.cos_t) {
        if constexpr (BN % 128 == 0) epilogue_row_rope<BN>(p, epi_mode, t_row, t.n_tile, x, valid, e, bias_s, rr, lane);
      } else if (geglu)
        epilogue_row_m<true>(epi_mode, t_row, out_bn, BN / 2, t.n_tile * BN, t.n_tile * out_bn, valid, e, nullptr, 0,
                             false, bias_s, rr, my_gn, p.gn_cpg, lane);
      else
        epilogue_row_m<false>(epi_mode, t_row, out_bn, 0, t.n_tile * BN, t.n_tile * out_bn, valid, e, nullptr, 0,
                              false, bias_s, rr, my_gn, p.gn_cpg, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (p.bias && next_n != bias_tile) {
        bias_commit<BN>(my_bias, bias_nx, lane);
        bias_tile = next_n;
      }
    }
    if (p.gn_sums && gn_b >= 0) gn_flush(my_gn, gn_n, gn_slot(p.gn_sums, gn_b, gn_n, blockIdx.x * 4 + quarter, p.gn_slots), lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs computes a 256 x 256 output tile with ONE tcgen05.mma per
// K-step (M = 256, N = 256). Each CTA loads its own 128-pixel A tile and HALF (128 rows) of the weight tile, holds its
// own 128 x 256 fp32 accumulator in TMEM and runs its own epilogue. Versus the single-CTA kernel the per-CTA
// L2->SMEM stream per k-block drops from 48 KB to 32 KB and the shared-memory operand reads per MMA from 12 KB to 8 KB.
// Only the leader CTA (rank 0) issues MMAs; TMA bytes of both CTAs are credited to the leader's full barriers; the
// MMA completion is multicast to the empty / accumulator-full barriers of both CTAs.
// ------------------------------------------------------------------------------------------------
template <int BN>
struct Cfg2 {
  static constexpr int kABytes = kBM * kBK * 2;           // 16 KB: this CTA's 128 pixels
  static constexpr int kBBytes = (BN / 2) * kBK * 2;      // this CTA's half of the BN-row weight tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN >= 256) ? 6 : 7;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + 1024 + 16 * BN;
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const ConvGemmParams p) {
  using C = Cfg2<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // offset, not a cast: keeps the shared address space (LDS / STS)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;                        // [kStages]  (the leader's are the live ones)
  uint64_t* empty_bar = bars + C::kStages;          // [kStages]  per CTA
  uint64_t* acc_full = bars + 2 * C::kStages;       // [2]        per CTA
  uint64_t* acc_empty = bars + 2 * C::kStages + 2;  // [2]        leader's: 4 epilogue warps x 2 CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 4);
  float* gn_acc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform to the compiler (see the MMA issuer)
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full_bar[i], 2);   // one arrival per CTA's producer (the leader's carries the expected bytes)
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_launch_dependents();
  pdl_wait();  // nothing above touches global memory

  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  const int m_pairs = (m_tiles + 1) >> 1;
  const int total_pairs = m_pairs * p.n_tiles;
  const int num_kb = p.R * p.S * p.kc_per_tap;
  const int tw = 1 << p.log_tw, th = 1 << p.log_th;
  const int tb = kBM >> (p.log_tw + p.log_th);
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
        const int n_tile = pt % p.n_tiles;
        const int m = 2 * (pt / p.n_tiles) + static_cast<int>(rank);
        const bool dummy = m >= m_tiles;  // odd tile count: the pair's second half loads zeros and stores nothing
        const int tx = m % p.tiles_x, ty = (m / p.tiles_x) % p.tiles_y, tbi = m / (p.tiles_x * p.tiles_y);
        const int x0 = tx * tw, y0 = ty * th, b0 = dummy ? p.B : tbi * tb;
        int kb = 0;
        for (int r = 0; r < p.R; ++r) {
          for (int s = 0; s < p.S; ++s) {
            int c1, c2, c3, c4, cbase;
            if (!p.stride2) {
              cbase = 0; c1 = x0 + s - p.pad_l; c2 = y0 + r - p.pad_t; c3 = b0; c4 = 0;
            } else {
              const int dy = r - p.pad_t, dx = s - p.pad_l;
              const int py = dy & 1, px = dx & 1;
              cbase = px * p.pair_off; c1 = x0 + ((dx - px) >> 1); c2 = py; c3 = y0 + ((dy - py) >> 1); c4 = b0;
            }
            for (int kc = 0; kc < p.kc_per_tap; ++kc, ++kb) {
              mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = smem + stage * C::kStageBytes;
              uint8_t* sb = sa + C::kABytes;
              if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (p.a_bytes + C::kBBytes));
              else mbar_arrive_remote(&full_bar[stage], 0);
              tma2_load_5d(sa, &tmA, &full_bar[stage], cbase + kc * kBK, c1, c2, c3, c4);
              tma2_load_4d(sb, &tmB, &full_bar[stage], kb * kBK, n_tile * BN + static_cast<int>(rank) * (BN / 2), 0, 0);
              if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // The WHOLE warp runs this loop on warp-uniform values and one elected lane issues the MMAs and commits. Round 1
      // issued from a divergent `lane == 0` branch: every operand was then thread-private to the compiler and each
      // tcgen05.mma sat behind an ELECT / 7 x R2UR.BROADCAST / BRA.U.ANY sequence (~20 dependent instructions, 130-170
      // cycles per MMA where the tensor pipe needs 128 for N = 256 and 64 for N = 128: profiles/r02/mma_probe6_lean.txt).
      const bool leader = elect_one();
      const uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
      const uint64_t desc0 = umma_desc_sw128(0, 16, 1024);  // only the start-address field of the low word changes
      const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32), desc_lo = static_cast<uint32_t>(desc0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t a_lo = desc_lo | (sa >> 4), b_lo = desc_lo | ((sa + C::kABytes) >> 4);
          if (leader) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              tc_mma2_ss(d_tmem, (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2 * k),
                         (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2 * k), idesc, (kb | k) != 0);
            tc_commit2_mc(&empty_bar[stage]);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (leader) tc_commit2_mc(&acc_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int xl = row & (tw - 1);
    const int yl = (row >> p.log_tw) & (th - 1);
    const int bl = row >> (p.log_tw + p.log_th);
    const bool geglu = (p.act == OMGSR_ACT_GEGLU);
    const int out_bn = geglu ? BN / 2 : BN;
    int acc = 0;
    uint32_t acc_phase = 0;
    float* my_gn = gn_acc + quarter * 64;
    float* my_bias = gn_acc + 256 + quarter * BN;
    int bias_tile = -1;
    const int epi_mode = epilogue_mode(p.act, p.gate != nullptr, p.res != nullptr, p.out_f32 != 0, p.gn_sums != nullptr,
                                       p.res_f32 != 0);
    const int gn_n = p.gn_groups * 2;
    int gn_b = -1;
    if (p.gn_sums) { my_gn[lane] = 0.f; my_gn[lane + 32] = 0.f; __syncwarp(); }
    for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
      const int n_tile = pt % p.n_tiles;
      if (p.bias && n_tile != bias_tile) {  // first tile only: later ones are prefetched below
        stage_bias<BN>(my_bias, p.bias, n_tile * BN, p.Cout, lane);
        bias_tile = n_tile;
      }
      float bias_nx[(BN + 31) / 32];
      const int next_pt = pt + num_clusters;
      const int next_n = next_pt < total_pairs ? next_pt % p.n_tiles : bias_tile;
      if (p.bias && next_n != bias_tile) bias_prefetch<BN>(bias_nx, p.bias, next_n * BN, p.Cout, lane);
      const int m = 2 * (pt / p.n_tiles) + static_cast<int>(rank);
      const bool dummy = m >= m_tiles;
      const int tx = m % p.tiles_x, ty = (m / p.tiles_x) % p.tiles_y, tbi = m / (p.tiles_x * p.tiles_y);
      const int x = tx * tw + xl, y = ty * th + yl, b = tbi * tb + bl;
      if (p.gn_sums && !dummy && tbi != gn_b) {
        if (gn_b >= 0) gn_flush(my_gn, gn_n, gn_slot(p.gn_sums, gn_b, gn_n, blockIdx.x * 4 + quarter, p.gn_slots), lane);
        gn_b = tbi;
      }
      const bool valid = !dummy && (x < p.W) && (y < p.H) && (b < p.B);
      EpiArgs e;
      e.bias = p.bias;
      e.out = p.out;
      e.o_off = b * p.o_sb + y * p.o_sy + x * p.o_sx;
      e.res_row = (p.res && valid) ? res_row_ptr(p.res, b * p.r_sb + y * p.r_sy + x * p.r_sx, p.res_f32) : nullptr;
      e.res_f32 = p.res_f32;
      const int gate_b = p.gate ? (p.gate_rows > 0 ? x / p.gate_rows : min(b, p.B - 1)) : 0;
      e.gate_row = p.gate ? p.gate + static_cast<long long>(gate_b) * p.gate_ld : nullptr;
      e.Cout = p.Cout; e.cout_out = geglu ? p.Cout / 2 : p.Cout; e.act = p.act; e.act_c0 = p.act_c0; e.act_c1 = p.act_c1;
      e.out_f32 = p.out_f32; e.vec32 = p.out_vec32; e.res_vec32 = p.res_vec32;
      float gn_part[16];
      e.gn_part = p.gn_sums ? gn_part : nullptr;

      uint4 rr[4] = {};
      if (!p.res_f32) {
        prefetch_res_row(e, n_tile * out_bn, out_bn, valid);
        load_res32(rr, e, n_tile * out_bn, valid);
      }
      mbar_wait_relaxed(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * kAccStride + (static_cast<uint32_t>(quarter * 32) << 16);
      const float* bias_s = p.bias ? my_bias : nullptr;
      if (p.rope.cos_t) {
        if constexpr (BN % 128 == 0) epilogue_row_rope<BN>(p, epi_mode, t_row, n_tile, x, valid, e, bias_s, rr, lane);
      } else if (geglu)
        epilogue_row_m<true>(epi_mode, t_row, out_bn, BN / 2, n_tile * BN, n_tile * out_bn, valid, e, nullptr, 0,
                             false, bias_s, rr, my_gn, p.gn_cpg, lane);
      else
        epilogue_row_m<false>(epi_mode, t_row, out_bn, 0, n_tile * BN, n_tile * out_bn, valid, e, nullptr, 0,
                              false, bias_s, rr, my_gn, p.gn_cpg, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&acc_empty[acc]);
        else mbar_arrive_remote(&acc_empty[acc], 0);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (p.bias && next_n != bias_tile) {
        bias_commit<BN>(my_bias, bias_nx, lane);
        bias_tile = next_n;
      }
    }
    if (p.gn_sums && gn_b >= 0) gn_flush(my_gn, gn_n, gn_slot(p.gn_sums, gn_b, gn_n, blockIdx.x * 4 + quarter, p.gn_slots), lane);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// measured on B200 (profiles/r01/gemm_1cta_vs_2cta.txt): +7..19 % over the single-CTA kernel on every 256-wide shape
static int g_use_2cta = 1;
extern const char* g_trace_variant;

template <int BN>
static int launch_cfg2(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvGemmParams& p, cudaStream_t st) {
  static unsigned long long attr_done = 0;  // one bit per device
  if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(&conv_gemm_tc2_kernel<BN>), Cfg2<BN>::kSmemBytes, &attr_done, "conv_gemm_tc2")) return rc;
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  const long long pairs = static_cast<long long>((m_tiles + 1) / 2) * p.n_tiles;
  const int max_clusters = sm_count() / 2;
  const int clusters = pairs < max_clusters ? static_cast<int>(pairs) : max_clusters;
  g_trace_variant = BN == 256 ? "conv_gemm_tc2_kernel<256>" : BN == 160 ? "conv_gemm_tc2_kernel<160>" : "conv_gemm_tc2_kernel<128>";
  launch_pdl(conv_gemm_tc2_kernel<BN>, dim3(2 * clusters), dim3(kThreads), Cfg2<BN>::kSmemBytes, st, tmA, tmB, p);
  return check_launch("conv_gemm_tc2_kernel");
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

template <int BN>
static int launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvGemmParams& p, cudaStream_t st) {
  using C = Cfg<BN>;
  static unsigned long long attr_done = 0;  // one bit per device
  if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(&conv_gemm_tc_kernel<BN>), C::kSmemBytes, &attr_done, "conv_gemm")) return rc;
  const int total = p.tiles_x * p.tiles_y * p.tiles_b * p.n_tiles;
  const int grid = total < sm_count() ? total : sm_count();
  g_trace_variant = BN == 256 ? "conv_gemm_tc_kernel<256>" : BN == 160 ? "conv_gemm_tc_kernel<160>" : BN == 128 ? "conv_gemm_tc_kernel<128>" : "conv_gemm_tc_kernel<64>";
  launch_pdl(conv_gemm_tc_kernel<BN>, dim3(grid), dim3(kThreads), C::kSmemBytes, st, tmA, tmB, p);
  return check_launch("conv_gemm_tc_kernel");
}

static int pick_bn(int Cout, long long m_tiles, int forced, bool geglu, bool mn_major, bool pair_ok) {
  if (forced > 0) return forced;
  if (mn_major) return (Cout % 128 == 0) ? 128 : 64;
  // Cost model per candidate width (a divisor of Cout): waves x max(MMA cycles, operand bytes / L2->SM port rate) per
  // 64-deep k-block. One CTA streams 16 KB of A plus its share of B (the CTA pair splits B) at ~64 B/clk; the MMAs of
  // a k-block take 2 * bn cycles. Many-wave problems come out at the widest tile (least operand traffic per output
  // column); skinny ones (UNet 8x8 / 16x16 levels: M = 1024..4096) now get 128/160-wide pair tiles on fewer SMs
  // instead of 64-wide tiles on all of them, which were L2-port bound (1280->1280 3x3 @8x8: 130 us at BN = 64).
  const int cands[4] = {256, 160, 128, 64};
  int best = 0;
  double best_cost = 0.0;
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    if (Cout % bn) continue;
    if (geglu && (bn % 64)) continue;
    const bool pair = pair_ok && bn >= 128 && m_tiles >= 2;
    const long long n_tiles = Cout / bn;
    const long long ctas = pair ? 2 * ((m_tiles + 1) / 2) * n_tiles : m_tiles * n_tiles;
    const long long waves = (ctas + sm_count() - 1) / sm_count();
    const double bytes = 16384.0 + (pair ? 64.0 : 128.0) * bn;
    const double mma = 2.0 * bn, port = bytes / 64.0;
    const double cost = static_cast<double>(waves) * (mma > port ? mma : port);
    if (!best || cost < best_cost) { best = bn; best_cost = cost; }
  }
  if (best) return best;
  return 64;  // ragged Cout: 64-wide tiles, last one masked
}

struct ConvDesc {
  const void* x; const void* w; const float* bias; const void* res; const float* gate; void* y;
  int B, H, W;   // GEMM-side pixel space (== output pixels enumerated by the kernel)
  int IH, IW;    // input tensor spatial extents
  int Cin, Cout, R, S, stride2, pad_t, pad_l;
  long long x_sx, x_sy, x_sb;  // input element strides per x / y / b (stride-1 path)
  long long o_sx, o_sy, o_sb, r_sx, r_sy, r_sb;
  int act, gate_rows, force_bn, out_f32, res_f32;
  long long gate_ld; int act_c0, act_c1;
  float* gn_sums; int gn_groups;
  const float* gn_in;  // optional [B][Cin][2] input GroupNorm coefficients (+SiLU), halo pair kernel only
  RopeArgs rope;       // rope.cos_t != null: RMSNorm + RoPE on the q / k head columns (plain GEMMs only)
  long long w_ld;  // elements per weight row (R*S*Cin)
  // batched B operand (attention)
  int b_batched, b_mn_major, flat;
  int halo_mode;  // 0: generic kernel only; 1|2: try the halo kernel (descriptor base-offset mode 0|1)
  long long b_row, b_head, b_batch;
  int b_rows;  // rows of the B matrix (N for K-major, K for MN-major)
};

int try_conv_halo(const void* x, const void* w, const float* bias, const void* res, void* y, int B, int IH, int IW,
                  int OH, int OW, int Cin, int Cout, int R, int S, int pad_t, int pad_l, long long x_sx, long long x_sy,
                  long long x_sb, long long o_sx, long long o_sy, long long o_sb, long long r_sx, long long r_sy,
                  long long r_sb, long long w_ld, int act, int out_f32, float* gn_sums, int gn_groups, const float* gn_in,
                  cudaStream_t st);

static int run_conv_gemm_impl(const ConvDesc& d, cudaStream_t st) {
  // K need not fill the last 64-deep k-block of a plain GEMM (R = S = 1, K-major B): the tensor maps carry the true
  // K as their inner extent and TMA zero-fills the rest of the box for BOTH operands (used by the K = 32 im2col form
  // of the 3-channel conv_in). With taps, a partial block would run into the next tap's weights.
  const bool ragged_k = (d.Cin % kBK) != 0;
  if (ragged_k && (d.R * d.S != 1 || d.b_mn_major || d.stride2 || (d.Cin % 8)))
    return set_error(OMGSR_ERR_SHAPE, "conv_gemm: K/Cin=%d must be a multiple of %d (or of 8 for a plain GEMM)", d.Cin, kBK);
  if ((reinterpret_cast<uintptr_t>(d.x) | reinterpret_cast<uintptr_t>(d.w) | reinterpret_cast<uintptr_t>(d.y) |
       reinterpret_cast<uintptr_t>(d.res)) & 15)
    return set_error(OMGSR_ERR_ALIGN, "conv_gemm: pointers must be 16-byte aligned");
  if ((d.x_sx | d.x_sy | d.x_sb | d.o_sx | d.o_sy | d.o_sb | d.r_sx | d.r_sy | d.r_sb | d.w_ld) & 7)
    return set_error(OMGSR_ERR_ALIGN, "conv_gemm: strides must be multiples of 8 elements");
  const bool geglu = d.act == OMGSR_ACT_GEGLU;
  if (d.halo_mode && !d.stride2 && !d.b_batched && !d.flat && !d.gate && (d.force_bn == 0)) {
    // wide stride-1 3x3 / 2x2 layers: input-halo kernel (conv_halo_tc.cu); returns 1 when not eligible
    const int rc = try_conv_halo(d.x, d.w, d.bias, d.res, d.y, d.B, d.IH, d.IW, d.H, d.W, d.Cin, d.Cout, d.R, d.S,
                                 d.pad_t, d.pad_l, d.x_sx, d.x_sy, d.x_sb, d.o_sx, d.o_sy, d.o_sb, d.r_sx, d.r_sy,
                                 d.r_sb, d.w_ld, d.act, d.out_f32, d.gn_sums, d.gn_groups, d.gn_in, st);
    if (rc != 1) return rc;
  }
  if (d.gn_in)
    return set_error(OMGSR_ERR_SHAPE, "conv2d: the fused input GroupNorm is only available on the input-halo path "
                                      "(3x3 stride 1, W >= 96, Cin %% 64 == 0, 16 < Cout <= 128)");

  // tile shape {tw, th, tb} (powers of two, product 128): minimise the padded pixel count, prefer wide rows
  int tw = 128, th = 1, tb = 1;
  if (!d.flat) {
    long long best = -1;
    for (int c = 128; c >= 1; c >>= 1) {
      if (c > d.W && c > 1) continue;
      int h = 128 / c;
      while (h > 1 && h > d.H) h >>= 1;
      const int nb = 128 / (c * h);
      if (d.gn_sums && nb != 1) continue;  // fused GroupNorm statistics: a tile must not straddle two images
      const long long cost = static_cast<long long>((d.W + c - 1) / c) * c * ((d.H + h - 1) / h) * h *
                             ((d.B + nb - 1) / nb) * nb;
      if (best < 0 || cost < best) { best = cost; tw = c; th = h; tb = nb; }
    }
  }
  // never ask TMA for a box larger than the tensor along the batch axis; the unused accumulator rows
  // are masked in the epilogue
  const int tb_box = tb > d.B ? d.B : tb;

  ConvGemmParams p;
  p.log_tw = ilog2(tw); p.log_th = ilog2(th);
  p.tiles_x = (d.W + tw - 1) / tw; p.tiles_y = (d.H + th - 1) / th; p.tiles_b = (d.B + tb - 1) / tb;
  p.W = d.W; p.H = d.H; p.B = d.B;
  p.Cout = d.Cout; p.kc_per_tap = (d.Cin + kBK - 1) / kBK; p.R = d.R; p.S = d.S; p.pad_t = d.pad_t; p.pad_l = d.pad_l;
  p.stride2 = d.stride2; p.pair_off = static_cast<int>(d.x_sx);
  p.b_batched = d.b_batched; p.b_mn_major = d.b_mn_major;
  p.o_sx = d.o_sx; p.o_sy = d.o_sy; p.o_sb = d.o_sb; p.r_sx = d.r_sx; p.r_sy = d.r_sy; p.r_sb = d.r_sb;
  p.out = d.y; p.bias = d.bias; p.res = static_cast<const bf16*>(d.res);
  p.gate = d.gate; p.gate_rows = d.gate_rows; p.act = d.act; p.out_f32 = d.out_f32;
  p.out_vec32 = !d.out_f32 && (reinterpret_cast<uintptr_t>(d.y) % 32 == 0) && ((d.o_sx | d.o_sy | d.o_sb) % 16 == 0);
  p.res_vec32 = d.res && (reinterpret_cast<uintptr_t>(d.res) % 32 == 0) && ((d.r_sx | d.r_sy | d.r_sb) % 16 == 0);
  p.res_f32 = 0;
  if (d.res_f32) {
    // fp32 residual stream: out = res + gate * (acc + bias) with fp32 res / out rows (epilogue_row_f32stream)
    if (!d.res || !d.out_f32 || d.act != OMGSR_ACT_NONE || d.gn_sums || (d.Cout % 32) ||
        (reinterpret_cast<uintptr_t>(d.res) % 32) || (reinterpret_cast<uintptr_t>(d.y) % 32))
      return set_error(OMGSR_ERR_SHAPE, "conv_gemm: OMGSR_RES_F32 needs OMGSR_OUT_F32, a residual, no activation, "
                                        "N %% 32 == 0 and 32-byte aligned rows (N=%d)", d.Cout);
    p.res_f32 = 1;
  }
  p.gate_ld = d.gate_ld > 0 ? d.gate_ld : (geglu ? d.Cout / 2 : d.Cout);
  p.act_c0 = d.act_c1 > d.act_c0 ? d.act_c0 : 0; p.act_c1 = d.act_c1 > d.act_c0 ? d.act_c1 : 0x7fffffff;
  p.a_bytes = kBK * 2 * tw * th * tb_box;
  p.rope = d.rope;
  if (d.rope.cos_t) {
    if (!d.flat || d.out_f32 || d.res || d.gate || d.gn_sums || geglu || d.b_batched || (d.Cout % 128) ||
        ((d.rope.q0 | d.rope.q1 | d.rope.k0 | d.rope.k1) % 128) || !d.rope.wq || !d.rope.wk || !d.rope.sin_t)
      return set_error(OMGSR_ERR_SHAPE, "gemm_rope: plain bf16 GEMM with N %% 128 == 0 and head-aligned q / k ranges only");
    if ((reinterpret_cast<uintptr_t>(d.y) % 32) || (d.o_sx % 16) || (d.o_sb % 16))
      return set_error(OMGSR_ERR_ALIGN, "gemm_rope: output rows must be 32-byte aligned");
  }
  p.gn_sums = nullptr; p.gn_cpg = 0; p.gn_groups = 0;
  if (d.gn_sums) {
    // fused statistics need: bf16 output, one image per tile, 8-channel store granules that map to whole groups
    const int cpg = d.gn_groups > 0 ? d.Cout / d.gn_groups : 0;
    if (d.out_f32 || geglu || tb != 1 || d.gn_groups <= 0 || d.gn_groups > 32 || d.Cout % d.gn_groups || !(cpg == 4 || cpg == 8 || cpg == 16) || (d.Cout % 32))
      return set_error(OMGSR_ERR_SHAPE, "conv_gemm: fused GroupNorm statistics unsupported for this shape (Cout=%d groups=%d tb=%d)",
                       d.Cout, d.gn_groups, tb);
    p.gn_sums = d.gn_sums; p.gn_cpg = cpg; p.gn_groups = d.gn_groups; p.gn_slots = gn_slots_per_image();
  }

  const long long m_tiles = static_cast<long long>(p.tiles_x) * p.tiles_y * p.tiles_b;
  int bn = pick_bn(d.Cout, m_tiles, d.force_bn, geglu, d.b_mn_major, g_use_2cta && !d.b_batched);
  if (d.rope.cos_t && bn != 256 && bn != 128) bn = (d.Cout % 256 == 0) ? 256 : 128;  // whole 128-column heads per tile
  if (bn != 256 && bn != 160 && bn != 128 && bn != 64) return set_error(OMGSR_ERR_SHAPE, "conv_gemm: BN=%d", bn);
  if (geglu && ((bn % 64) || (d.Cout % bn)))
    return set_error(OMGSR_ERR_SHAPE, "conv_gemm: GEGLU needs Cout %% BN == 0 and BN %% 64 == 0");
  if (d.b_mn_major && bn != 64 && bn != 128)
    return set_error(OMGSR_ERR_SHAPE, "conv_gemm: MN-major B needs BN in {64,128}");
  if ((d.Cout % 8) && d.o_sx < ((d.Cout + 7) & ~7))
    // ragged Cout: the 8-channel store granule writes up to the next multiple of 8
    return set_error(OMGSR_ERR_SHAPE, "conv_gemm: Cout=%d needs an output row stride >= %d", d.Cout,
                     (d.Cout + 7) & ~7);
  p.n_tiles = (d.Cout + bn - 1) / bn;
  if (m_tiles * p.n_tiles > 0x7fffffffLL) return set_error(OMGSR_ERR_SHAPE, "conv_gemm: too many tiles");

  // ---- tensor maps ----
  CUtensorMap tmA, tmB;
  uint64_t dims[5], strides[4];
  uint32_t box[5];
  const uint64_t eb = 2;
  if (!d.stride2) {
    dims[0] = d.Cin; dims[1] = d.IW; dims[2] = d.IH; dims[3] = d.B; dims[4] = 1;
    strides[0] = d.x_sx * eb; strides[1] = d.x_sy * eb; strides[2] = d.x_sb * eb;
    strides[3] = strides[2] * d.B;
    box[0] = kBK; box[1] = tw; box[2] = th; box[3] = tb_box; box[4] = 1;
  } else {
    if ((d.IW & 1) || (d.IH & 1)) return set_error(OMGSR_ERR_SHAPE, "conv_gemm: stride-2 needs even input extents");
    if (d.x_sy != d.x_sx * d.IW || d.x_sb != d.x_sy * d.IH)
      return set_error(OMGSR_ERR_SHAPE, "conv_gemm: stride-2 needs an input dense in x/y/b (channel stride allowed)");
    // {pixel pair (ld + Cin channels: even pixel at 0, odd pixel at ld), W/2, row parity, H/2, B}
    dims[0] = d.x_sx + d.Cin; dims[1] = d.IW / 2; dims[2] = 2; dims[3] = d.IH / 2; dims[4] = d.B;
    strides[0] = 2 * d.x_sx * eb;                            // next even x
    strides[1] = static_cast<uint64_t>(d.IW) * d.x_sx * eb;  // next row parity
    strides[2] = 2 * strides[1];                             // next even y
    strides[3] = strides[1] * d.IH;                          // next image
    box[0] = kBK; box[1] = tw; box[2] = 1; box[3] = th; box[4] = tb_box;
  }
  int rc = make_tmap_bf16(&tmA, d.x, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  uint64_t wd[4], ws[3];
  uint32_t wb[4];
  if (!d.b_batched) {
    wd[0] = d.w_ld; wd[1] = d.Cout; wd[2] = 1; wd[3] = 1;
    ws[0] = d.w_ld * eb; ws[1] = ws[0] * d.Cout; ws[2] = ws[1];
    wb[0] = kBK; wb[1] = bn; wb[2] = 1; wb[3] = 1;
    if (g_use_2cta && (bn == 256 || bn == 160 || bn == 128) && (d.Cout % bn == 0) && m_tiles >= 2) {
      // CTA-pair kernel: each CTA of the pair loads half of the bn-row weight tile
      wb[1] = bn / 2;
      rc = make_tmap_bf16(&tmB, d.w, 4, wd, ws, wb, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
      if (bn == 256) return launch_cfg2<256>(tmA, tmB, p, st);
      if (bn == 160) return launch_cfg2<160>(tmA, tmB, p, st);
      return launch_cfg2<128>(tmA, tmB, p, st);
    }
  } else if (!d.b_mn_major) {
    wd[0] = d.Cin; wd[1] = d.b_rows; wd[2] = d.H; wd[3] = d.B;
    ws[0] = d.b_row * eb; ws[1] = d.b_head * eb; ws[2] = d.b_batch * eb;
    wb[0] = kBK; wb[1] = bn; wb[2] = 1; wb[3] = 1;
  } else {
    wd[0] = d.Cout; wd[1] = d.b_rows; wd[2] = d.H; wd[3] = d.B;
    ws[0] = d.b_row * eb; ws[1] = d.b_head * eb; ws[2] = d.b_batch * eb;
    wb[0] = 64; wb[1] = kBK; wb[2] = 1; wb[3] = 1;
  }
  rc = make_tmap_bf16(&tmB, d.w, 4, wd, ws, wb, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;

  switch (bn) {
    case 256: return launch_cfg<256>(tmA, tmB, p, st);
    case 160: return launch_cfg<160>(tmA, tmB, p, st);
    case 128: return launch_cfg<128>(tmA, tmB, p, st);
    default: return launch_cfg<64>(tmA, tmB, p, st);
  }
}

// OMGSR_TRACE=1: time every GEMM / convolution launch with CUDA events (serialising) and print its shape, the
// kernel variant that served it and the achieved TFLOP/s to stderr - the per-layer table under profiles/.
const char* g_trace_variant = "";
static int run_conv_gemm(const ConvDesc& d, cudaStream_t st) {
  static const bool trace = getenv("OMGSR_TRACE") != nullptr;
  if (!trace) return run_conv_gemm_impl(d, st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  const int rc = run_conv_gemm_impl(d, st);
  cudaEventRecord(e1, st);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  const double flops = 2.0 * d.B * d.H * d.W * d.Cout * d.Cin * d.R * d.S;
  fprintf(stderr, "TRACE %s B=%d H=%d W=%d Cin=%d Cout=%d k=%dx%d s2=%d flat=%d batched=%d act=%d res=%d gn=%d %.4f ms %.0f TFLOP/s\n",
          g_trace_variant, d.B, d.H, d.W, d.Cin, d.Cout, d.R, d.S, d.stride2, d.flat, d.b_batched, d.act,
          d.res ? 1 : 0, d.gn_sums ? 1 : 0, ms, flops / ms * 1e-9);
  return rc;
}

}  // namespace omgsr

using namespace omgsr;

/* 1: route 256-wide weight-stationary GEMMs / convs to the CTA-pair (cta_group::2) kernel */
namespace omgsr { extern int g_halo_2cta; }
extern "C" int omgsr_set_gemm_2cta(int on) { omgsr::g_use_2cta = on; omgsr::g_halo_2cta = on; return OMGSR_OK; }

extern "C" int omgsr_gemm_bf16(const void* A, const void* Wt, const float* bias, const void* res, const float* gate,
                               void* D, long long M, int N, int K, long long lda, long long ldd, long long ldr,
                               int act, int gate_rows, int flags, int force_bn, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(OMGSR_ERR_SHAPE, "gemm: empty problem");
  if (M > 0x7fffffffLL) return set_error(OMGSR_ERR_SHAPE, "gemm: M too large");
  if (gate && gate_rows <= 0) return set_error(OMGSR_ERR_SHAPE, "gemm: gate needs gate_rows > 0");
  ConvDesc d{};
  d.x = A; d.w = Wt; d.bias = bias; d.res = res; d.gate = gate; d.y = D;
  d.B = 1; d.H = 1; d.W = static_cast<int>(M); d.IH = 1; d.IW = static_cast<int>(M);
  d.Cin = K; d.Cout = N; d.R = 1; d.S = 1;
  d.x_sx = lda; d.x_sy = lda * M; d.x_sb = lda * M;
  d.o_sx = ldd; d.o_sy = 0; d.o_sb = 0; d.r_sx = ldr; d.r_sy = 0; d.r_sb = 0;
  d.act = act; d.gate_rows = gate ? gate_rows : 0; d.force_bn = force_bn; d.w_ld = K;
  d.out_f32 = (flags & OMGSR_OUT_F32) ? 1 : 0;
  d.res_f32 = (flags & OMGSR_RES_F32) ? 1 : 0;
  d.flat = 1;
  return run_conv_gemm(d, static_cast<cudaStream_t>(stream));
}

extern "C" int omgsr_gemm_batched_bf16(const void* A, const void* Wt, const float* bias, const void* res,
                                       const float* gate, void* D, int batch, int M, int N, int K, long long lda,
                                       long long a_batch, long long ldd, long long d_batch, long long ldr,
                                       long long r_batch, long long gate_ld, int act, int act_c0, int act_c1,
                                       int flags, int force_bn, void* stream) {
  if (batch <= 0 || M <= 0 || N <= 0 || K <= 0) return set_error(OMGSR_ERR_SHAPE, "gemm_batched: empty problem");
  ConvDesc d{};
  d.x = A; d.w = Wt; d.bias = bias; d.res = res; d.gate = gate; d.y = D;
  d.B = batch; d.H = 1; d.W = M; d.IH = 1; d.IW = M;
  d.Cin = K; d.Cout = N; d.R = 1; d.S = 1;
  d.x_sx = lda; d.x_sy = lda * M; d.x_sb = a_batch;
  d.o_sx = ldd; d.o_sy = 0; d.o_sb = d_batch; d.r_sx = ldr; d.r_sy = 0; d.r_sb = r_batch;
  d.act = act; d.gate_rows = 0; d.force_bn = force_bn; d.w_ld = K; d.gate_ld = gate_ld;
  d.act_c0 = act_c0; d.act_c1 = act_c1;
  d.out_f32 = (flags & OMGSR_OUT_F32) ? 1 : 0;
  d.res_f32 = (flags & OMGSR_RES_F32) ? 1 : 0;
  d.flat = 1;  // 128-row tiles never straddle two batch entries
  return run_conv_gemm(d, static_cast<cudaStream_t>(stream));
}

extern "C" int omgsr_gemm_rope_bf16(const void* A, const void* Wt, const float* bias, void* D, int batch, int M, int N,
                                    int K, long long lda, long long a_batch, long long ldd, long long d_batch, int act,
                                    int act_c0, int act_c1, const float* wq, const float* wk, int q0, int q1, int k0,
                                    int k1, const float* cos_t, const float* sin_t, int pos_offset, float eps,
                                    int force_bn, void* stream) {
  if (batch <= 0 || M <= 0 || N <= 0 || K <= 0) return set_error(OMGSR_ERR_SHAPE, "gemm_rope: empty problem");
  if (!cos_t || !sin_t || !wq || !wk) return set_error(OMGSR_ERR_SHAPE, "gemm_rope: tables and norm weights are required");
  if (q0 < 0 || q1 > N || k0 < 0 || k1 > N || q1 < q0 || k1 < k0) return set_error(OMGSR_ERR_SHAPE, "gemm_rope: column ranges");
  ConvDesc d{};
  d.x = A; d.w = Wt; d.bias = bias; d.y = D;
  d.B = batch; d.H = 1; d.W = M; d.IH = 1; d.IW = M;
  d.Cin = K; d.Cout = N; d.R = 1; d.S = 1;
  d.x_sx = lda; d.x_sy = lda * M; d.x_sb = a_batch;
  d.o_sx = ldd; d.o_sy = 0; d.o_sb = d_batch;
  d.act = act; d.force_bn = force_bn; d.w_ld = K; d.act_c0 = act_c0; d.act_c1 = act_c1;
  d.flat = 1;
  d.rope.wq = wq; d.rope.wk = wk; d.rope.cos_t = cos_t; d.rope.sin_t = sin_t;
  d.rope.q0 = q0; d.rope.q1 = q1; d.rope.k0 = k0; d.rope.k1 = k1; d.rope.pos_offset = pos_offset; d.rope.eps = eps;
  return run_conv_gemm(d, static_cast<cudaStream_t>(stream));
}

extern "C" int omgsr_bgemm_bf16(const void* A, const void* Bm, void* D, int batch, int heads, int M, int N, int K,
                                long long a_row, long long a_head, long long a_batch, long long b_row,
                                long long b_head, long long b_batch, long long d_row, long long d_head,
                                long long d_batch, int b_valid_rows, int flags, int force_bn, void* stream) {
  if (batch <= 0 || heads <= 0 || M <= 0 || N <= 0 || K <= 0)
    return set_error(OMGSR_ERR_SHAPE, "bgemm: empty problem");
  if ((b_row | b_head | b_batch) & 7) return set_error(OMGSR_ERR_ALIGN, "bgemm: B strides must be multiples of 8");
  ConvDesc d{};
  d.x = A; d.w = Bm; d.y = D;
  d.B = batch; d.H = heads; d.W = M; d.IH = heads; d.IW = M;
  d.Cin = K; d.Cout = N; d.R = 1; d.S = 1;
  d.x_sx = a_row; d.x_sy = a_head; d.x_sb = a_batch;
  d.o_sx = d_row; d.o_sy = d_head; d.o_sb = d_batch;
  d.force_bn = force_bn; d.w_ld = 8;
  d.out_f32 = (flags & OMGSR_OUT_F32) ? 1 : 0;
  d.flat = 1; d.b_batched = 1; d.b_mn_major = (flags & OMGSR_B_MN_MAJOR) ? 1 : 0;
  d.b_row = b_row; d.b_head = b_head; d.b_batch = b_batch;
  // rows of Bm that really exist (K may be padded up to the operand granule; the rest must read as zero)
  d.b_rows = b_valid_rows > 0 ? b_valid_rows : (d.b_mn_major ? K : N);
  return run_conv_gemm(d, static_cast<cudaStream_t>(stream));
}

static int conv2d_impl(const void* x, const void* w, const float* bias, const void* res, void* y, int B, int H, int W,
                       int Cin, int Cout, int ksize, int stride, int pad_t, int pad_l, int pad_b, int pad_r,
                       int upsample2x, int act, long long ldx, long long ldy, long long ldr, float* gn_sums,
                       int gn_groups, float* gn_scratch, long long gn_scratch_floats, const float* gn_in, int flags,
                       int force_bn, void* stream);

extern "C" int omgsr_conv2d_nhwc_bf16(const void* x, const void* w, const float* bias, const void* res, void* y,
                                      int B, int H, int W, int Cin, int Cout, int ksize, int stride, int pad_t,
                                      int pad_l, int pad_b, int pad_r, int upsample2x, int act, long long ldx,
                                      long long ldy, long long ldr, float* gn_sums, int gn_groups, float* gn_scratch,
                                      long long gn_scratch_floats, int flags, int force_bn, void* stream) {
  return conv2d_impl(x, w, bias, res, y, B, H, W, Cin, Cout, ksize, stride, pad_t, pad_l, pad_b, pad_r, upsample2x, act,
                     ldx, ldy, ldr, gn_sums, gn_groups, gn_scratch, gn_scratch_floats, nullptr, flags, force_bn, stream);
}

extern "C" int omgsr_conv2d_gn_nhwc_bf16(const void* x, const float* gn_in_coef, const void* w, const float* bias,
                                         const void* res, void* y, int B, int H, int W, int Cin, int Cout,
                                         long long ldx, long long ldy, long long ldr, float* gn_sums, int gn_groups,
                                         float* gn_scratch, long long gn_scratch_floats, void* stream) {
  if (!gn_in_coef) return set_error(OMGSR_ERR_SHAPE, "conv2d_gn: gn_in_coef is required");
  return conv2d_impl(x, w, bias, res, y, B, H, W, Cin, Cout, 3, 1, 1, 1, 1, 1, 0, OMGSR_ACT_NONE, ldx, ldy, ldr, gn_sums,
                     gn_groups, gn_scratch, gn_scratch_floats, gn_in_coef, 0, 0, stream);
}

static int conv2d_impl(const void* x, const void* w, const float* bias, const void* res, void* y, int B, int H, int W,
                       int Cin, int Cout, int ksize, int stride, int pad_t, int pad_l, int pad_b, int pad_r,
                       int upsample2x, int act, long long ldx, long long ldy, long long ldr, float* gn_sums,
                       int gn_groups, float* gn_scratch, long long gn_scratch_floats, const float* gn_in, int flags,
                       int force_bn, void* stream) {
  if (B <= 0 || H <= 0 || W <= 0) return set_error(OMGSR_ERR_SHAPE, "conv2d: empty problem");
  if (ksize != 1 && ksize != 3) return set_error(OMGSR_ERR_SHAPE, "conv2d: ksize must be 1 or 3");
  if (stride != 1 && stride != 2) return set_error(OMGSR_ERR_SHAPE, "conv2d: stride must be 1 or 2");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ConvDesc d{};
  d.x = x; d.bias = bias; d.res = res; d.y = y;
  d.B = B; d.IH = H; d.IW = W; d.Cin = Cin; d.Cout = Cout;
  d.x_sx = ldx; d.x_sy = ldx * W; d.x_sb = ldx * W * H;
  d.act = act; d.force_bn = force_bn; d.out_f32 = (flags & OMGSR_OUT_F32) ? 1 : 0;
  d.halo_mode = (flags & OMGSR_CONV_NO_HALO) ? 0 : 1;
  d.gn_sums = nullptr; d.gn_groups = gn_groups; d.gn_in = gn_in;
  // deterministic fused statistics: the kernels write per-warp partial sums into gn_scratch [launch][image][slot][2G]
  // (zeroed here: slots of CTAs that never see an image stay zero), gn_finalize folds them into gn_sums afterwards
  const int gn_launches = upsample2x ? 4 : 1;
  const long long gn_per_launch = static_cast<long long>(B) * gn_slots_per_image() * 2 * gn_groups;
  if (gn_sums) {
    if (gn_groups <= 0) return set_error(OMGSR_ERR_SHAPE, "conv2d: gn_groups must be > 0 with gn_sums");
    if (!gn_scratch || gn_scratch_floats < gn_launches * gn_per_launch)
      return set_error(OMGSR_ERR_SHAPE, "conv2d: gn_scratch needs %lld floats (omgsr_gn_scratch_floats), got %lld",
                       gn_launches * gn_per_launch, gn_scratch ? gn_scratch_floats : 0LL);
    cudaError_t e = cudaMemsetAsync(gn_scratch, 0, sizeof(float) * gn_launches * gn_per_launch, st);
    if (e != cudaSuccess) return set_error(OMGSR_ERR_CUDA, "conv2d gn memset: %s", cudaGetErrorString(e));
    d.gn_sums = gn_scratch;
  }
  if (upsample2x) {
    // nearest-2x + 3x3/p1 == four 2x2 convolutions on the un-upsampled input (weights pre-collapsed
    // per output parity by the packer): out(2i+py, 2j+px).
    if (ksize != 3 || stride != 1 || pad_t != 1 || pad_l != 1)
      return set_error(OMGSR_ERR_SHAPE, "conv2d: upsample2x fusion needs 3x3 stride 1 pad 1");
    d.H = H; d.W = W; d.R = 2; d.S = 2; d.w_ld = 4LL * Cin;
    const long long OW = 2LL * W, OH = 2LL * H;
    d.o_sx = 2 * ldy; d.o_sy = 2 * OW * ldy; d.o_sb = OH * OW * ldy;
    d.r_sx = 2 * ldr; d.r_sy = 2 * OW * ldr; d.r_sb = OH * OW * ldr;
    const size_t esz = d.out_f32 ? 4 : 2;
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        d.w = static_cast<const bf16*>(w) + static_cast<long long>(py * 2 + px) * Cout * d.w_ld;
        d.pad_t = 1 - py; d.pad_l = 1 - px;
        const long long shift = py * OW + px;
        d.y = static_cast<char*>(y) + shift * ldy * esz;
        d.res = res ? static_cast<const bf16*>(res) + shift * ldr : nullptr;
        if (gn_sums) d.gn_sums = gn_scratch + (py * 2 + px) * gn_per_launch;
        int rc = run_conv_gemm(d, st);
        if (rc) return rc;
      }
    if (gn_sums) return launch_gn_finalize(gn_scratch, 4, B, gn_slots_per_image(), 2 * gn_groups, gn_sums, st);
    return OMGSR_OK;
  }
  const int OHh = (H + pad_t + pad_b - ksize) / stride + 1;
  const int OWw = (W + pad_l + pad_r - ksize) / stride + 1;
  if (OHh <= 0 || OWw <= 0) return set_error(OMGSR_ERR_SHAPE, "conv2d: empty output");
  d.w = w; d.H = OHh; d.W = OWw; d.R = ksize; d.S = ksize; d.stride2 = (stride == 2);
  d.pad_t = pad_t; d.pad_l = pad_l;
  d.o_sx = ldy; d.o_sy = ldy * OWw; d.o_sb = ldy * OWw * OHh;
  d.r_sx = ldr; d.r_sy = ldr * OWw; d.r_sb = ldr * OWw * OHh;
  d.w_ld = static_cast<long long>(ksize) * ksize * Cin;
  const int rc = run_conv_gemm(d, st);
  if (rc || !gn_sums) return rc;
  return launch_gn_finalize(gn_scratch, 1, B, gn_slots_per_image(), 2 * gn_groups, gn_sums, st);
}
