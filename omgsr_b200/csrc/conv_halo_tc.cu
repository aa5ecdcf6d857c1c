// 3x3 / 2x2 stride-1 convolution on tcgen05 with input-halo reuse ("halo kernel").
//
// The generic implicit-GEMM kernel (gemm_conv_tc.cu) re-fetches the input tile once per filter tap: for the wide,
// thin-channel VAE levels (C = 128/256 at 256^2..1024^2) that makes the L2->SMEM stream, not the tensor pipe, the
// limit (ncu: 19 % of peak on the 128->128 @512^2 layer). This kernel instead loads, per 64-channel slice, ONE
// halo tile {64 ch, 128+S-1 pixels, TM+R-1 rows} and forms the A operand of every (tap, output row) directly
// from it: in the 128B-swizzled K-major layout a shift by one pixel is a shift of the descriptor start address by
// one 128-byte row, so all R*S*TM operand views alias the same shared-memory tile. One CTA computes TM output
// rows x 128 pixels (TM accumulators in TMEM) so each weight k-block is also reused TM times.
//   traffic per 2x128 output pixels, C=128: 4 rows x 130 px x 256 B + 295 KB weights = 428 KB
//   (generic kernel: 2 x 18 k-blocks x 32 KB = 1152 KB).
// Warp roles as in gemm_conv_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue.
// Three instantiations: conv_halo_tc2_kernel<2,false> (CTA pair, the default for 16 < Cout <= 128),
// conv_halo_tc2_kernel<2,true> (the same plus four transform warps that apply the INPUT's GroupNorm + SiLU to the halo
// tile in shared memory, omgsr_conv2d_gn_nhwc_bf16) and the single-CTA conv_halo_tc_kernel<BN,2> (BN = 16 for thin
// Cout such as the decoder's conv_out, BN = 128 for single-tile problems).
#include <vector>

#include "common.cuh"
#include "epilogue.cuh"
#include "../../include/omgsr_b200.h"

namespace omgsr {

extern const char* g_trace_variant;

struct HaloParams {
  int tiles_x, tiles_y, B;    // tiles_y counts groups of TM rows
  int W, H;                   // output extents
  int n_tiles, Cout;
  int chunks;                 // Cin / 64
  int R, S, pad_t, pad_l, Cin;
  int pw;                     // halo pitch in pixels = 128 + S - 1
  int a_bytes;                // bytes of one halo load
  int debug;                  // experiment knob (OMGSR_HALO_DEBUG): 1 no stores, 2 no epilogue, 4 no MMA
  long long o_sx, o_sy, o_sb, r_sx, r_sy, r_sb;
  void* out;
  const float* bias;
  const bf16* res;
  int act, out_f32;
  int out_vec32, res_vec32;
  float* gn_sums;
  int gn_cpg, gn_groups, gn_slots;  // gn_sums: scratch [image][gn_slots][groups * 2], see gn_flush (epilogue.cuh)
  const float2* gn_in;        // optional [B][Cin] (scale, shift): GroupNorm + SiLU applied to the INPUT (pair kernel)
  int IH, IW;                 // input extents (validity of halo pixels for the input transform)
  int res_tap;                // rolling-row kernel: the residual is accumulated as an identity-weight tap (see there)
};

template <int BN, int TM>
struct HaloCfg {
  static constexpr int kABytes = (((TM + 2) * 130 * 128) + 1023) / 1024 * 1024;
  static constexpr int kAStages = 2;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kBStages = (BN == 128) ? 5 : (BN <= 32 ? 8 : 2);
  static constexpr int kAccSets = (BN * TM * 2 <= 512) ? 2 : 1;
  static constexpr int kBiasSlice = BN < 32 ? 32 : BN;  // floats per epilogue warp (a chunk reads 32)
  static constexpr int kSmemBytes = kAStages * kABytes + kBStages * kBBytes + 1024 + 256 + 1024 /*GN partial sums*/ +
                                    16 * kBiasSlice /*bias slices*/;
};

// NOTE on the shifted operand views: the 128B swizzle is a function of the absolute shared-memory address bits
// [4:6] ^ [7:9], so an operand view that starts at any 128-byte row of a 1024-byte aligned, TMA-written tile needs NO
// base-offset correction in its descriptor (verified on B200: setting the base-offset field to (addr >> 7) & 7 gives
// wrong results), and aligning the views to whole 8-row swizzle atoms does not change the MMA rate either.

template <int BN, int TM>
__global__ void __launch_bounds__(192, 1)
conv_halo_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const HaloParams p) {
  using C = HaloCfg<BN, TM>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // offset, not a cast: keeps the shared address space (LDS / STS)
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::kAStages * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + C::kBStages * C::kBBytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + C::kAStages;
  uint64_t* b_full = a_empty + C::kAStages;
  uint64_t* b_empty = b_full + C::kBStages;
  uint64_t* acc_full = b_empty + C::kBStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* gn_acc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [4 warps][64]

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // uniform: see the MMA issuer
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < C::kAStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < C::kBStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
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

  const int total_tiles = p.tiles_x * p.tiles_y * p.B * p.n_tiles;
  const int taps = p.R * p.S;

  if (warp == 0) {
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        int m = tile / p.n_tiles;
        const int tx = m % p.tiles_x; m /= p.tiles_x;
        const int ty = m % p.tiles_y;
        const int b = m / p.tiles_y;
        for (int c = 0; c < p.chunks; ++c) {
          mbar_wait(&a_empty[as], aph ^ 1);
          mbar_expect_tx(&a_full[as], p.a_bytes);
          tma_load_5d(sA + as * C::kABytes, &tmA, &a_full[as], c * 64, tx * 128 - p.pad_l, ty * TM - p.pad_t, b, 0);
          if (++as == C::kAStages) { as = 0; aph ^= 1; }
          for (int tap = 0; tap < taps; ++tap) {
            mbar_wait(&b_empty[bs], bph ^ 1);
            mbar_expect_tx(&b_full[bs], C::kBBytes);
            tma_load_4d(sB + bs * C::kBBytes, &tmB, &b_full[bs], tap * p.Cin + c * 64, n_tile * BN, 0, 0);
            if (++bs == C::kBStages) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    {  // whole warp on warp-uniform values, one elected lane issues (gemm_conv_tc.cu has the argument)
      const bool leader = elect_one();
      const uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      const uint64_t desc0 = umma_desc_sw128(0, 16, 1024);
      const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32), desc_lo = static_cast<uint32_t>(desc0);
      const bool no_mma = (p.debug & 4) != 0;
      int as = 0, bs = 0, set = 0;
      uint32_t aph = 0, bph = 0, sph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty[set], sph ^ 1);
        tc_fence_after();
        const uint32_t acc0 = tmem_base + set * (TM * BN);
        for (int c = 0; c < p.chunks; ++c) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + as * C::kABytes);
          int tap = 0;
          for (int r = 0; r < p.R; ++r) {
            for (int s = 0; s < p.S; ++s, ++tap) {
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              const uint32_t b_lo = desc_lo | (smem_u32(sB + bs * C::kBBytes) >> 4);
              const uint32_t a_lo = desc_lo | ((a_base + static_cast<uint32_t>((r * p.pw + s) * 128)) >> 4);
              const uint32_t a_row = static_cast<uint32_t>(p.pw * 8);  // one halo row down, in 16-byte units
              if (leader && !no_mma) {
#pragma unroll
                for (int t = 0; t < TM; ++t) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    tc_mma_ss(acc0 + t * BN, (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + t * a_row + 2 * k),
                              (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2 * k), idesc, (c | tap | k) != 0);
                }
              }
              if (leader) tc_commit(&b_empty[bs]);
              if (++bs == C::kBStages) { bs = 0; bph ^= 1; }
            }
          }
          if (leader) tc_commit(&a_empty[as]);
          if (++as == C::kAStages) { as = 0; aph ^= 1; }
        }
        if (leader) tc_commit(&acc_full[set]);
        if (C::kAccSets == 2) { if (++set == 2) { set = 0; sph ^= 1; } }
        else { sph ^= 1; }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int set = 0;
    uint32_t sph = 0;
    float* my_gn = gn_acc + quarter * 64;  // see gemm_conv_tc.cu: per-warp partial sums, flushed when the image changes
    float* my_bias = gn_acc + 256 + quarter * C::kBiasSlice;
    int bias_tile = -1;
    const int epi_mode = epilogue_mode(p.act, false, p.res != nullptr, p.out_f32 != 0, p.gn_sums != nullptr);
    const int gn_n = p.gn_groups * 2;  // <= 64: the slice holds every group of one image, whatever the n_tile
    int gn_b = -1;
    if (p.gn_sums) { my_gn[lane] = 0.f; my_gn[lane + 32] = 0.f; __syncwarp(); }
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      if (p.bias && n_tile != bias_tile) {
        stage_bias<BN>(my_bias, p.bias, n_tile * BN, p.Cout, lane);
        bias_tile = n_tile;
      }
      int m = tile / p.n_tiles;
      const int tx = m % p.tiles_x; m /= p.tiles_x;
      const int ty = m % p.tiles_y;
      const int b = m / p.tiles_y;
      const int x = tx * 128 + row;
      if (p.gn_sums && b != gn_b) {
        if (gn_b >= 0) gn_flush(my_gn, gn_n, gn_slot(p.gn_sums, gn_b, gn_n, blockIdx.x * 4 + quarter, p.gn_slots), lane);
        gn_b = b;
      }
      // residual rows of this thread's TM pixels: pulled into L2 now, then loaded one 32-channel chunk ahead of use
      EpiArgs e;
      e.bias = p.bias; e.gate_row = nullptr; e.out = p.out;
      e.Cout = p.Cout; e.cout_out = p.Cout; e.act = p.act; e.act_c0 = 0; e.act_c1 = 0x7fffffff;
      e.out_f32 = p.out_f32; e.vec32 = p.out_vec32; e.res_vec32 = p.res_vec32; e.res_f32 = 0;
      float gn_part[16];
      e.gn_part = p.gn_sums ? gn_part : nullptr;
      uint4 rr[4] = {};
      if (p.res) {
#pragma unroll
        for (int t = TM - 1; t >= 0; --t) {
          const int y = ty * TM + t;
          e.res_row = p.res + (b * p.r_sb + y * p.r_sy + x * p.r_sx);
          prefetch_res_row(e, n_tile * BN, BN, (x < p.W) && (y < p.H));
        }
        load_res32(rr, e, n_tile * BN, (x < p.W) && (ty * TM < p.H));  // e.res_row is row t = 0 here
      }
      mbar_wait(&acc_full[set], sph);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < TM; ++t) {
        const int y = ty * TM + t;
        const bool valid = (x < p.W) && (y < p.H) && !(p.debug & 1);
        if (p.debug & 2) continue;
        e.o_off = b * p.o_sb + y * p.o_sy + x * p.o_sx;
        const bf16* res_this = p.res ? p.res + (b * p.r_sb + y * p.r_sy + x * p.r_sx) : nullptr;
        const uint32_t t_row = tmem_base + set * (TM * BN) + t * BN + (static_cast<uint32_t>(quarter * 32) << 16);
        e.res_row = res_this;
        const bf16* next_res = (p.res && t + 1 < TM) ? res_this + p.r_sy : nullptr;
        epilogue_row_m<false>(epi_mode, t_row, BN, 0, n_tile * BN, n_tile * BN, valid, e, next_res, n_tile * BN,
                              (x < p.W) && (y + 1 < p.H), p.bias ? my_bias : nullptr, rr, my_gn, p.gn_cpg, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[set]);
      if (C::kAccSets == 2) { if (++set == 2) { set = 0; sph ^= 1; } }
      else { sph ^= 1; }
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
// CTA-pair variant (cta_group::2, see gemm_conv_tc.cu): the two CTAs of a cluster each own one halo tile (TM rows x
// 128 pixels) and HALF (64 rows) of every 128-row weight block; one M=256 tcgen05.mma per (tap, row, K-step) covers
// both. At N = 128 the single-CTA MMA reads 4 KB of A + 4 KB of B from shared memory per 64 cycles (the 128 B/clk
// limit); the pair reads 4 + 2 KB per CTA.
// ------------------------------------------------------------------------------------------------
template <int TM>
struct Halo2Cfg {
  static constexpr int kABytes = (((TM + 2) * 130 * 128) + 1023) / 1024 * 1024;
  static constexpr int kAStages = 2;
  static constexpr int kBBytes = 64 * 128;  // this CTA's half of the 128-row weight block
  static constexpr int kBStages = 8;
  static constexpr int kSmemBytes = kAStages * kABytes + kBStages * kBBytes + 1024 + 256 + 1024 + 16 * 128;
};

// XF = true adds four TRANSFORM warps (320 threads): the input of the convolution is the raw output of the previous
// layer and its GroupNorm + SiLU is applied to the halo tile in shared memory, between the TMA load and the MMAs
// (y = silu(a[b,c] * x + s[b,c]) with the per-(image, channel) coefficients of gn_coef_kernel; halo pixels outside
// the image stay zero = the conv's zero padding). This removes the separate normalise pass: 2 x 1 GB of HBM traffic
// per 128-channel layer at 512^2 x 16. Each CTA's halo load then signals its OWN barrier; the transform warps
// publish the tile through the leader's a_ready barrier (cluster-scope release / acquire).
template <int TM, bool XF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(XF ? 320 : 192, 1)
conv_halo_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const HaloParams p) {
  using C = Halo2Cfg<TM>;
  constexpr int BN = 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // offset, not a cast: keeps the shared address space (LDS / STS)
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::kAStages * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + C::kBStages * C::kBBytes);
  uint64_t* a_full = bars;                      // leader's are live (2 arrivals)
  uint64_t* a_empty = a_full + C::kAStages;     // per CTA
  uint64_t* b_full = a_empty + C::kAStages;     // leader's
  uint64_t* b_empty = b_full + C::kBStages;     // per CTA
  uint64_t* acc_full = b_empty + C::kBStages;   // per CTA
  uint64_t* acc_empty = acc_full + 2;           // leader's: 8 arrivals
  uint64_t* a_ready = acc_empty + 2;            // XF: leader's [stage][half], 8 arrivals (4 transform warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 2 * C::kAStages);
  float* gn_acc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // uniform: see the MMA issuer
  const uint32_t rank = cluster_ctarank();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < C::kAStages; ++i) {
      mbar_init(&a_full[i], XF ? 1 : 2);
      mbar_init(&a_empty[i], 1);
      mbar_init(&a_ready[2 * i], 8);
      mbar_init(&a_ready[2 * i + 1], 8);
    }
    for (int i = 0; i < C::kBStages; ++i) { mbar_init(&b_full[i], 2); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
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

  const int halo_tiles = p.tiles_x * p.tiles_y * p.B;
  const int total_pairs = (halo_tiles + 1) >> 1;   // n_tiles == 1 (Cout <= 128)
  const int taps = p.R * p.S;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
        int m = 2 * pt + static_cast<int>(rank);
        const bool dummy = m >= halo_tiles;
        const int tx = m % p.tiles_x; m /= p.tiles_x;
        const int ty = m % p.tiles_y;
        const int b = dummy ? p.B : m / p.tiles_y;
        for (int c = 0; c < p.chunks; ++c) {
          mbar_wait(&a_empty[as], aph ^ 1);
          if (XF) {  // own barrier: this CTA's transform warps consume the tile first
            mbar_expect_tx(&a_full[as], p.a_bytes);
            tma_load_5d(sA + as * C::kABytes, &tmA, &a_full[as], c * 64, tx * 128 - p.pad_l, ty * TM - p.pad_t, b, 0);
          } else {
            if (rank == 0) mbar_expect_tx(&a_full[as], 2 * p.a_bytes);
            else mbar_arrive_remote(&a_full[as], 0);
            tma2_load_5d(sA + as * C::kABytes, &tmA, &a_full[as], c * 64, tx * 128 - p.pad_l, ty * TM - p.pad_t, b, 0);
          }
          if (++as == C::kAStages) { as = 0; aph ^= 1; }
          for (int tap = 0; tap < taps; ++tap) {
            mbar_wait(&b_empty[bs], bph ^ 1);
            if (rank == 0) mbar_expect_tx(&b_full[bs], 2 * C::kBBytes);
            else mbar_arrive_remote(&b_full[bs], 0);
            tma2_load_4d(sB + bs * C::kBBytes, &tmB, &b_full[bs], tap * p.Cin + c * 64, static_cast<int>(rank) * 64, 0, 0);
            if (++bs == C::kBStages) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {  // whole warp on warp-uniform values, one elected lane issues (gemm_conv_tc.cu has the argument)
      const bool leader = elect_one();
      const uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
      const uint64_t desc0 = umma_desc_sw128(0, 16, 1024);
      const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32), desc_lo = static_cast<uint32_t>(desc0);
      const bool no_mma = (p.debug & 4) != 0;
      int as = 0, bs = 0, set = 0;
      uint32_t aph = 0, bph = 0, sph = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
        mbar_wait(&acc_empty[set], sph ^ 1);
        tc_fence_after();
        const uint32_t acc0 = tmem_base + set * (TM * BN);
        for (int c = 0; c < p.chunks; ++c) {
          // XF: the transform publishes the halo tile in two halves (first TM rows, then the rest): the r = 0 taps only
          // read rows [0, TM) and start while the lower rows are still being normalised
          if (XF) mbar_wait(&a_ready[2 * as], aph);
          else mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + as * C::kABytes);
          int tap = 0;
          for (int r = 0; r < p.R; ++r) {
            if (XF && r == 1) { mbar_wait(&a_ready[2 * as + 1], aph); tc_fence_after(); }
            for (int s = 0; s < p.S; ++s, ++tap) {
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              const uint32_t b_lo = desc_lo | (smem_u32(sB + bs * C::kBBytes) >> 4);
              const uint32_t a_lo = desc_lo | ((a_base + static_cast<uint32_t>((r * p.pw + s) * 128)) >> 4);
              const uint32_t a_row = static_cast<uint32_t>(p.pw * 8);  // one halo row down, in 16-byte units
              if (leader && !no_mma) {
#pragma unroll
                for (int t = 0; t < TM; ++t) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    tc_mma2_ss(acc0 + t * BN, (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + t * a_row + 2 * k),
                               (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2 * k), idesc, (c | tap | k) != 0);
                }
              }
              if (leader) tc_commit2_mc(&b_empty[bs]);
              if (++bs == C::kBStages) { bs = 0; bph ^= 1; }
            }
          }
          if (leader) tc_commit2_mc(&a_empty[as]);
          if (++as == C::kAStages) { as = 0; aph ^= 1; }
        }
        if (leader) tc_commit2_mc(&acc_full[set]);
        if (++set == 2) { set = 0; sph ^= 1; }
      }
    }
  } else if (XF && warp >= 6) {
    // ================= input transform (warps 6..9): GroupNorm + SiLU on the halo tile, in place =================
    const int tid = threadIdx.x - 192;  // 0..127
    const int j = tid & 7;              // 16-byte unit (8 channels) of a 128-byte pixel row: fixed per thread
    const int p0 = tid >> 3;            // first halo pixel of this thread; then every 16th
    const int n_pix = (TM + p.R - 1) * p.pw;
    int as = 0;
    uint32_t aph = 0;
    for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
      int m = 2 * pt + static_cast<int>(rank);
      const bool dummy = m >= halo_tiles;
      const int tx = m % p.tiles_x; m /= p.tiles_x;
      const int ty = m % p.tiles_y;
      const int b = dummy ? 0 : m / p.tiles_y;
      const int x0 = tx * 128 - p.pad_l, y0 = ty * TM - p.pad_t;
      for (int c = 0; c < p.chunks; ++c) {
        float ca[8], cb[8];
        if (!dummy) {
          const float4* cf = reinterpret_cast<const float4*>(p.gn_in + static_cast<long long>(b) * p.Cin + c * 64 + j * 8);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 v = __ldg(cf + q);
            ca[2 * q] = v.x; cb[2 * q] = v.y; ca[2 * q + 1] = v.z; cb[2 * q + 1] = v.w;
          }
        }
        mbar_wait(&a_full[as], aph);
        bool half_sent = false;
        if (!dummy) {
          uint8_t* tile = sA + as * C::kABytes;
          // four independent pixels per iteration (loads, math and stores of the four interleave): one transform
          // warp per scheduler has nothing else to hide the LDS / MUFU latencies behind
          int hx = p0, hy = 0;
          const int half_pix = TM * p.pw;  // rows [0, TM): everything the r = 0 taps read
          for (int pix = p0; pix < n_pix; pix += 64) {
            if (!half_sent && pix - p0 >= half_pix) {  // warp-uniform: every lane's remaining pixels are >= half_pix
              half_sent = true;
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) { if (rank == 0) mbar_arrive(&a_ready[2 * as]); else mbar_arrive_remote_relaxed(&a_ready[2 * as], 0); }
            }
            uint4* u[4];
            uint4 v[4];
            bool on[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int pi = pix + 16 * i;
              const int xx = x0 + hx, yy = y0 + hy;
              // outside the image: TMA wrote zeros = the conv padding, which must stay zero
              on[i] = pi < n_pix && xx >= 0 && xx < p.IW && yy >= 0 && yy < p.IH;
              u[i] = reinterpret_cast<uint4*>(tile + pi * 128 + ((j ^ (pi & 7)) << 4));
              if (on[i]) v[i] = *u[i];
              hx += 16;
              if (hx >= p.pw) { hx -= p.pw; ++hy; }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float f[8];
              const float2 f0 = unpack_bf16x2(v[i].x), f1 = unpack_bf16x2(v[i].y), f2 = unpack_bf16x2(v[i].z),
                           f3 = unpack_bf16x2(v[i].w);
              f[0] = f0.x; f[1] = f0.y; f[2] = f1.x; f[3] = f1.y; f[4] = f2.x; f[5] = f2.y; f[6] = f3.x; f[7] = f3.y;
#pragma unroll
              for (int q = 0; q < 8; ++q) f[q] = silu_f(fmaf(f[q], ca[q], cb[q]));
              v[i].x = pack_bf16x2(f[0], f[1]); v[i].y = pack_bf16x2(f[2], f[3]);
              v[i].z = pack_bf16x2(f[4], f[5]); v[i].w = pack_bf16x2(f[6], f[7]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (on[i]) *u[i] = v[i];
          }
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) {
          // plain (CTA-scope) arrives: the tile is only ever read by THIS CTA's tensor core, and fence.proxy.async has
          // already made the writes visible to it; a .release.cluster arrive compiled to MEMBAR.ALL.GPU (9 % of all
          // stall samples, on the transform -> MMA critical path)
          if (!half_sent) { if (rank == 0) mbar_arrive(&a_ready[2 * as]); else mbar_arrive_remote_relaxed(&a_ready[2 * as], 0); }
          if (rank == 0) mbar_arrive(&a_ready[2 * as + 1]); else mbar_arrive_remote_relaxed(&a_ready[2 * as + 1], 0);
        }
        if (++as == C::kAStages) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp < 6) {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int set = 0;
    uint32_t sph = 0;
    float* my_gn = gn_acc + quarter * 64;
    float* my_bias = gn_acc + 256 + quarter * BN;
    if (p.bias) stage_bias<BN>(my_bias, p.bias, 0, p.Cout, lane);
    const int epi_mode = epilogue_mode(p.act, false, p.res != nullptr, p.out_f32 != 0, p.gn_sums != nullptr);
    const int gn_n = p.gn_groups * 2;
    int gn_b = -1;
    if (p.gn_sums) { my_gn[lane] = 0.f; my_gn[lane + 32] = 0.f; __syncwarp(); }
    for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
      int m = 2 * pt + static_cast<int>(rank);
      const bool dummy = m >= halo_tiles;
      const int tx = m % p.tiles_x; m /= p.tiles_x;
      const int ty = m % p.tiles_y;
      const int b = dummy ? 0 : m / p.tiles_y;
      const int x = tx * 128 + row;
      if (p.gn_sums && !dummy && b != gn_b) {
        if (gn_b >= 0) gn_flush(my_gn, gn_n, gn_slot(p.gn_sums, gn_b, gn_n, blockIdx.x * 4 + quarter, p.gn_slots), lane);
        gn_b = b;
      }
      // residual rows of this thread's TM pixels: pulled into L2 now, then loaded one 32-channel chunk ahead of use
      EpiArgs e;
      e.bias = p.bias; e.gate_row = nullptr; e.out = p.out;
      e.Cout = p.Cout; e.cout_out = p.Cout; e.act = p.act; e.act_c0 = 0; e.act_c1 = 0x7fffffff;
      e.out_f32 = p.out_f32; e.vec32 = p.out_vec32; e.res_vec32 = p.res_vec32; e.res_f32 = 0;
      float gn_part[16];
      e.gn_part = p.gn_sums ? gn_part : nullptr;
      uint4 rr[4] = {};
      const bool use_res = p.res && !dummy;
      if (use_res) {
#pragma unroll
        for (int t = TM - 1; t >= 0; --t) {
          const int y = ty * TM + t;
          e.res_row = p.res + (b * p.r_sb + y * p.r_sy + x * p.r_sx);
          prefetch_res_row(e, 0, BN, (x < p.W) && (y < p.H));
        }
        load_res32(rr, e, 0, (x < p.W) && (ty * TM < p.H));  // e.res_row is row t = 0 here
      }
      mbar_wait(&acc_full[set], sph);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < TM; ++t) {
        const int y = ty * TM + t;
        const bool valid = !dummy && (x < p.W) && (y < p.H) && !(p.debug & 1);
        if (p.debug & 2) continue;
        e.o_off = b * p.o_sb + y * p.o_sy + x * p.o_sx;
        const bf16* res_this = use_res ? p.res + (b * p.r_sb + y * p.r_sy + x * p.r_sx) : nullptr;
        const uint32_t t_row = tmem_base + set * (TM * BN) + t * BN + (static_cast<uint32_t>(quarter * 32) << 16);
        e.res_row = res_this;
        const bf16* next_res = (use_res && t + 1 < TM) ? res_this + p.r_sy : nullptr;
        epilogue_row_m<false>(epi_mode, t_row, BN, 0, 0, 0, valid, e, next_res, 0, (x < p.W) && (y + 1 < p.H),
                              p.bias ? my_bias : nullptr, rr, my_gn, p.gn_cpg, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&acc_empty[set]);
        else mbar_arrive_remote(&acc_empty[set], 0);
      }
      if (++set == 2) { set = 0; sph ^= 1; }
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

// ------------------------------------------------------------------------------------------------
// Row-shared CTA-pair kernel (round 2).
//
// Design history: built when a probe showed ~130-170 cycles per tcgen05.mma whatever N <= 128 is, which made N = 256
// instructions look twice as cheap per column. That figure was the single-lane issue loop (the
// loops below are warp-uniform now and an N = 128 instruction costs its 64 pipe cycles). The pairing still halves the
// instruction count of the shared rows; conv_halo_rr_kernel below supersedes this kernel wherever it applies (this one
// remains for the 2x2 parity convolutions of the fused upsample and for short images). A 3x3 convolution with 128
// output channels has N = 256 instructions in it: for a fixed tap column s, output row t with tap row r reads
// input row rho = t + r, so (t, r) and (t + 1, r - 1) read the SAME A operand. With the two accumulators of a TM = 2
// tile adjacent in TMEM, one instruction
//     D[256 px pair, 256] (+)= A[row rho, shifted by s] x [ W(r = rho, s) | W(r = rho - 1, s) ]^T
// feeds both. Per (64-channel chunk, s) the six N = 128 instructions become N = 128, 256, 256, 128 over input rows
// 0..3: 2.33-2.40 k instead of 3.04-3.38 k cycles on the probe, and a third fewer A reads.
//   * cta_group::2 splits the B rows of an instruction between the CTAs, so for the N = 256 instructions CTA 0 holds
//     all 128 rows of W(rho, s) (-> accumulator t0, columns 0..127) and CTA 1 all of W(rho - 1, s) (-> t1); for the
//     N = 128 leftovers (rho = 0 -> t0, rho = TM + R - 2 -> t1) each CTA holds its 64-row half as before.
//   * an input ROW is the pipeline stage: stage rho = one halo row (17 KB slot, 1024-byte aligned so it is its own
//     swizzle domain) + the S weight units that multiply it (S x 16 KB or S x 8 KB), at FIXED shared-memory offsets.
//     Loop order chunk -> row -> s; ONE tcgen05.commit per row releases both (the probe: a commit behind short
//     instructions costs 60-120 cycles, so per-unit commits ate the gain), and the stage is refilled with the same row of
//     the next chunk / tile, i.e. operands are requested ~3 rows (5-6 k cycles) ahead of their use. Rows and weights have
//     their own producer warps. Row 1 is visited first: its N = 256 instruction initialises BOTH accumulators.
// Epilogue, statistics and the transform arithmetic are those of conv_halo_tc2_kernel.
// ------------------------------------------------------------------------------------------------
struct HaloRsCfg {
  static constexpr int TM = 2;
  static constexpr int kRowBytes = 17 * 1024;  // 130 px x 128 B = 16640 B, padded to a multiple of 1024
  static constexpr int kMaxRows = 4;           // TM + R - 1 input rows per tile
  static constexpr int kBBytes = 144 * 1024;   // R = S = 3: 24 + 48 + 48 + 24 KB of weight units per chunk
  static constexpr int kBarBytes = 512;
  // + statistics partial sums (8 epilogue warps x 64 floats) + bias slices (8 x 128 floats)
  static constexpr int kSmemBytes = kMaxRows * kRowBytes + kBBytes + 1024 + kBarBytes + 2048 + 32 * 128;
};

// input rows of one chunk in visiting order: 1 first (see above), then 0, 2, 3, ...
__device__ __forceinline__ int rs_row(int i) { return i == 0 ? 1 : (i == 1 ? 0 : i); }
// byte offset of row rho's weight units inside the B region: leftover rows hold S x 8 KB, shared rows S x 16 KB
__device__ __forceinline__ int rs_boff(int rho, int R, int S) {
  // rows before rho: row 0 (leftover) and rows 1 .. rho - 1 (shared, as long as they are <= R - 1)
  if (rho == 0) return 0;
  const int shared_before = (rho - 1 < R - 1) ? rho - 1 : R - 1;
  return S * 8192 + shared_before * S * 16384 + (rho - 1 - shared_before) * S * 8192;
}

template <bool XF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(XF ? 480 : 352, 1)
conv_halo_rs_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB128,
                    const __grid_constant__ CUtensorMap tmB64, const HaloParams p) {
  using C = HaloRsCfg;
  constexpr int BN = 128, TM = C::TM, NS = C::kMaxRows;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // offset, not a cast: keeps the shared address space (LDS / STS)
  uint8_t* sA = smem;
  uint8_t* sB = smem + NS * C::kRowBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + C::kBBytes);
  uint64_t* a_full = bars;                // [NS] XF: per CTA (own bytes); else leader's (2 arrivals)
  uint64_t* a_ready = a_full + NS;        // [NS] XF: leader's, 8 arrivals (4 transform warps x 2 CTAs)
  uint64_t* b_full = a_ready + NS;        // [NS] leader's (2 arrivals)
  uint64_t* empty = b_full + NS;          // [NS] per CTA: one multicast commit per row frees the row AND its weights
  uint64_t* acc_full = empty + NS;        // [2] per CTA
  uint64_t* acc_empty = acc_full + 2;     // [2] leader's: 16 arrivals (8 epilogue warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* gn_acc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + C::kBarBytes);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // uniform: see the MMA issuer
  const uint32_t rank = cluster_ctarank();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB128);
    tma_prefetch_desc(&tmB64);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&a_full[i], XF ? 1 : 2);
      mbar_init(&a_ready[i], 8);
      mbar_init(&b_full[i], 2);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 16); }
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

  const int halo_tiles = p.tiles_x * p.tiles_y * p.B;
  const int total_pairs = (halo_tiles + 1) >> 1;   // n_tiles == 1 (Cout <= 128)
  const int n_rows = TM + p.R - 1;                 // input rows per tile (<= NS)
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int row_bytes = p.pw * 128;
  // warps: 0 row producer, 1 MMA issuer, 2..5 epilogue of accumulator row t0, 6..9 epilogue of t1 (two epilogue warps
  // per scheduler: with one, every TMEM / residual-load latency was exposed and the residual + statistics variants were
  // epilogue-bound at 1.7 ms), 10 weight producer, XF: 11..14 input transform
  constexpr int kBWarp = 10;

  // every role walks the same (tile pair, chunk, row) sequence; stage rho is used once per chunk, so the parity of a
  // stage's barriers is the parity of the running chunk count
  if (warp == 0) {
    if (lane == 0) {
      // ================= input-row producer =================
      uint32_t ph = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
        int m = 2 * pt + static_cast<int>(rank);
        const bool dummy = m >= halo_tiles;
        const int tx = m % p.tiles_x; m /= p.tiles_x;
        const int ty = m % p.tiles_y;
        const int b = dummy ? p.B : m / p.tiles_y;
        for (int c = 0; c < p.chunks; ++c, ph ^= 1) {
          for (int i = 0; i < n_rows; ++i) {
            const int rho = rs_row(i);
            mbar_wait_relaxed(&empty[rho], ph ^ 1);
            if (XF) {  // own barrier: this CTA's transform warps consume the row first
              mbar_expect_tx(&a_full[rho], row_bytes);
              tma_load_5d(sA + rho * C::kRowBytes, &tmA, &a_full[rho], c * 64, tx * 128 - p.pad_l, ty * TM - p.pad_t + rho, b, 0);
            } else {
              if (rank == 0) mbar_expect_tx(&a_full[rho], 2 * row_bytes);
              else mbar_arrive_remote(&a_full[rho], 0);
              tma2_load_5d(sA + rho * C::kRowBytes, &tmA, &a_full[rho], c * 64, tx * 128 - p.pad_l, ty * TM - p.pad_t + rho, b, 0);
            }
          }
        }
      }
    }
  } else if (warp == kBWarp) {
    if (lane == 0) {
      // ================= weight-unit producer =================
      uint32_t ph = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
        for (int c = 0; c < p.chunks; ++c, ph ^= 1) {
          for (int i = 0; i < n_rows; ++i) {
            const int rho = rs_row(i);
            // weight units of this input row: both accumulators (N = 256: this CTA supplies all 128 rows of ITS tap
            // row) or one (N = 128: this CTA supplies its 64-row half)
            const bool both = rho >= 1 && rho <= p.R - 1;
            const int r_mine = both ? rho - static_cast<int>(rank) : (rho == 0 ? 0 : p.R - 1);
            const int unit = both ? 16384 : 8192;
            uint8_t* dst = sB + rs_boff(rho, p.R, p.S);
            mbar_wait_relaxed(&empty[rho], ph ^ 1);
            if (rank == 0) mbar_expect_tx(&b_full[rho], 2 * p.S * unit);
            else mbar_arrive_remote(&b_full[rho], 0);
            for (int sx = 0; sx < p.S; ++sx) {
              const int kcol = (r_mine * p.S + sx) * p.Cin + c * 64;
              if (both) tma2_load_4d(dst + sx * unit, &tmB128, &b_full[rho], kcol, 0, 0, 0);
              else tma2_load_4d(dst + sx * unit, &tmB64, &b_full[rho], kcol, static_cast<int>(rank) * 64, 0, 0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ================= MMA issuer =================
      // whole warp on warp-uniform values, one elected lane issues (gemm_conv_tc.cu has the argument)
      const bool leader = elect_one();
      const uint32_t idesc256 = umma_idesc_bf16(256, 256, 0, 0), idesc128 = umma_idesc_bf16(256, 128, 0, 0);
      const uint64_t desc0 = umma_desc_sw128(0, 16, 1024);
      const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32), desc_lo = static_cast<uint32_t>(desc0);
      const bool no_mma = (p.debug & 4) != 0;  // timing-only ablation (OMGSR_HALO_DEBUG): no MMAs, results invalid
      int set = 0;
      uint32_t ph = 0, sph = 0;
      for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
        mbar_wait(&acc_empty[set], sph ^ 1);
        tc_fence_after();
        const uint32_t acc0 = tmem_base + set * (TM * BN);
        uint32_t first = 1;  // the first instruction of a tile (row 1: N = 256) initialises both accumulators
        for (int c = 0; c < p.chunks; ++c, ph ^= 1) {
          for (int i = 0; i < n_rows; ++i) {
            const int rho = rs_row(i);
            if (XF) mbar_wait(&a_ready[rho], ph);
            else mbar_wait(&a_full[rho], ph);
            mbar_wait(&b_full[rho], ph);
            tc_fence_after();
            const uint32_t a_row = smem_u32(sA + rho * C::kRowBytes);
            const bool both = rho >= 1 && rho <= p.R - 1;
            const uint32_t d = both ? acc0 : (rho == 0 ? acc0 : acc0 + (TM - 1) * BN);
            const uint32_t idesc = both ? idesc256 : idesc128;
            const uint32_t b_row = smem_u32(sB + rs_boff(rho, p.R, p.S));
            const uint32_t unit = both ? 16384u : 8192u;
            for (int sx = 0; sx < p.S; ++sx) {
              const uint32_t b_lo = desc_lo | ((b_row + sx * unit) >> 4);
              const uint32_t a_lo = desc_lo | ((a_row + static_cast<uint32_t>(sx * 128)) >> 4);
              if (leader && !no_mma) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  tc_mma2_ss(d, (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2 * k),
                             (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2 * k), idesc, (first && k == 0) ? 0u : 1u);
              }
              first = 0;
            }
            if (leader) tc_commit2_mc(&empty[rho]);  // the row and its weight units are free once everything issued so far retires
          }
        }
        if (leader) tc_commit2_mc(&acc_full[set]);
        if (++set == 2) { set = 0; sph ^= 1; }
      }
    }
  } else if (XF && warp >= 11) {
    // ================= input transform (warps 11..14): GroupNorm + SiLU on each halo row, in place =================
    const int tid = threadIdx.x - 352;  // 0..127
    const int j = tid & 7;              // 16-byte unit (8 channels) of a 128-byte pixel row: fixed per thread
    const int p0 = tid >> 3;            // first pixel of this thread in a row; then every 16th (<= 9 per row)
    uint32_t ph = 0;
    for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
      int m = 2 * pt + static_cast<int>(rank);
      const bool dummy = m >= halo_tiles;
      const int tx = m % p.tiles_x; m /= p.tiles_x;
      const int ty = m % p.tiles_y;
      const int b = dummy ? 0 : m / p.tiles_y;
      const int x0 = tx * 128 - p.pad_l, y0 = ty * TM - p.pad_t;
      for (int c = 0; c < p.chunks; ++c, ph ^= 1) {
        float ca[8], cb[8];
        if (!dummy) {
          const float4* cf = reinterpret_cast<const float4*>(p.gn_in + static_cast<long long>(b) * p.Cin + c * 64 + j * 8);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 v = __ldg(cf + q);
            // halved: silu(a x + b) = h + h tanh(h) with h = (a / 2) x + b / 2
            ca[2 * q] = 0.5f * v.x; cb[2 * q] = 0.5f * v.y; ca[2 * q + 1] = 0.5f * v.z; cb[2 * q + 1] = 0.5f * v.w;
          }
        }
        for (int i = 0; i < n_rows; ++i) {
          const int rho = rs_row(i);
          const int yy = y0 + rho;
          mbar_wait_relaxed(&a_full[rho], ph, 32);
          // rows / pixels outside the image: TMA wrote zeros = the conv padding, which must stay zero
          if (!dummy && yy >= 0 && yy < p.IH) {
            uint8_t* row = sA + rho * C::kRowBytes;
            // ALL of this thread's units of the row (<= 9) are in flight at once: loads first, then the math, then the
            // stores. One transform warp per scheduler has nothing else to hide LDS / MUFU latency behind (three
            // units in flight: 2.1 k cycles per row, slower than the row's MMAs).
            uint4 v[9];
            bool on[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) {
              const int pi = p0 + 16 * q;
              const int xx = x0 + pi;
              on[q] = pi < p.pw && xx >= 0 && xx < p.IW;
              if (on[q]) v[q] = *reinterpret_cast<const uint4*>(row + pi * 128 + ((j ^ (pi & 7)) << 4));
            }
#pragma unroll
            for (int q = 0; q < 9; ++q) {
              float f[8];
              const float2 f0 = unpack_bf16x2(v[q].x), f1 = unpack_bf16x2(v[q].y), f2 = unpack_bf16x2(v[q].z),
                           f3 = unpack_bf16x2(v[q].w);
              f[0] = f0.x; f[1] = f0.y; f[2] = f1.x; f[3] = f1.y; f[4] = f2.x; f[5] = f2.y; f[6] = f3.x; f[7] = f3.y;
#pragma unroll
              for (int e = 0; e < 8; e += 2) {  // packed fp32 pairs: h = (a / 2) x + b / 2, y = h + h tanh(h)
                float h0, h1;
                ffma2p(h0, h1, f[e], f[e + 1], ca[e], ca[e + 1], cb[e], cb[e + 1]);
                ffma2p(f[e], f[e + 1], h0, h1, tanh_approx(h0), tanh_approx(h1), h0, h1);
              }
              v[q].x = pack_bf16x2(f[0], f[1]); v[q].y = pack_bf16x2(f[2], f[3]);
              v[q].z = pack_bf16x2(f[4], f[5]); v[q].w = pack_bf16x2(f[6], f[7]);
            }
#pragma unroll
            for (int q = 0; q < 9; ++q) {
              const int pi = p0 + 16 * q;
              if (on[q]) *reinterpret_cast<uint4*>(row + pi * 128 + ((j ^ (pi & 7)) << 4)) = v[q];
            }
          }
          fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) {
            // plain / relaxed arrives: the row is only ever read by THIS CTA's tensor core, and fence.proxy.async has
            // already made the writes visible to it (a .release.cluster arrive compiles to MEMBAR.ALL.GPU)
            if (rank == 0) mbar_arrive(&a_ready[rho]); else mbar_arrive_remote_relaxed(&a_ready[rho], 0);
          }
        }
      }
    }
  } else if (warp >= 2 && warp < 10) {
    // ================= epilogue: warps 2..5 drain accumulator row t0, warps 6..9 row t1 =================
    const int t = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int ew = t * 4 + quarter;       // epilogue warp index inside the CTA: statistics slot / bias slice
    const int row = quarter * 32 + lane;
    int set = 0;
    uint32_t sph = 0;
    float* my_gn = gn_acc + ew * 64;
    float* my_bias = gn_acc + 8 * 64 + ew * BN;
    if (p.bias) stage_bias<BN>(my_bias, p.bias, 0, p.Cout, lane);
    const int epi_mode = epilogue_mode(p.act, false, p.res != nullptr, p.out_f32 != 0, p.gn_sums != nullptr);
    const int gn_n = p.gn_groups * 2;
    int gn_b = -1;
    if (p.gn_sums) { my_gn[lane] = 0.f; my_gn[lane + 32] = 0.f; __syncwarp(); }
    for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
      int m = 2 * pt + static_cast<int>(rank);
      const bool dummy = m >= halo_tiles;
      const int tx = m % p.tiles_x; m /= p.tiles_x;
      const int ty = m % p.tiles_y;
      const int b = dummy ? 0 : m / p.tiles_y;
      const int x = tx * 128 + row, y = ty * TM + t;
      if (p.gn_sums && !dummy && b != gn_b) {
        if (gn_b >= 0) gn_flush(my_gn, gn_n, gn_slot(p.gn_sums, gn_b, gn_n, blockIdx.x * 8 + ew, p.gn_slots), lane);
        gn_b = b;
      }
      EpiArgs e;
      e.bias = p.bias; e.gate_row = nullptr; e.out = p.out;
      e.Cout = p.Cout; e.cout_out = p.Cout; e.act = p.act; e.act_c0 = 0; e.act_c1 = 0x7fffffff;
      e.out_f32 = p.out_f32; e.vec32 = p.out_vec32; e.res_vec32 = p.res_vec32; e.res_f32 = 0;
      float gn_part[16];
      e.gn_part = p.gn_sums ? gn_part : nullptr;
      const bool valid = !dummy && (x < p.W) && (y < p.H) && !(p.debug & 1);
      uint4 rr[4] = {};
      e.res_row = (p.res && !dummy) ? p.res + (b * p.r_sb + y * p.r_sy + x * p.r_sx) : nullptr;
      if (e.res_row) {  // this thread's residual row: pulled into L2 now, first chunk loaded before the accumulator wait
        prefetch_res_row(e, 0, BN, valid);
        load_res32(rr, e, 0, valid);
      }
      e.o_off = b * p.o_sb + y * p.o_sy + x * p.o_sx;
      mbar_wait_relaxed(&acc_full[set], sph);
      tc_fence_after();
      if (!(p.debug & 2)) {  // (timing-only ablation: no epilogue)
        const uint32_t t_row = tmem_base + set * (TM * BN) + t * BN + (static_cast<uint32_t>(quarter * 32) << 16);
        epilogue_row_m<false>(epi_mode, t_row, BN, 0, 0, 0, valid, e, nullptr, 0, false, p.bias ? my_bias : nullptr, rr,
                              my_gn, p.gn_cpg, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&acc_empty[set]);
        else mbar_arrive_remote(&acc_empty[set], 0);
      }
      if (++set == 2) { set = 0; sph ^= 1; }
    }
    if (p.gn_sums && gn_b >= 0) gn_flush(my_gn, gn_n, gn_slot(p.gn_sums, gn_b, gn_n, blockIdx.x * 8 + ew, p.gn_slots), lane);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Rolling-row kernel (round 2, 3x3 only): conv_halo_rs_kernel loads (and, with the fused input GroupNorm, TRANSFORMS)
// 4 input rows per 2 output rows, i.e. every input row twice; the transform pipeline alone takes 0.80 ms of the 0.9 ms
// of MMAs on the 512^2 x 16 layer (profiles/r02/halo_rowshared_ablation_v3_8epi.txt) and the two stages do not overlap
// perfectly through 4 row slots. Here a CTA pair walks DOWN a pair of 128-pixel column strips: each input row is loaded
// and transformed ONCE and feeds the three output rows it touches (tap rows 2, 1, 0 of outputs y-1, y, y+1).
//   * TMEM: a ring of four 128-column accumulators; output n (running count) lives in slot n & 3. Three are being
//     accumulated while the fourth is drained. The epilogue ZEROES a slot after reading it, so every MMA accumulates
//     (enable-input-d = 1) and an N = 256 instruction may cover ANY two adjacent slots - also a fresh one next to an
//     old one (one instruction has one accumulate flag). Per input row, chunk and tap column: one N = 256 + one N = 128
//     instruction, the same 315 cycles per three row-taps as the row-shared schedule.
//   * unit of the pipeline = (input row, 64-channel chunk): A ring of seven 17 KB rows (the rows come from DRAM and, with
//     the input transform, pass through a second stage before the MMAs: with four slots the issuer waited on the
//     transform and the transform on TMA, ncu profiles/r02/ncu_r02_halo_rr_gn_lean.txt). The weights of a unit arrive
//     per tap column: B ring of four 24 KB stages (pair unit 16 KB: this CTA's 128 rows of ITS tap row; single unit
//     8 KB: its 64-row half); one tcgen05.commit per B stage and one per unit (A slot).
//   * work split: the B * ceil(tiles_x / 2) * H pair-rows are cut into one contiguous range per cluster (a range crosses
//     at most a few strip boundaries; each segment costs two extra input rows), so all clusters finish together.
//   * residual (Cout = 128): accumulated as an IDENTITY TAP - the residual row of output o is one more A row (TMA, full
//     lines from DRAM instead of per-thread 32-byte loads), multiplied by the 128 x 128 identity (8 N = 128 instructions
//     per output row, +11 % MMA time; exact: 1.0 x bf16 into the fp32 accumulator). The epilogue then has no residual
//     loads, unpacks or adds: the heaviest variant (input GroupNorm + residual + statistics) went from 1.24 ms to the
//     cost of the variant without residual plus the extra MMAs.
//   * epilogue: warps 4..7 take columns 0..63 of every output row, warps 8..11 columns 64..127. A slot must be free
//     again ONE row-step (~6.5 k cycles) after it completes, so a warp copies its 64 columns to registers, zeroes
//     them and returns the slot before it does any arithmetic or stores.
// ------------------------------------------------------------------------------------------------
struct HaloRrCfg {
  static constexpr int kRowBytes = 17 * 1024;
  static constexpr int kASlots = 7;             // (input row, chunk) units in flight: the rows come from DRAM
  static constexpr int kBSlotBytes = 24 * 1024; // weight units of ONE tap column of a unit: 16 KB pair + 8 KB single
  static constexpr int kBSlots = 4;
  static constexpr int kBarBytes = 512;
  // + statistics partials (8 warps x 64 floats) + bias slices (8 x 128 floats) + input GroupNorm coefficient table (4 KB: 512 channels)
  static constexpr int kSmemBytes = kASlots * kRowBytes + kBSlots * kBSlotBytes + 1024 + kBarBytes + 2048 + 32 * 128 + 4096;
};

// the MMAs of one input row (step i of a segment of L output rows whose first output has running index n0): up to two
// ops, a pair (N = 256 over slots lo, lo + 1) and / or singles (N = 128)
struct RrPlan {
  int n_ops;
  int pair[2];   // 1: N = 256
  int slot[2];   // first TMEM slot
  int tap0[2];   // tap row of the (first) accumulator
  int tap1[2];   // pair: tap row of the second accumulator
  int boff[2];   // byte offset of the op's weight unit inside a B stage (one tap column)
};
__device__ __forceinline__ RrPlan rr_plan(int i, int L, uint32_t n0) {
  RrPlan pl;
  const int lo = i - 2 > 0 ? i - 2 : 0, hi = i < L - 1 ? i : L - 1;
  const int cnt = hi - lo + 1;
  const int s_lo = static_cast<int>((n0 + lo) & 3u);
  auto set = [&](int k, int is_pair, int o, int boff) {
    pl.pair[k] = is_pair; pl.slot[k] = static_cast<int>((n0 + o) & 3u); pl.tap0[k] = i - o; pl.tap1[k] = i - o - 1;
    pl.boff[k] = boff;
  };
  if (cnt == 3) {
    pl.n_ops = 2;
    if (s_lo != 3) { set(0, 1, lo, 0); set(1, 0, hi, 16384); }
    else { set(0, 0, lo, 16384); set(1, 1, lo + 1, 0); }
  } else if (cnt == 2) {
    if (s_lo != 3) { pl.n_ops = 1; set(0, 1, lo, 0); }
    else { pl.n_ops = 2; set(0, 0, lo, 0); set(1, 0, hi, 8192); }
  } else {
    pl.n_ops = 1; set(0, 0, lo, 0);
  }
  return pl;
}

// MODE: the epilogue option set (EPI_*), a template parameter so that ONE epilogue body is compiled in: with the
// run-time switch the kernel had 11 k instructions (7.9 k of them epilogue) and 'no instruction' was the top stall reason
// of every role, the MMA issuer included (ncu, profiles/r02/ncu_halo_rr_gn_res_v1.txt).
template <bool XF, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(XF ? 512 : 384, 1)
conv_halo_rr_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB128,
                    const __grid_constant__ CUtensorMap tmB64, const __grid_constant__ CUtensorMap tmR,
                    const __grid_constant__ CUtensorMap tmI64, const HaloParams p) {
  using C = HaloRrCfg;
  constexpr int BN = 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // offset, not a cast: keeps the shared address space (LDS / STS)
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::kASlots * C::kRowBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + C::kBSlots * C::kBSlotBytes);
  constexpr int NA = C::kASlots, NB = C::kBSlots;
  uint64_t* a_full = bars;                // [NA] XF: per CTA (own bytes); else leader's (2 arrivals)
  uint64_t* a_ready = a_full + NA;        // [NA] XF: leader's, 8 arrivals (4 transform warps x 2 CTAs)
  uint64_t* a_empty = a_ready + NA;       // [NA] per CTA: one multicast commit per unit
  uint64_t* b_full = a_empty + NA;        // [NB] leader's (2 arrivals)
  uint64_t* b_empty = b_full + NB;        // [NB] per CTA: one multicast commit per tap column
  uint64_t* acc_full = b_empty + NB;      // [4] per CTA
  uint64_t* acc_empty = acc_full + 4;     // [4] leader's: 16 arrivals (8 epilogue warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 4);
  float* gn_acc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + C::kBarBytes);

  // warp index through a shuffle: the compiler then knows the role branches are warp-uniform and keeps the MMA
  // issuer's address arithmetic in uniform registers (see the MMA issuer below)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB128);
    tma_prefetch_desc(&tmB64);
    if (p.res_tap) { tma_prefetch_desc(&tmR); tma_prefetch_desc(&tmI64); }
    for (int i = 0; i < NA; ++i) {
      mbar_init(&a_full[i], XF ? 1 : 2);
      mbar_init(&a_ready[i], 8);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < NB; ++i) { mbar_init(&b_full[i], 2); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 16); }
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
  // warps: 0 row producer, 1 MMA issuer, 2 weight producer, 3 idle | 4..11 epilogue | XF: 12..15 input transform.
  // One role class per warpgroup so that the register file can be re-split (XF: 512 threads start at 128 registers;
  // the epilogue holds 64 accumulator columns + residual + statistics partials and spilled at that cap).
  // (the setmaxnreg of each class is the first instruction of its branch below: ptxas sizes a region by the
  // setmaxnreg that dominates it)
  if (warp >= 4 && warp < 12) {  // all four accumulators start at zero (every MMA accumulates)
    const uint32_t z[32] = {};
    const uint32_t t0 = tmem_base + ((warp - 4) >> 2) * 256 + (static_cast<uint32_t>((warp & 3) * 32) << 16);
#pragma unroll
    for (int c0 = 0; c0 < 256; c0 += 32) tmem_st32(t0 + c0, z);
    tmem_wait_st();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();  // nothing above touches global memory

  // this cluster's contiguous range of pair-rows: pair-strip q = (image, pair of 128-pixel strips), H rows each
  // (Cout = 256: two passes of 128 output channels over the same pair-strip, back to back in the range, so the second
  // pass finds the input rows in L2)
  const int xp = (p.tiles_x + 1) >> 1;
  const int nt = p.n_tiles;
  const long long G = static_cast<long long>(p.B) * xp * p.H * nt;  // < 2^31 (host check)
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int g_begin = static_cast<int>(G * cluster_id / num_clusters), g_end = static_cast<int>(G * (cluster_id + 1) / num_clusters);
  const int row_bytes = p.pw * 128;
  constexpr int kBWarp = 2;
  // segment decode shared by all roles
  int cur_n_tile = 0;  // output-channel tile (128 wide) of the current segment
  auto segment = [&](int g, int& b, int& tx, bool& dummy, int& ya, int& L) {
    const int q0 = g / p.H;
    ya = g - q0 * p.H;
    L = min(p.H - ya, g_end - g);
    cur_n_tile = q0 % nt;
    const int q = q0 / nt;
    b = q / xp;
    tx = 2 * (q - b * xp) + static_cast<int>(rank);
    dummy = tx >= p.tiles_x;
  };

  if (warp < 4) {
  if (XF) setmaxnreg_dec<56>();
  if (warp == 0) {
    if (lane == 0) {
      // ================= input-row producer =================
      int slot = 0;
      uint32_t ph = 0;
      for (int g = g_begin; g < g_end;) {
        int b, tx, ya, L; bool dummy;
        segment(g, b, tx, dummy, ya, L);
        if (dummy) b = p.B;  // out of range: TMA fills zeros
        for (int i = 0; i < L + 2; ++i) {
          for (int c = 0; c < p.chunks; ++c) {
            mbar_wait_relaxed(&a_empty[slot], ph ^ 1);
            uint8_t* dst = sA + slot * C::kRowBytes;
            if (XF) {  // own barrier: this CTA's transform warps consume the row first
              mbar_expect_tx(&a_full[slot], row_bytes);
              tma_load_5d(dst, &tmA, &a_full[slot], c * 64, tx * 128 - p.pad_l, ya - p.pad_t + i, b, 0);
            } else {
              if (rank == 0) mbar_expect_tx(&a_full[slot], 2 * row_bytes);
              else mbar_arrive_remote(&a_full[slot], 0);
              tma2_load_5d(dst, &tmA, &a_full[slot], c * 64, tx * 128 - p.pad_l, ya - p.pad_t + i, b, 0);
            }
            if (++slot == NA) { slot = 0; ph ^= 1; }
          }
          // residual as an identity tap: the residual row of output i - 1 follows the conv units of input row i
          if (p.res_tap && i >= 1 && i <= L) {
            for (int c = 0; c < 2; ++c) {
              mbar_wait_relaxed(&a_empty[slot], ph ^ 1);
              uint8_t* dst = sA + slot * C::kRowBytes;
              if (XF) {
                mbar_expect_tx(&a_full[slot], row_bytes);
                tma_load_5d(dst, &tmR, &a_full[slot], c * 64, tx * 128 - p.pad_l, ya + i - 1, b, 0);
              } else {
                if (rank == 0) mbar_expect_tx(&a_full[slot], 2 * row_bytes);
                else mbar_arrive_remote(&a_full[slot], 0);
                tma2_load_5d(dst, &tmR, &a_full[slot], c * 64, tx * 128 - p.pad_l, ya + i - 1, b, 0);
              }
              if (++slot == NA) { slot = 0; ph ^= 1; }
            }
          }
        }
        g += L;
      }
    }
  } else if (warp == kBWarp) {
    if (lane == 0) {
      // ================= weight-unit producer =================
      uint32_t n0 = 0;
      int bs = 0;
      uint32_t bph = 0;
      for (int g = g_begin; g < g_end;) {
        int b, tx, ya, L; bool dummy;
        segment(g, b, tx, dummy, ya, L);
        for (int i = 0; i < L + 2; ++i) {
          const RrPlan pl = rr_plan(i, L, n0);
          int bytes = 0;
          for (int k = 0; k < pl.n_ops; ++k) bytes += pl.pair[k] ? 16384 : 8192;
          for (int c = 0; c < p.chunks; ++c) {
            for (int sx = 0; sx < 3; ++sx) {
              uint8_t* dst = sB + bs * C::kBSlotBytes;
              mbar_wait_relaxed(&b_empty[bs], bph ^ 1);
              if (rank == 0) mbar_expect_tx(&b_full[bs], 2 * bytes);
              else mbar_arrive_remote(&b_full[bs], 0);
              for (int k = 0; k < pl.n_ops; ++k) {
                const int r_mine = pl.pair[k] ? (rank == 0 ? pl.tap0[k] : pl.tap1[k]) : pl.tap0[k];
                const int kcol = (r_mine * 3 + sx) * p.Cin + c * 64;
                if (pl.pair[k]) tma2_load_4d(dst + pl.boff[k], &tmB128, &b_full[bs], kcol, cur_n_tile * 128, 0, 0);
                else tma2_load_4d(dst + pl.boff[k], &tmB64, &b_full[bs], kcol, cur_n_tile * 128 + static_cast<int>(rank) * 64, 0, 0);
              }
              if (++bs == NB) { bs = 0; bph ^= 1; }
            }
          }
          if (p.res_tap && i >= 1 && i <= L) {  // identity weights of the residual tap: this CTA's 64 rows, one stage
            for (int c = 0; c < 2; ++c) {
              uint8_t* dst = sB + bs * C::kBSlotBytes;
              mbar_wait_relaxed(&b_empty[bs], bph ^ 1);
              if (rank == 0) mbar_expect_tx(&b_full[bs], 2 * 8192);
              else mbar_arrive_remote(&b_full[bs], 0);
              tma2_load_4d(dst, &tmI64, &b_full[bs], c * 64, static_cast<int>(rank) * 64, 0, 0);
              if (++bs == NB) { bs = 0; bph ^= 1; }
            }
          }
        }
        g += L; n0 += L;
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ================= MMA issuer =================
      // The WHOLE warp runs the loop on warp-uniform values and one elected lane issues. Issuing from a divergent
      // single-lane branch (all kernels of round 1 did) makes every MMA operand thread-private to the compiler: each
      // tcgen05.mma then sits in an ELECT / 7 x R2UR.BROADCAST / BRA.U.ANY sequence, ~20 dependent instructions =
      // 130-170 cycles per MMA - the "instruction cost floor" of profiles/r02/mma_probe*.txt was this loop, not the
      // tensor pipe (profiles/r02/mma_probe6_lean.txt: 64 cycles per N = 128 instruction, 128 per N = 256 = 99.5 %).
      const bool leader = elect_one();
      const uint32_t idesc256 = umma_idesc_bf16(256, 256, 0, 0), idesc128 = umma_idesc_bf16(256, 128, 0, 0);
      // shared-memory descriptors (128B swizzle, LBO 16, SBO 1024): only the start-address field [0, 14) of the low
      // word changes, so the issue loop adds to a 32-bit word instead of rebuilding two 64-bit descriptors per MMA
      const uint64_t desc0 = umma_desc_sw128(0, 16, 1024);
      const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32), desc_lo = static_cast<uint32_t>(desc0);
      auto dsc = [&](uint32_t lo) { return (static_cast<uint64_t>(desc_hi) << 32) | lo; };
      const bool no_mma = (p.debug & 4) != 0;  // timing-only ablation (OMGSR_HALO_DEBUG): no MMAs, results invalid
      uint32_t n0 = 0;
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int g = g_begin; g < g_end;) {
        int b, tx, ya, L; bool dummy;
        segment(g, b, tx, dummy, ya, L);
        for (int i = 0; i < L + 2; ++i) {
          if (i < L) {  // output i starts in this step: its slot must have been drained (and zeroed)
            const uint32_t n = n0 + i;
            mbar_wait(&acc_empty[n & 3], ((n >> 2) & 1) ^ 1);
          }
          const RrPlan pl = rr_plan(i, L, n0);
          const uint32_t d0 = tmem_base + pl.slot[0] * BN, id0 = pl.pair[0] ? idesc256 : idesc128;
          const bool two = pl.n_ops > 1;
          const uint32_t d1 = tmem_base + pl.slot[1] * BN, id1 = pl.pair[1] ? idesc256 : idesc128;
          for (int c = 0; c < p.chunks; ++c) {
            if (XF) mbar_wait(&a_ready[as], aph);
            else mbar_wait(&a_full[as], aph);
            uint32_t a_lo = desc_lo | (smem_u32(sA + as * C::kRowBytes) >> 4);
#pragma unroll 1
            for (int sx = 0; sx < 3; ++sx) {  // tap column: the A view shifts by one pixel (128 B)
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              const uint32_t b_base = smem_u32(sB + bs * C::kBSlotBytes);
              const uint32_t b0_lo = desc_lo | ((b_base + pl.boff[0]) >> 4), b1_lo = desc_lo | ((b_base + pl.boff[1]) >> 4);
              if (leader) {
                if (!no_mma) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    tc_mma2_ss(d0, dsc(a_lo + 2 * k), dsc(b0_lo + 2 * k), id0, 1u);
                    if (two) tc_mma2_ss(d1, dsc(a_lo + 2 * k), dsc(b1_lo + 2 * k), id1, 1u);
                  }
                }
                tc_commit2_mc(&b_empty[bs]);
              }
              a_lo += 8;
              if (++bs == NB) { bs = 0; bph ^= 1; }
            }
            if (leader) tc_commit2_mc(&a_empty[as]);  // the row is free once everything issued so far retires
            if (++as == NA) { as = 0; aph ^= 1; }
          }
          if (p.res_tap && i >= 1 && i <= L) {
            // residual of output i - 1 (its accumulator is live from step i - 1 to step i + 1): D += R[row, shifted by
            // pad_l pixels] x I^T, N = 128, exact (1.0 x bf16 in an fp32 accumulator)
            const uint32_t dr = tmem_base + ((n0 + i - 1) & 3u) * BN;
            for (int c = 0; c < 2; ++c) {
              if (XF) mbar_wait(&a_ready[as], aph);
              else mbar_wait(&a_full[as], aph);
              mbar_wait(&b_full[bs], bph);
              tc_fence_after();
              const uint32_t a_lo = desc_lo | ((smem_u32(sA + as * C::kRowBytes) + static_cast<uint32_t>(p.pad_l * 128)) >> 4);
              const uint32_t b_lo = desc_lo | (smem_u32(sB + bs * C::kBSlotBytes) >> 4);
              if (leader) {
                if (!no_mma) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) tc_mma2_ss(dr, dsc(a_lo + 2 * k), dsc(b_lo + 2 * k), idesc128, 1u);
                }
                tc_commit2_mc(&b_empty[bs]);
                tc_commit2_mc(&a_empty[as]);
              }
              if (++bs == NB) { bs = 0; bph ^= 1; }
              if (++as == NA) { as = 0; aph ^= 1; }
            }
          }
          if (i >= 2 && leader) tc_commit2_mc(&acc_full[(n0 + i - 2) & 3]);  // output i - 2 has received its three tap rows
        }
        g += L; n0 += L;
      }
    }
  }
  } else if (XF && warp >= 12) {
    setmaxnreg_dec<112>();
    // ================= input transform (warps 12..15): GroupNorm + SiLU on each halo row, in place =================
    // One warp per scheduler, so the schedule inside the warp matters (the first version ran at 0.15 IPC, fully
    // busy, and set the pace of the kernel: ncu profiles/r02/ncu_r02_halo_rr_gn_lean.txt). Per unit a thread owns
    // 9 x 16 bytes (8 channels of pixels p0, p0 + 16, ...): fixed swizzle term, so the addresses are one base plus
    // immediates; three 16-byte units at a time go through three PHASES (24 affine results, 24 tanh, 24 outputs) so
    // that the MUFU latencies overlap; the coefficients of the image come from a 1 KB table in shared memory.
    const int tid = threadIdx.x - 384;  // 0..127
    const int j = tid & 7;              // 16-byte unit (8 channels) of a 128-byte pixel row: fixed per thread
    const int p0 = tid >> 3;            // first pixel of this thread in a row; then every 16th (<= 9 per row)
    const uint32_t off0 = static_cast<uint32_t>(p0 * 128 + ((j ^ (p0 & 7)) << 4));
    float* coef_tab = gn_acc + 8 * 64 + 8 * BN;  // scale / 2 [512] then shift / 2 [512] (separate arrays: the packed
    const uint32_t coef_s = smem_u32(coef_tab);  // fp32 FMAs want (scale, scale) and (shift, shift) register pairs)
    int slot = 0;
    uint32_t ph = 0;
    int coef_b = -1;
    for (int g = g_begin; g < g_end;) {
      int b, tx, ya, L; bool dummy;
      segment(g, b, tx, dummy, ya, L);
      const int x0 = tx * 128 - p.pad_l;
      uint32_t onmask = 0;  // this thread's pixels that lie inside the image (the others are conv padding: stay zero)
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        const int pi = p0 + 16 * q, xx = x0 + pi;
        if (pi < p.pw && xx >= 0 && xx < p.IW) onmask |= 1u << q;
      }
      if (!dummy && b != coef_b) {  // (warp-uniform per CTA: all four transform warps take this branch together)
        named_bar_sync(3, 128);  // nobody still reads the previous image's table
        for (int ch = tid; ch < p.Cin; ch += 128) {
          const float2 v = __ldg(p.gn_in + static_cast<long long>(b) * p.Cin + ch);
          // halved: silu(a x + b) = h + h tanh(h) with h = (a / 2) x + b / 2
          coef_tab[ch] = 0.5f * v.x;
          coef_tab[512 + ch] = 0.5f * v.y;
        }
        named_bar_sync(3, 128);
        coef_b = b;
      }
      for (int i = 0; i < L + 2; ++i) {
        const int yy = ya - p.pad_t + i;
        const bool live = !dummy && yy >= 0 && yy < p.IH;
        for (int c = 0; c < p.chunks; ++c) {
          float ca[8], cb[8];
          if (live) {
            const uint32_t ct = coef_s + static_cast<uint32_t>((c * 64 + j * 8) * 4);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const uint4 va = lds128(ct + q * 16), vb = lds128(ct + 2048 + q * 16);
              ca[4 * q] = __uint_as_float(va.x); ca[4 * q + 1] = __uint_as_float(va.y);
              ca[4 * q + 2] = __uint_as_float(va.z); ca[4 * q + 3] = __uint_as_float(va.w);
              cb[4 * q] = __uint_as_float(vb.x); cb[4 * q + 1] = __uint_as_float(vb.y);
              cb[4 * q + 2] = __uint_as_float(vb.z); cb[4 * q + 3] = __uint_as_float(vb.w);
            }
          }
          mbar_wait_relaxed(&a_full[slot], ph, 32);
          if (live) {
            const uint32_t row_s = smem_u32(sA + slot * C::kRowBytes) + off0;
#pragma unroll
            for (int qb = 0; qb < 9; qb += 3) {
              uint4 v[3];
              float h[24], th[24];
#pragma unroll
              for (int t = 0; t < 3; ++t) v[t] = lds128(row_s + (qb + t) * 2048);  // (past the row for q = 8: harmless)
#pragma unroll
              for (int t = 0; t < 3; ++t) {
                const float2 f0 = unpack_bf16x2(v[t].x), f1 = unpack_bf16x2(v[t].y), f2 = unpack_bf16x2(v[t].z),
                             f3 = unpack_bf16x2(v[t].w);
                ffma2p(h[t * 8 + 0], h[t * 8 + 1], f0.x, f0.y, ca[0], ca[1], cb[0], cb[1]);
                ffma2p(h[t * 8 + 2], h[t * 8 + 3], f1.x, f1.y, ca[2], ca[3], cb[2], cb[3]);
                ffma2p(h[t * 8 + 4], h[t * 8 + 5], f2.x, f2.y, ca[4], ca[5], cb[4], cb[5]);
                ffma2p(h[t * 8 + 6], h[t * 8 + 7], f3.x, f3.y, ca[6], ca[7], cb[6], cb[7]);
              }
#pragma unroll
              for (int e = 0; e < 24; ++e) th[e] = tanh_approx(h[e]);
#pragma unroll
              for (int e = 0; e < 24; e += 2) ffma2p(h[e], h[e + 1], h[e], h[e + 1], th[e], th[e + 1], h[e], h[e + 1]);
#pragma unroll
              for (int t = 0; t < 3; ++t) {
                uint4 o;
                o.x = pack_bf16x2(h[t * 8 + 0], h[t * 8 + 1]); o.y = pack_bf16x2(h[t * 8 + 2], h[t * 8 + 3]);
                o.z = pack_bf16x2(h[t * 8 + 4], h[t * 8 + 5]); o.w = pack_bf16x2(h[t * 8 + 6], h[t * 8 + 7]);
                if ((onmask >> (qb + t)) & 1u) sts128(row_s + (qb + t) * 2048, o);
              }
            }
          }
          fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(&a_ready[slot]); else mbar_arrive_remote_relaxed(&a_ready[slot], 0);
          }
          if (++slot == NA) { slot = 0; ph ^= 1; }
        }
        if (p.res_tap && i >= 1 && i <= L) {  // residual rows are not normalised: hand them on as they are
          for (int c = 0; c < 2; ++c) {
            mbar_wait_relaxed(&a_full[slot], ph, 32);
            __syncwarp();
            if (lane == 0) {
              if (rank == 0) mbar_arrive(&a_ready[slot]); else mbar_arrive_remote_relaxed(&a_ready[slot], 0);
            }
            if (++slot == NA) { slot = 0; ph ^= 1; }
          }
        }
      }
      g += L;
    }
  } else {
    // ================= epilogue: warps 4..7 drain columns 0..63 of every output row, warps 8..11 columns 64..127 ====
    if (XF) setmaxnreg_inc<168>();
    const int half = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int ew = half * 4 + quarter;    // epilogue warp index inside the CTA: statistics slot / bias slice
    const int row = quarter * 32 + lane;
    const int c_off = half * 64;
    float* my_gn = gn_acc + ew * 64;
    float* my_bias = gn_acc + 8 * 64 + ew * BN;
    int bias_tile = -1;
    const int gn_n = p.gn_groups * 2;
    int gn_b = -1;
    if (p.gn_sums) { my_gn[lane] = 0.f; my_gn[lane + 32] = 0.f; __syncwarp(); }
    const uint32_t zeros[32] = {};
    uint32_t n = 0;
    for (int g = g_begin; g < g_end;) {
      int b, tx, ya, L; bool dummy;
      segment(g, b, tx, dummy, ya, L);
      if (p.gn_sums && !dummy && b != gn_b) {
        if (gn_b >= 0) gn_flush(my_gn, gn_n, gn_slot(p.gn_sums, gn_b, gn_n, blockIdx.x * 8 + ew, p.gn_slots), lane);
        gn_b = b;
      }
      const int ch0 = cur_n_tile * BN + c_off;  // first output channel of this warp's 64 columns
      if (p.bias && cur_n_tile != bias_tile) {
        stage_bias<BN>(my_bias, p.bias, cur_n_tile * BN, p.Cout, lane);
        bias_tile = cur_n_tile;
      }
      const int x = tx * 128 + row;
      for (int o = 0; o < L; ++o, ++n) {
        const int y = ya + o;
        EpiArgs e;
        e.bias = p.bias; e.gate_row = nullptr; e.out = p.out;
        e.Cout = p.Cout; e.cout_out = p.Cout; e.act = p.act; e.act_c0 = 0; e.act_c1 = 0x7fffffff;
        e.out_f32 = p.out_f32; e.vec32 = p.out_vec32; e.res_vec32 = p.res_vec32; e.res_f32 = 0;
        float gn_part[16];
        e.gn_part = p.gn_sums ? gn_part : nullptr;
        const bool valid = !dummy && (x < p.W) && !(p.debug & 1);
        uint4 rr0[4] = {}, rr1[4] = {};
        e.res_row = (p.res && !p.res_tap && !dummy) ? p.res + (b * p.r_sb + y * p.r_sy + x * p.r_sx) : nullptr;
        if (e.res_row) {  // this thread's 64 residual channels are in flight before the accumulator wait
          load_res32(rr0, e, ch0, valid);
          load_res32(rr1, e, ch0 + 32, valid);
          // the residual comes from DRAM: pull the NEXT row's 128-byte line of this thread into L2 now (the first use of
          // rr0 was the largest long-scoreboard stall of the epilogue, ncu profiles/r02/ncu_r02_halo_rr_gn_res_final.txt)
          if (valid && o + 1 < L) asm volatile("prefetch.global.L2 [%0];" ::"l"(e.res_row + p.r_sy + ch0));
        }
        e.o_off = b * p.o_sb + y * p.o_sy + x * p.o_sx;
        const int slot = n & 3;
        mbar_wait_relaxed(&acc_full[slot], (n >> 2) & 1);
        tc_fence_after();
        const uint32_t t_row = tmem_base + slot * BN + c_off + (static_cast<uint32_t>(quarter * 32) << 16);
        // the slot is needed again ONE row-step after it completes: copy the 64 columns out, zero them and hand the
        // slot back before any arithmetic (the bias / residual / store / statistics work then overlaps the next MMAs)
        uint32_t v0[32], v1[32];
        tmem_ld32(t_row, v0);
        tmem_ld32(t_row + 32, v1);
        tmem_wait_ld();
        tmem_st32(t_row, zeros);
        tmem_st32(t_row + 32, zeros);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(&acc_empty[slot]);
          else mbar_arrive_remote(&acc_empty[slot], 0);
        }
        if (!(p.debug & 2)) {  // (timing-only ablation: no epilogue)
          epilogue_regs64<MODE>(v0, v1, ch0, ch0, valid, e, p.bias ? my_bias + c_off : nullptr, rr0, rr1, my_gn,
                                p.gn_cpg, lane);
        }
      }
      g += L;
    }
    if (p.gn_sums && gn_b >= 0) gn_flush(my_gn, gn_n, gn_slot(p.gn_sums, gn_b, gn_n, blockIdx.x * 8 + ew, p.gn_slots), lane);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// conv_halo_rr_kernel where it applies (3x3, a long enough share of output rows per cluster): 1 every eligible
// convolution (default: measured faster in all six epilogue / transform variants, profiles/r02/halo_rolling_ab.txt),
// 0 never, -1 only the fused input GroupNorm variants. OMGSR_HALO_ROLLING / omgsr_set_halo_rolling.
static const int g_halo_rolling_default = getenv("OMGSR_HALO_ROLLING") ? atoi(getenv("OMGSR_HALO_ROLLING")) : 1;
int g_halo_rolling = g_halo_rolling_default;
int halo_rolling_default() { return g_halo_rolling_default; }

// 128 x 128 bf16 identity per device: the weight matrix of the residual tap
static const void* identity128(cudaStream_t st) {
  static void* ptr[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return nullptr;
  if (!ptr[dev]) {
    // first use on this device allocates and copies: not allowed while `st` is being captured into a CUDA graph (the
    // caller then keeps the residual add in the epilogue; every later, un-captured call installs the buffer)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return nullptr;
    std::vector<uint16_t> h(128 * 128, 0);
    for (int i = 0; i < 128; ++i) h[i * 128 + i] = 0x3f80;  // bf16 1.0
    void* d = nullptr;
    if (cudaMalloc(&d, h.size() * 2) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
    ptr[dev] = d;
  }
  return ptr[dev];
}

template <bool XF, int MODE>
static int launch_halo_rr_m(const CUtensorMap& tmA, const CUtensorMap& tmB128, const CUtensorMap& tmB64,
                            const CUtensorMap& tmR, const CUtensorMap& tmI64, const HaloParams& p, int clusters,
                            cudaStream_t st) {
  using C = HaloRrCfg;
  static unsigned long long attr_done = 0;  // one bit per device
  if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(&conv_halo_rr_kernel<XF, MODE>), C::kSmemBytes, &attr_done,
                                "conv_halo_rr")) return rc;
  g_trace_variant = XF ? "conv_halo_rr_kernel<1>" : "conv_halo_rr_kernel<0>";
  launch_pdl(conv_halo_rr_kernel<XF, MODE>, dim3(2 * clusters), dim3(XF ? 512 : 384), C::kSmemBytes, st, tmA, tmB128, tmB64,
             tmR, tmI64, p);
  return check_launch("conv_halo_rr_kernel");
}
template <bool XF>
static int launch_halo_rr(const CUtensorMap& tmA, const CUtensorMap& tmB128, const CUtensorMap& tmB64,
                          const CUtensorMap& tmR, const CUtensorMap& tmI64, const HaloParams& p, int clusters,
                          cudaStream_t st) {
  // with the residual folded into the MMAs (p.res_tap) the epilogue is the one without residual
  const bool epi_res = p.res != nullptr && !p.res_tap;
  switch (epilogue_mode(p.act, false, epi_res, p.out_f32 != 0, p.gn_sums != nullptr)) {
    case EPI_PLAIN: return launch_halo_rr_m<XF, EPI_PLAIN>(tmA, tmB128, tmB64, tmR, tmI64, p, clusters, st);
    case EPI_RES: return launch_halo_rr_m<XF, EPI_RES>(tmA, tmB128, tmB64, tmR, tmI64, p, clusters, st);
    case EPI_GN: return launch_halo_rr_m<XF, EPI_GN>(tmA, tmB128, tmB64, tmR, tmI64, p, clusters, st);
    case EPI_RES_GN: return launch_halo_rr_m<XF, EPI_RES_GN>(tmA, tmB128, tmB64, tmR, tmI64, p, clusters, st);
    default: return launch_halo_rr_m<XF, EPI_GENERIC>(tmA, tmB128, tmB64, tmR, tmI64, p, clusters, st);
  }
}

// 0: conv_halo_tc2_kernel (every instruction N = 128), 1: conv_halo_rs_kernel (row-shared N = 256 instructions)
int g_halo_row_shared = 1;

template <bool XF>
static int launch_halo_rs(const CUtensorMap& tmA, const CUtensorMap& tmB128, const CUtensorMap& tmB64,
                          const HaloParams& p, cudaStream_t st) {
  using C = HaloRsCfg;
  static unsigned long long attr_done = 0;  // one bit per device
  if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(&conv_halo_rs_kernel<XF>), C::kSmemBytes, &attr_done,
                                "conv_halo_rs")) return rc;
  const long long halo_tiles = static_cast<long long>(p.tiles_x) * p.tiles_y * p.B;
  const long long pairs = (halo_tiles + 1) / 2;
  const int max_clusters = sm_count() / 2;
  const int clusters = pairs < max_clusters ? static_cast<int>(pairs) : max_clusters;
  g_trace_variant = XF ? "conv_halo_rs_kernel<1>" : "conv_halo_rs_kernel<0>";
  launch_pdl(conv_halo_rs_kernel<XF>, dim3(2 * clusters), dim3(XF ? 480 : 352), C::kSmemBytes, st, tmA, tmB128, tmB64, p);
  return check_launch("conv_halo_rs_kernel");
}

int g_halo_2cta = 1;

template <int TM, bool XF>
static int launch_halo2(const CUtensorMap& tmA, const CUtensorMap& tmB, const HaloParams& p, cudaStream_t st) {
  using C = Halo2Cfg<TM>;
  static unsigned long long attr_done = 0;  // one bit per device
  if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(&conv_halo_tc2_kernel<TM, XF>), C::kSmemBytes, &attr_done,
                                "conv_halo_tc2")) return rc;
  const long long halo_tiles = static_cast<long long>(p.tiles_x) * p.tiles_y * p.B;
  const long long pairs = (halo_tiles + 1) / 2;
  const int max_clusters = sm_count() / 2;
  const int clusters = pairs < max_clusters ? static_cast<int>(pairs) : max_clusters;
  g_trace_variant = XF ? "conv_halo_tc2_kernel<2,true>" : "conv_halo_tc2_kernel<2,false>";
  launch_pdl(conv_halo_tc2_kernel<TM, XF>, dim3(2 * clusters), dim3(XF ? 320 : 192), C::kSmemBytes, st, tmA, tmB, p);
  return check_launch("conv_halo_tc2_kernel");
}

template <int BN, int TM>
static int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const HaloParams& p, cudaStream_t st) {
  using C = HaloCfg<BN, TM>;
  static unsigned long long attr_done = 0;  // one bit per device
  if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(&conv_halo_tc_kernel<BN, TM>), C::kSmemBytes, &attr_done,
                                "conv_halo_tc")) return rc;
  const long long total = static_cast<long long>(p.tiles_x) * p.tiles_y * p.B * p.n_tiles;
  const int grid = total < sm_count() ? static_cast<int>(total) : sm_count();
  g_trace_variant = BN == 16 ? "conv_halo_tc_kernel<16,2>" : "conv_halo_tc_kernel<128,2>";
  launch_pdl(conv_halo_tc_kernel<BN, TM>, dim3(grid), dim3(192), C::kSmemBytes, st, tmA, tmB, p);
  return check_launch("conv_halo_tc_kernel");
}

// does the rolling-row kernel take this 3x3 problem? (a long enough share of pair-rows per cluster; the input GroupNorm
// coefficient table in shared memory holds 512 channels; at most two 128-wide output-channel passes)
static bool halo_rr_fits(int B, int OH, int OW, int Cin, int Cout) {
  const int tiles_x = (OW + 127) / 128, nt = (Cout + 127) / 128;
  const long long pair_rows = static_cast<long long>(B) * ((tiles_x + 1) / 2) * OH * nt;
  return nt <= 2 && Cin <= 512 && Cin % 64 == 0 && pair_rows >= static_cast<long long>(sm_count() / 2) * 48 &&
         pair_rows < 0x7fffffffLL;
}

// Returns 1 if this problem is not eligible (caller falls through to the generic kernel), 0 on success, <0 on error.
int try_conv_halo(const void* x, const void* w, const float* bias, const void* res, void* y, int B, int IH, int IW,
                  int OH, int OW, int Cin, int Cout, int R, int S, int pad_t, int pad_l, long long x_sx, long long x_sy,
                  long long x_sb, long long o_sx, long long o_sy, long long o_sb, long long r_sx, long long r_sy,
                  long long r_sb, long long w_ld, int act, int out_f32, float* gn_sums, int gn_groups, const float* gn_in,
                  cudaStream_t st) {
  if (R * S < 4 || R > 3 || S > 3) return 1;
  if (OW < 96) return 1;                        // a 128-pixel row tile would be mostly padding
  // Cout = 128, or a thin Cout < 128 served by one 128-wide tile whose missing weight rows TMA zero-fills and whose
  // stores are masked (decoder conv_out 128 -> 3: 1.1 ms instead of 2.3 ms on the generic 64-wide path)
  if (Cin % 64 || (Cout > 128 && Cout % 128)) return 1;
  if (Cout < 128 && gn_sums) return 1;
  if ((Cout % 8) && o_sx < ((Cout + 7) & ~7)) return 1;  // the 8-channel store granule needs the room
  if (act == OMGSR_ACT_GEGLU) return 1;
  // measured on B200 (profiles/r01/conv_gemm_timing_v2.txt): the halo kernel wins where the generic kernel is
  // bound by its L2->SMEM operand stream, i.e. Cout = 128 (128-wide tiles re-read A per 128 output channels):
  // 1150 vs 918 TFLOP/s at 128->128 @512^2. With 256-wide tiles (no room for a second TMEM accumulator set at
  // TM=2) the generic kernel is faster (1333 vs 1209 at 256->256 @256^2), so those stay on the generic path.
  // Exception (round 2): with the GroupNorm + SiLU of the input fused in, Cout = 256 runs as TWO 128-channel passes of
  // the rolling-row kernel over the same input rows (0.85 ms at 256->256 @256^2 x16 against 0.80 ms generic + 0.19 ms for
  // the separate normalisation pass it replaces).
  const bool wide_rr = Cout == 256 && gn_in != nullptr && R == 3 && S == 3 && OW >= 256 && g_halo_rolling != 0 &&
                       halo_rr_fits(B, OH, OW, Cin, Cout);
  if (Cout % 256 == 0 && !wide_rr) return 1;
  const int bn = 128;
  constexpr int TM = 2;
  HaloParams p;
  p.tiles_x = (OW + 127) / 128; p.tiles_y = (OH + TM - 1) / TM; p.B = B; p.W = OW; p.H = OH;
  p.n_tiles = (Cout + bn - 1) / bn; p.Cout = Cout; p.chunks = Cin / 64; p.R = R; p.S = S; p.pad_t = pad_t; p.pad_l = pad_l;
  p.Cin = Cin; p.pw = 128 + S - 1;
  p.a_bytes = 128 * p.pw * (TM + R - 1);
  { static const int dbg = getenv("OMGSR_HALO_DEBUG") ? atoi(getenv("OMGSR_HALO_DEBUG")) : 0; p.debug = dbg; }
  p.gn_sums = nullptr; p.gn_cpg = 0; p.gn_groups = 0;
  p.gn_in = reinterpret_cast<const float2*>(gn_in); p.IH = IH; p.IW = IW; p.res_tap = 0;
  if (gn_sums) {
    const int cpg = gn_groups > 0 ? Cout / gn_groups : 0;
    if (out_f32 || gn_groups <= 0 || gn_groups > 32 || Cout % gn_groups || !(cpg == 4 || cpg == 8 || cpg == 16)) return 1;
    p.gn_sums = gn_sums; p.gn_cpg = cpg; p.gn_groups = gn_groups; p.gn_slots = gn_slots_per_image();
  }
  p.o_sx = o_sx; p.o_sy = o_sy; p.o_sb = o_sb; p.r_sx = r_sx; p.r_sy = r_sy; p.r_sb = r_sb;
  p.out = y; p.bias = bias; p.res = static_cast<const bf16*>(res); p.act = act; p.out_f32 = out_f32;
  p.out_vec32 = !out_f32 && (reinterpret_cast<uintptr_t>(y) % 32 == 0) && ((o_sx | o_sy | o_sb) % 16 == 0);
  p.res_vec32 = res && (reinterpret_cast<uintptr_t>(res) % 32 == 0) && ((r_sx | r_sy | r_sb) % 16 == 0);
  if (static_cast<long long>(p.tiles_x) * p.tiles_y * p.B * p.n_tiles > 0x7fffffffLL) return 1;
  CUtensorMap tmA, tmB;
  const uint64_t eb = 2;
  uint64_t dims[5] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(IW), static_cast<uint64_t>(IH),
                      static_cast<uint64_t>(B), 1};
  uint64_t strides[4] = {x_sx * eb, x_sy * eb, x_sb * eb, x_sb * eb * B};
  uint32_t box[5] = {64, static_cast<uint32_t>(p.pw), static_cast<uint32_t>(TM + R - 1), 1, 1};
  int rc = make_tmap_bf16(&tmA, x, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  uint64_t wd[4] = {static_cast<uint64_t>(w_ld), static_cast<uint64_t>(Cout), 1, 1};
  uint64_t ws[3] = {w_ld * eb, w_ld * eb * Cout, w_ld * eb * Cout};
  // thin Cout (decoder conv_out 128 -> 3): N = 16 MMAs, bound by the 4 KB A read per MMA (32 instead of 64 cycles)
  const bool thin = Cout <= 16;
  const bool pair = !thin && g_halo_2cta && p.n_tiles == 1 && static_cast<long long>(p.tiles_x) * p.tiles_y * p.B >= 2;
  uint32_t wb[4] = {64, static_cast<uint32_t>(thin ? 16 : (pair ? 64 : bn)), 1, 1};
  rc = make_tmap_bf16(&tmB, w, 4, wd, ws, wb, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  if (thin) return launch_halo<16, TM>(tmA, tmB, p, st);
  if (gn_in && !pair && !wide_rr)
    return set_error(OMGSR_ERR_SHAPE, "conv_halo: the fused input GroupNorm needs the CTA-pair halo kernel (16 < Cout <= 128, >= 2 tiles; or Cout = 256 on the rolling-row kernel)");
  static const bool rs_off = getenv("OMGSR_HALO_ROW_SHARED") && atoi(getenv("OMGSR_HALO_ROW_SHARED")) == 0;  // A/B timing
  if ((pair || wide_rr) && g_halo_row_shared && !rs_off && R >= 2) {
    // row-shared schedule: one input ROW per A load, weight units of 128 rows (N = 256 instructions) or 64 rows
    CUtensorMap tmRow, tmB128, tmB64;
    uint32_t rbox[5] = {64, static_cast<uint32_t>(p.pw), 1, 1, 1};
    rc = make_tmap_bf16(&tmRow, x, 5, dims, strides, rbox, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    uint32_t wb128[4] = {64, 128, 1, 1}, wb64[4] = {64, 64, 1, 1};
    rc = make_tmap_bf16(&tmB128, w, 4, wd, ws, wb128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmB64, w, 4, wd, ws, wb64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    // rolling rows: every input row loaded / transformed once; worth it when a cluster's share of output rows is long
    // against the two extra input rows per segment
    const int max_clusters = sm_count() / 2;
    const bool rr_fit = R == 3 && S == 3 && halo_rr_fits(B, OH, OW, Cin, Cout);
    const bool rr_on = g_halo_rolling >= 0 ? g_halo_rolling != 0 : gn_in != nullptr;
    if (rr_fit && rr_on) {
      // residual as an identity tap (Cout = 128, bf16 residual with TMA-compatible strides): OMGSR_HALO_RES_TAP=0 keeps
      // the residual add in the epilogue (A/B timing)
      static const bool tap_off = getenv("OMGSR_HALO_RES_TAP") && atoi(getenv("OMGSR_HALO_RES_TAP")) == 0;
      CUtensorMap tmR = tmRow, tmI64 = tmB64;
      HaloParams q = p;
      q.res_tap = 0;
      if (res && !tap_off && Cout == 128 && !out_f32 && r_sx >= 128 && (r_sx % 8) == 0 && (r_sy % 8) == 0 &&
          (r_sb % 8) == 0 && (reinterpret_cast<uintptr_t>(res) % 16) == 0) {
        const void* eye = identity128(st);
        if (eye) {
          uint64_t rdims[5] = {128, static_cast<uint64_t>(OW), static_cast<uint64_t>(OH), static_cast<uint64_t>(B), 1};
          uint64_t rstr[4] = {static_cast<uint64_t>(r_sx) * eb, static_cast<uint64_t>(r_sy) * eb,
                              static_cast<uint64_t>(r_sb) * eb, static_cast<uint64_t>(r_sb) * eb * B};
          rc = make_tmap_bf16(&tmR, res, 5, rdims, rstr, rbox, CU_TENSOR_MAP_SWIZZLE_128B);
          if (rc) return rc;
          uint64_t id[4] = {128, 128, 1, 1};
          uint64_t is[3] = {128 * eb, 128 * 128 * eb, 128 * 128 * eb};
          rc = make_tmap_bf16(&tmI64, eye, 4, id, is, wb64, CU_TENSOR_MAP_SWIZZLE_128B);
          if (rc) return rc;
          q.res_tap = 1;
        }
      }
      if (gn_in) return launch_halo_rr<true>(tmRow, tmB128, tmB64, tmR, tmI64, q, max_clusters, st);
      return launch_halo_rr<false>(tmRow, tmB128, tmB64, tmR, tmI64, q, max_clusters, st);
    }
    if (wide_rr) return set_error(OMGSR_ERR_SHAPE, "conv_halo: Cout = 256 with a fused input GroupNorm needs the rolling-row kernel");
    if (gn_in) return launch_halo_rs<true>(tmRow, tmB128, tmB64, p, st);
    return launch_halo_rs<false>(tmRow, tmB128, tmB64, p, st);
  }
  if (gn_in) return launch_halo2<TM, true>(tmA, tmB, p, st);
  if (pair) return launch_halo2<TM, false>(tmA, tmB, p, st);
  return launch_halo<128, TM>(tmA, tmB, p, st);
}

}  // namespace omgsr

extern "C" int omgsr_set_halo_rolling(int mode) {
  omgsr::g_halo_rolling = mode < -1 ? omgsr::halo_rolling_default() : mode;  // < -1: back to the start-up default
  return OMGSR_OK;
}

/* can omgsr_conv2d_gn_nhwc_bf16 take a 3x3 / s1 / p1 convolution with 256 output channels (two passes of the rolling-row
 * kernel)? The engines ask before they defer a GroupNorm to the convolution that follows it. */
extern "C" int omgsr_conv_gn_wide_fusable(int B, int OH, int OW, int Cin, int Cout) {
  return (Cout == 256 && OW >= 256 && omgsr::g_halo_rolling != 0 && omgsr::g_halo_row_shared != 0 &&
          omgsr::halo_rr_fits(B, OH, OW, Cin, Cout)) ? 1 : 0;
}
