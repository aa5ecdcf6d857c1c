// Fused flash attention on tcgen05: out = softmax(scale * Q K^T) V for one (batch, head) per CTA row-tile.
//
//   * Q/K/V are [B, T, H, D] bf16 views (D contiguous, arbitrary row/head/batch strides), D in {64, 128}.
//   * CTA = 128 query rows. warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = softmax/accumulate
//     (one query row per thread, so row max / row sum need no shuffles).
//   * per KV block: S = Q K^T into TMEM (tcgen05.mma, K-major operands) -> threads read S with tcgen05.ld,
//     online softmax in fp32 (exp2 with the scale folded in), P written to shared memory as bf16 in the
//     128B-swizzled K-major layout -> O_blk = P V into TMEM (V consumed in place as an MN-major operand)
//     -> threads fold O_blk into register accumulators with the running rescale.
//   * TMEM: S (BLK_N columns) + O_blk (D columns) <= 256 columns and <= 112 KB of shared memory, so two
//     CTAs are co-resident per SM: while one CTA runs its softmax the other one owns the tensor pipe.
//
// Replaces F.scaled_dot_product_attention under diffusers' Attention processors for the UNet self/cross
// attention (d=64) and the FLUX joint attention (d=128) that the reference reaches from
// infer/omgsr_s_infer_model.py:75-79 and infer/omgsr_f_infer_model.py:191-200.
#include "common.cuh"
#include "../../include/omgsr_b200.h"

namespace omgsr {

template <int D>
struct AttnCfg {
  static constexpr int kBM = 128;
  static constexpr int kBN = (D == 64) ? 128 : 64;
  static constexpr int kStages = 2;
  static constexpr int kQBytes = kBM * D * 2;
  static constexpr int kKBytes = kBN * D * 2;
  static constexpr int kVBytes = kBN * D * 2;
  static constexpr int kPBytes = kBM * kBN * 2;
  static constexpr int kSmemBytes = kQBytes + kStages * (kKBytes + kVBytes) + kPBytes + 1024 + 256;
  static constexpr int kTmemCols = 256;
  static constexpr int kSCol = 0;
  static constexpr int kOCol = 128;
};

struct AttnParams {
  int M, N, H, B;
  long long o_row, o_head, o_batch;
  bf16* out;
  float scale_log2;
};

template <int D>
__global__ void __launch_bounds__(192, (D == 64) ? 2 : 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using C = AttnCfg<D>;
  constexpr int BN = C::kBN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // offset, not a cast: keeps the shared address space (LDS / STS)
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + C::kQBytes;  // [stage][K | V]
  uint8_t* sP = sKV + C::kStages * (C::kKBytes + C::kVBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + C::kPBytes);
  uint64_t* q_full = bars;                      // 1
  uint64_t* k_full = bars + 1;                  // [kStages]
  uint64_t* v_full = bars + 1 + C::kStages;     // [kStages]
  uint64_t* kv_empty = bars + 1 + 2 * C::kStages;  // [kStages]
  uint64_t* s_full = bars + 1 + 3 * C::kStages;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = s_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 3);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // uniform: see the MMA issuer
  const int m0 = blockIdx.x * C::kBM, h = blockIdx.y, b = blockIdx.z;
  const int nblk = (p.N + BN - 1) / BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_launch_dependents();
  pdl_wait();  // nothing above touches global memory

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer: Q once, then K_j / V_j through a kStages ring ----
      mbar_expect_tx(q_full, C::kQBytes);
#pragma unroll
      for (int pnl = 0; pnl < D / 64; ++pnl) tma_load_4d(sQ + pnl * (C::kBM * 128), &tmQ, q_full, pnl * 64, m0, h, b);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait_relaxed(&kv_empty[stage], phase ^ 1);
        uint8_t* sK = sKV + stage * (C::kKBytes + C::kVBytes);
        uint8_t* sV = sK + C::kKBytes;
        mbar_expect_tx(&k_full[stage], C::kKBytes);
#pragma unroll
        for (int pnl = 0; pnl < D / 64; ++pnl)
          tma_load_4d(sK + pnl * (BN * 128), &tmK, &k_full[stage], pnl * 64, j * BN, h, b);
        mbar_expect_tx(&v_full[stage], C::kVBytes);
#pragma unroll
        for (int pnl = 0; pnl < D / 64; ++pnl)
          tma_load_4d(sV + pnl * (BN * 128), &tmV, &v_full[stage], pnl * 64, j * BN, h, b);
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    {  // whole warp on warp-uniform values, one elected lane issues (gemm_conv_tc.cu has the argument)
      const bool leader = elect_one();
      // ---- MMA issuer ----
      const uint32_t idesc_s = umma_idesc_bf16(128, BN, 0, 0);  // S = Q K^T : both K-major
      const uint32_t idesc_o = umma_idesc_bf16(128, D, 0, 1);   // O = P V   : V is MN-major ([kv, d])
      const uint32_t tS = tmem_base + C::kSCol, tO = tmem_base + C::kOCol;
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      auto issue_s = [&](int stage) {
        const uint32_t aK = smem_u32(sKV + stage * (C::kKBytes + C::kVBytes));
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint64_t da = umma_desc_sw128(aQ + (kk >> 2) * (C::kBM * 128) + (kk & 3) * 32, 16, 1024);
          const uint64_t db = umma_desc_sw128(aK + (kk >> 2) * (BN * 128) + (kk & 3) * 32, 16, 1024);
          if (leader) tc_mma_ss(tS, da, db, idesc_s, kk != 0);
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0);
      if (leader) tc_commit(s_full);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(p_full, j & 1);  // softmax consumed S_j and wrote P_j
        mbar_wait(&v_full[stage], phase);
        tc_fence_after();
        const uint32_t aV = smem_u32(sKV + stage * (C::kKBytes + C::kVBytes) + C::kKBytes);
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
          const uint64_t da = umma_desc_sw128(aP + (kk >> 2) * (C::kBM * 128) + (kk & 3) * 32, 16, 1024);
          // MN-major V: 64-wide d atoms are BN*128 bytes apart (LBO), 8-row kv groups 1024 B apart (SBO)
          const uint64_t db = umma_desc_sw128(aV + kk * 2048, BN * 128, 1024);
          if (leader) tc_mma_ss(tO, da, db, idesc_o, kk != 0);
        }
        if (leader) tc_commit(o_full);
        const int nstage = (stage + 1 == C::kStages) ? 0 : stage + 1;
        const uint32_t nphase = (stage + 1 == C::kStages) ? (phase ^ 1) : phase;
        if (j + 1 < nblk) {
          mbar_wait(&k_full[nstage], nphase);
          tc_fence_after();
          issue_s(nstage);
          if (leader) tc_commit(s_full);
        }
        if (leader) tc_commit(&kv_empty[stage]);  // K_j and V_j are free once everything issued so far retires
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    // ---- softmax + accumulate: one query row per thread ----
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + C::kSCol + lane_off, tO = tmem_base + C::kOCol + lane_off;
    float o_acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* p_row = sP + row * 128;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int n_valid = min(BN, p.N - j * BN);
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tS + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < n_valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f((m_run - m_new) * p.scale_log2);
      const float mb = m_new * p.scale_log2;
      // pass 2: probabilities -> bf16 P tile (128B-swizzled K-major, 64-column panels)
      float lsum = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tS + c0, v);
        tmem_wait_ld();
        float pr[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e = exp2f(__uint_as_float(v[i]) * p.scale_log2 - mb);
          pr[i] = (c0 + i < n_valid) ? e : 0.f;
          lsum += pr[i];
        }
        uint8_t* panel = p_row + (c0 >> 6) * (C::kBM * 128);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((c0 & 63) >> 3) + q;  // 16-byte chunk inside the 128-byte row
          uint4 u;
          u.x = pack_bf16x2(pr[q * 8 + 0], pr[q * 8 + 1]);
          u.y = pack_bf16x2(pr[q * 8 + 2], pr[q * 8 + 3]);
          u.z = pack_bf16x2(pr[q * 8 + 4], pr[q * 8 + 5]);
          u.w = pack_bf16x2(pr[q * 8 + 6], pr[q * 8 + 7]);
          *reinterpret_cast<uint4*>(panel + ((chunk ^ (row & 7)) << 4)) = u;
        }
      }
      l_run = l_run * alpha + lsum;
      m_run = m_new;
      tc_fence_before();
      fence_proxy_async_smem();  // make the generic-proxy P writes visible to the tensor core
      mbar_arrive(p_full);
      // fold O_blk = P_j V_j into the running accumulator
      mbar_wait(o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < D; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tO + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c0 + i] = o_acc[c0 + i] * alpha + __uint_as_float(v[i]);
      }
      tc_fence_before();
    }
    const int m = m0 + row;
    if (m < p.M) {
      const float inv = 1.0f / l_run;
      bf16* o = p.out + static_cast<long long>(b) * p.o_batch + static_cast<long long>(m) * p.o_row +
                static_cast<long long>(h) * p.o_head;
#pragma unroll
      for (int c = 0; c < D; c += 8) {
        uint4 u;
        u.x = pack_bf16x2(o_acc[c + 0] * inv, o_acc[c + 1] * inv);
        u.y = pack_bf16x2(o_acc[c + 2] * inv, o_acc[c + 3] * inv);
        u.z = pack_bf16x2(o_acc[c + 4] * inv, o_acc[c + 5] * inv);
        u.w = pack_bf16x2(o_acc[c + 6] * inv, o_acc[c + 7] * inv);
        *reinterpret_cast<uint4*>(o + c) = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ================================================================================================================
// v2: two query tiles per CTA (ping-pong), O accumulated in TMEM, lazy rescale.
//
//   * CTA = 2 x 128 query rows of one (batch, head): softmax group A = warps 2..5, group B = warps 6..9, warp 0 =
//     TMA producer, warp 1 = MMA issuer (320 threads). K_j / V_j are loaded ONCE per block and used by both groups,
//     and while one group exponentiates, the tensor pipe runs the other group's S / PV MMAs.
//   * O never leaves TMEM during the loop: PV_j accumulates in place. The running row maximum is allowed to go stale:
//     probabilities are formed against m_stale and O is only rescaled (tcgen05.ld -> scale -> tcgen05.st) when a
//     row's new maximum exceeds m_stale by more than 8 in the log2 domain (p <= 256, safe in bf16 / fp32). After
//     the first few blocks this almost never fires, which removes the per-block O read-modify-write of v1.
//   * TMEM: S_a[2] @0/@64, S_b[2] @128/@192 (double-buffered scores), O_a @256, O_b @384 (512 columns, 1 CTA / SM).
// ================================================================================================================
template <int D>
struct Attn2Cfg {
  // 64 keys per block for both head sizes: the whole S row of a thread (64 fp32) is read from TMEM ONCE with one
  // tcgen05.ld.x64 and kept in registers for the max and the exponentiation (the first version re-read S in a second
  // pass and waited after every 32 columns: 30 % tensor-pipe activity in ncu)
  static constexpr int kBN = 64;
  static constexpr int kPBuf = 2;
  static constexpr int kStages = (D == 64) ? 4 : 3;
  static constexpr int kQBytes = 2 * 128 * D * 2;
  static constexpr int kKBytes = kBN * D * 2;
  static constexpr int kStageBytes = 2 * kKBytes;
  static constexpr int kPBytes = 128 * kBN * 2;  // one P buffer of one group
  // d = 64: the row sums l = sum_j p_j come out of the PV MMA itself. V is extended by 16 all-ones columns (a constant
  // 8 KB shared-memory panel reached through the descriptor's leading-dimension offset), so O gets a 65th column that
  // holds exactly the sum of the bf16 probabilities the MMA consumed, is accumulated / rescaled with O, and costs
  // the softmax warps no FADD at all. (d = 128 has no TMEM room for 2 x 144 columns next to the double-buffered S.)
  static constexpr bool kOnesCol = (D == 64);
  static constexpr int kOnesBytes = kOnesCol ? kBN * 128 : 0;
  static constexpr int kSmemBytes = kQBytes + kStages * kStageBytes + 2 * kPBuf * kPBytes + kOnesBytes + 1024 + 256;
};

// (ex2.approx.ftz.bf16x2 was tried for the exponentials: ptxas lowers it to TWO scalar MUFU.EX2.BF16 ops, one per
// half, so it saves no MUFU issue slots and adds pack / permute work.)

template <int D>
__global__ void __launch_bounds__(320, 1)
flash_attn2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using C = Attn2Cfg<D>;
  constexpr int BN = C::kBN, PB = C::kPBuf;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // offset, not a cast: keeps the shared address space (LDS / STS)
  uint8_t* sQ = smem;                                   // [group][panel][128 rows x 128 B]
  uint8_t* sKV = sQ + C::kQBytes;                       // [stage][K | V]
  uint8_t* sP = sKV + C::kStages * C::kStageBytes;      // [group][buf]
  uint8_t* sOnes = sP + 2 * PB * C::kPBytes;            // [64 keys x 128 B] of bf16 1.0 (d = 64 only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + C::kOnesBytes);
  uint64_t* q_full = bars;                   // 1
  uint64_t* k_full = q_full + 1;             // [stages]
  uint64_t* v_full = k_full + C::kStages;    // [stages]
  uint64_t* kv_empty = v_full + C::kStages;  // [stages]
  uint64_t* s_full = kv_empty + C::kStages;  // [2 groups][2 S buffers]
  uint64_t* p_full = s_full + 4;             // [2][PB]
  uint64_t* pv_done = p_full + 2 * PB;       // [2][PB]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2 * PB);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // uniform: see the MMA issuer
  const int m0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int nblk = (p.N + BN - 1) / BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < C::kStages; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g * 2], 1);
      mbar_init(&s_full[g * 2 + 1], 1);
      for (int i = 0; i < PB; ++i) { mbar_init(&p_full[g * PB + i], 128); mbar_init(&pv_done[g * PB + i], 1); }
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (C::kOnesCol) {
    for (int i = threadIdx.x; i < C::kOnesBytes / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_launch_dependents();
  pdl_wait();  // nothing above touches global memory

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, C::kQBytes);
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int pnl = 0; pnl < D / 64; ++pnl)
          tma_load_4d(sQ + (g * (D / 64) + pnl) * (128 * 128), &tmQ, q_full, pnl * 64, m0 + g * 128, h, b);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait_relaxed(&kv_empty[stage], phase ^ 1);
        uint8_t* sK = sKV + stage * C::kStageBytes;
        uint8_t* sV = sK + C::kKBytes;
        mbar_expect_tx(&k_full[stage], C::kKBytes);
#pragma unroll
        for (int pnl = 0; pnl < D / 64; ++pnl)
          tma_load_4d(sK + pnl * (BN * 128), &tmK, &k_full[stage], pnl * 64, j * BN, h, b);
        mbar_expect_tx(&v_full[stage], C::kKBytes);
#pragma unroll
        for (int pnl = 0; pnl < D / 64; ++pnl)
          tma_load_4d(sV + pnl * (BN * 128), &tmV, &v_full[stage], pnl * 64, j * BN, h, b);
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    {  // whole warp on warp-uniform values, one elected lane issues (gemm_conv_tc.cu has the argument)
      const bool leader = elect_one();
      const uint32_t idesc_s = umma_idesc_bf16(128, BN, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, C::kOnesCol ? D + 16 : D, 0, 1);
      // S_g(j) goes to buffer j & 1 of group g: the scores of block j + 1 are issued BEFORE waiting for the
      // probabilities of block j, so the tensor pipe computes them while the group is still exponentiating block j
      // (single S buffer: 403 / 727 TFLOP/s at d = 64 / 128; double-buffered: 448 / 743). Serving the groups in
      // whatever order their P becomes ready (mbarrier.test_wait polling) was tried and is SLOWER (383 / 647): the
      // polling thread steals issue slots from the two softmax warps on its scheduler.
      auto issue_s = [&](int g, int stage, int buf) {
        const uint32_t aQ = smem_u32(sQ + g * (D / 64) * (128 * 128));
        const uint32_t aK = smem_u32(sKV + stage * C::kStageBytes);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint64_t da = umma_desc_sw128(aQ + (kk >> 2) * (128 * 128) + (kk & 3) * 32, 16, 1024);
          const uint64_t db = umma_desc_sw128(aK + (kk >> 2) * (BN * 128) + (kk & 3) * 32, 16, 1024);
          if (leader) tc_mma_ss(tmem_base + g * 128 + buf * BN, da, db, idesc_s, kk != 0);
        }
        if (leader) tc_commit(&s_full[g * 2 + buf]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0, 0);
      issue_s(1, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nblk; ++j) {
        const int nstage = (stage + 1 == C::kStages) ? 0 : stage + 1;
        const uint32_t nphase = (stage + 1 == C::kStages) ? (phase ^ 1) : phase;
        const int pb = j % PB;
        const uint32_t pph = (j / PB) & 1;
        if (j + 1 < nblk) {
          // buffer (j + 1) & 1 was last read by softmax(j - 1), which finished before p_full(j - 1) (waited below,
          // one iteration ago)
          mbar_wait(&k_full[nstage], nphase);
          tc_fence_after();
          issue_s(0, nstage, (j + 1) & 1);
          issue_s(1, nstage, (j + 1) & 1);
        }
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&p_full[g * PB + pb], pph);  // group g consumed S_g(j) and wrote P_g(j)
          if (g == 0) mbar_wait(&v_full[stage], phase);
          tc_fence_after();
          const uint32_t aP = smem_u32(sP + (g * PB + pb) * C::kPBytes);
          const uint32_t aV = smem_u32(sKV + stage * C::kStageBytes + C::kKBytes);
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk) {
            const uint64_t da = umma_desc_sw128(aP + (kk >> 2) * (128 * 128) + (kk & 3) * 32, 16, 1024);
            // second 64-wide N atom: the next d panel (d = 128) or the all-ones panel (d = 64)
            const uint32_t lbo = C::kOnesCol ? smem_u32(sOnes) - aV : BN * 128;
            const uint64_t db = umma_desc_sw128(aV + kk * 2048, lbo, 1024);
            if (leader) tc_mma_ss(tmem_base + 256 + g * 128, da, db, idesc_o, (j | kk) != 0);
          }
          if (leader) tc_commit(&pv_done[g * PB + pb]);
        }
        if (leader) tc_commit(&kv_empty[stage]);
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + g * 128 + lane_off, tO = tmem_base + 256 + g * 128 + lane_off;
    float m_stale = 0.f, l_run = 0.f;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&s_full[g * 2 + (j & 1)], (j >> 1) & 1);
      tc_fence_after();
      const int n_valid = min(BN, p.N - j * BN);
      uint32_t sv[64];
      tmem_ld64(tS + (j & 1) * BN, sv);
      tmem_wait_ld();
      if (n_valid < BN) {  // ragged last block only: mask the columns that do not exist
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= n_valid) sv[i] = 0xff800000u;  // -inf
      }
      float mx0 = __uint_as_float(sv[0]), mx1 = __uint_as_float(sv[1]);
#pragma unroll
      for (int i = 2; i < 64; i += 2) {
        mx0 = fmaxf(mx0, __uint_as_float(sv[i]));
        mx1 = fmaxf(mx1, __uint_as_float(sv[i + 1]));
      }
      const float mx = fmaxf(mx0, mx1);
      const int pb = j % PB;
      if (j == 0) {
        m_stale = mx;
      } else {
        const bool need = (mx - m_stale) * p.scale_log2 > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // every PV issued so far must have retired before O is touched
          mbar_wait(&pv_done[g * PB + ((j - 1) % PB)], ((j - 1) / PB) & 1);
          tc_fence_after();
          const float alpha = need ? exp2f((m_stale - mx) * p.scale_log2) : 1.0f;
#pragma unroll
          for (int c0 = 0; c0 < (C::kOnesCol ? D + 32 : D); c0 += 32) {  // + the row-sum column(s) at D
            uint32_t v[32];
            tmem_ld32(tO + c0, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st32(tO + c0, v);
          }
          tmem_wait_st();
          l_run *= alpha;
          if (need) m_stale = mx;
        }
      }
      // the P buffer we are about to overwrite was last read by PV(j - PB)
      if (j >= PB) mbar_wait(&pv_done[g * PB + pb], ((j - PB) / PB) & 1);
      const float mb = m_stale * p.scale_log2;
      uint8_t* p_row = sP + (g * PB + pb) * C::kPBytes + row * 128;
      float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float pr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pr[i] = exp2f(fmaf(__uint_as_float(sv[q * 8 + i]), p.scale_log2, -mb));  // exp2(-inf) = 0
        if (!C::kOnesCol) {
          ls0 += (pr[0] + pr[1]) + (pr[2] + pr[3]);
          ls1 += (pr[4] + pr[5]) + (pr[6] + pr[7]);
        }
        uint4 u;
        u.x = pack_bf16x2(pr[0], pr[1]); u.y = pack_bf16x2(pr[2], pr[3]);
        u.z = pack_bf16x2(pr[4], pr[5]); u.w = pack_bf16x2(pr[6], pr[7]);
        *reinterpret_cast<uint4*>(p_row + ((q ^ (row & 7)) << 4)) = u;
      }
      const float lsum = ls0 + ls1;
      l_run += lsum;
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&p_full[g * PB + pb]);
    }
    // epilogue: O / l
    mbar_wait(&pv_done[g * PB + ((nblk - 1) % PB)], ((nblk - 1) / PB) & 1);
    tc_fence_after();
    const int m = m0 + g * 128 + row;
    if (C::kOnesCol) {  // l = the ones column of O
      uint32_t lv[16];
      tmem_ld16(tO + D, lv);
      tmem_wait_ld();
      l_run = __uint_as_float(lv[0]);
    }
    const float inv = 1.0f / l_run;
    bf16* o = p.out + static_cast<long long>(b) * p.o_batch + static_cast<long long>(m) * p.o_row +
              static_cast<long long>(h) * p.o_head;
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tO + c0, v);
      tmem_wait_ld();
      if (m < p.M) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]) * inv, __uint_as_float(v[q * 8 + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]) * inv, __uint_as_float(v[q * 8 + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]) * inv, __uint_as_float(v[q * 8 + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]) * inv, __uint_as_float(v[q * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(o + c0 + q * 8) = u;
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// exp2 of a PAIR on the FMA pipe (no MUFU): round-to-nearest split t = r + f with the 1.5 * 2^23 trick, degree-3 minimax
// polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, P is stored as bf16: 3.9e-3), 2^r added into the exponent
// field. 2 FMNMX + 6 FFMA2 + 2 SHL + 2 IADD per pair against 2 MUFU.EX2, which occupy the 4-lane MUFU unit of the
// scheduler for 16 cycles: the softmax of these kernels is MUFU-bound (128 exponentials per row and key block = 2048
// MUFU cycles per 256 x 128 block, as much as its MMAs), so a fraction of the pairs goes this way.
__device__ __forceinline__ void exp2_poly2(float& t0, float& t1) {
  const float magic = 12582912.0f;  // 1.5 * 2^23
  t0 = fmaxf(t0, -125.0f);
  t1 = fmaxf(t1, -125.0f);
  float i0, i1, r0, r1, f0, f1, p0, p1;
  ffma2p(i0, i1, t0, t1, 1.0f, 1.0f, magic, magic);      // low mantissa bits = round(t)
  ffma2p(r0, r1, i0, i1, 1.0f, 1.0f, -magic, -magic);    // round(t) as a float
  ffma2p(f0, f1, r0, r1, -1.0f, -1.0f, t0, t1);          // f = t - round(t)
  ffma2p(p0, p1, f0, f1, 0.0551716648f, 0.0551716648f, 0.2426111251f, 0.2426111251f);
  ffma2p(p0, p1, p0, p1, f0, f1, 0.6932609677f, 0.6932609677f);
  ffma2p(p0, p1, p0, p1, f0, f1, 0.9999280572f, 0.9999280572f);
  t0 = __int_as_float(__float_as_int(p0) + (__float_as_int(i0) << 23));
  t1 = __int_as_float(__float_as_int(p1) + (__float_as_int(i1) << 23));
}

constexpr bool kAttnV5Default = true;   // v5 vs v3, same box: d = 128 0.881 -> 0.853 ms, d = 64 0.534 -> 0.461 ms (profiles/r02/attn_v5_ab.txt)
constexpr int kAttnPolyDefault = 2;  // d = 64 (MUFU-bound): half of the pairs, 0.561 -> 0.472 ms on v5; d = 128: a quarter
                                     // (0.850 -> 0.833 ms; half: 0.861) (profiles/r02/attn_poly_exp2.txt, attn_v5_poly.txt)
// one-pass softmax (v3 / v4): a block whose row sum exceeds this moves the row's reference exponent up by log2(sum)
constexpr float kAttnRebase = 1024.0f;

// ================================================================================================================
// v3 (round 2): 128 keys per block, single-buffered S and P, ping-pong order on the tensor pipe.
//
// Design history: this kernel was motivated by a probe that showed ~154 cycles per tcgen05.mma for ANY N <= 128 (so v2's
// N = 64 key blocks looked twice as expensive as N = 128 ones). That figure was the single-lane issue loop, not the
// tensor pipe (with the warp-uniform issue loop below an N = 128 instruction costs its 64 pipe
// cycles, an N = 64 one ~45). 128 keys per block still halve the number of instructions, barrier round trips and
// softmax passes per key, and this is the fastest of the four kernels at both head sizes.
//   * TMEM: S_a @0, S_b @128 (128 fp32 columns each, ONE buffer per group), O_a @256, O_b @384.
//   * tensor-pipe order per block j:  PV_a(j), S_a(j+1), PV_b(j), S_b(j+1).  Group a exponentiates block j while
//     the pipe runs PV_b(j-1) and S_b(j), group b while it runs PV_a(j) and S_a(j+1): each group has a window of two
//     instruction groups (~2.5 k cycles) for 128 x 128 exponentials on four warps, so neither S nor P needs a second
//     buffer (a double-buffered P alone would be 128 KB of shared memory at d = 128).
//   * K and V tiles (128 keys each) travel through ONE ring in consumption order K0 V0 K1 V1 ... (3 slots of 32 KB at
//     d = 128: live at any time are V(j), K(j+1) and the tile in flight).
//   * softmax: the whole 128-score row of a thread is read once (2 x tcgen05.ld.x64) and kept in registers; stale-maximum
//     scheme, lazy O rescale and the d = 64 ones-column row sums are those of v2.
// ================================================================================================================
template <int D>
struct Attn3Cfg {
  static constexpr int kBN = 128;
  static constexpr int kSlots = (D == 64) ? 4 : 3;
  static constexpr int kQBytes = 2 * 128 * D * 2;
  static constexpr int kTileBytes = kBN * D * 2;         // one K or V tile
  static constexpr int kPBytes = 128 * kBN * 2;          // P of one group: two 64-key panels of [128 rows x 128 B]
  static constexpr bool kOnesCol = (D == 64);
  static constexpr int kOnesBytes = kOnesCol ? kBN * 128 : 0;
  static constexpr int kSmemBytes = kQBytes + kSlots * kTileBytes + 2 * kPBytes + kOnesBytes + 1024 + 256;
};

// PMASK: which of the 4 pairs of every 8 scores take exp2_poly2 instead of MUFU.EX2 (bit i = pair i)
template <int D, int PMASK>
__global__ void __launch_bounds__(320, 1)
flash_attn3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using C = Attn3Cfg<D>;
  constexpr int BN = C::kBN, NS = C::kSlots;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                             // [group][d panel][128 rows x 128 B]
  uint8_t* sKV = sQ + C::kQBytes;                 // [slot][d panel][128 keys x 128 B]
  uint8_t* sP = sKV + NS * C::kTileBytes;         // [group][key panel][128 rows x 128 B]
  uint8_t* sOnes = sP + 2 * C::kPBytes;           // [128 keys x 128 B] of bf16 1.0 (d = 64 only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + C::kOnesBytes);
  uint64_t* q_full = bars;              // 1
  uint64_t* kv_full = q_full + 1;       // [NS]
  uint64_t* kv_empty = kv_full + NS;    // [NS]
  uint64_t* s_full = kv_empty + NS;     // [2 groups]
  uint64_t* p_full = s_full + 2;        // [2]
  uint64_t* pv_done = p_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // uniform: see the MMA issuer
  const int m0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int nblk = (p.N + BN - 1) / BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int g = 0; g < 2; ++g) { mbar_init(&s_full[g], 1); mbar_init(&p_full[g], 128); mbar_init(&pv_done[g], 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (C::kOnesCol) {
    for (int i = threadIdx.x; i < C::kOnesBytes / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_launch_dependents();
  pdl_wait();  // nothing above touches global memory

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer: Q (both groups) once, then K0 V0 K1 V1 ... through the ring ----
      mbar_expect_tx(q_full, C::kQBytes);
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int pnl = 0; pnl < D / 64; ++pnl)
          tma_load_4d(sQ + (g * (D / 64) + pnl) * (128 * 128), &tmQ, q_full, pnl * 64, m0 + g * 128, h, b);
      int slot = 0;
      uint32_t phase = 0;
      for (int t = 0; t < 2 * nblk; ++t) {
        mbar_wait_relaxed(&kv_empty[slot], phase ^ 1);
        uint8_t* dst = sKV + slot * C::kTileBytes;
        mbar_expect_tx(&kv_full[slot], C::kTileBytes);
        const CUtensorMap* tm = (t & 1) ? &tmV : &tmK;
#pragma unroll
        for (int pnl = 0; pnl < D / 64; ++pnl)
          tma_load_4d(dst + pnl * (BN * 128), tm, &kv_full[slot], pnl * 64, (t >> 1) * BN, h, b);
        if (++slot == NS) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    {  // whole warp on warp-uniform values, one elected lane issues (gemm_conv_tc.cu has the argument)
      const bool leader = elect_one();
      // ---- MMA issuer ----
      const uint32_t idesc_s = umma_idesc_bf16(128, BN, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, C::kOnesCol ? D + 16 : D, 0, 1);
      auto tile_addr = [&](int t) { return smem_u32(sKV + (t % NS) * C::kTileBytes); };
      auto wait_tile = [&](int t) { mbar_wait(&kv_full[t % NS], (t / NS) & 1); tc_fence_after(); };
      auto free_tile = [&](int t) { if (leader) tc_commit(&kv_empty[t % NS]); };
      auto issue_s = [&](int g, int t) {
        const uint32_t aQ = smem_u32(sQ + g * (D / 64) * (128 * 128));
        const uint32_t aK = tile_addr(t);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint64_t da = umma_desc_sw128(aQ + (kk >> 2) * (128 * 128) + (kk & 3) * 32, 16, 1024);
          const uint64_t db = umma_desc_sw128(aK + (kk >> 2) * (BN * 128) + (kk & 3) * 32, 16, 1024);
          if (leader) tc_mma_ss(tmem_base + g * 128, da, db, idesc_s, kk != 0);
        }
        if (leader) tc_commit(&s_full[g]);
      };
      auto issue_pv = [&](int g, int t, int j) {
        const uint32_t aP = smem_u32(sP + g * C::kPBytes);
        const uint32_t aV = tile_addr(t);
        // second 64-wide N atom: the next d panel (d = 128) or the all-ones panel (d = 64)
        const uint32_t lbo = C::kOnesCol ? smem_u32(sOnes) - aV : BN * 128;
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
          const uint64_t da = umma_desc_sw128(aP + (kk >> 2) * (128 * 128) + (kk & 3) * 32, 16, 1024);
          const uint64_t db = umma_desc_sw128(aV + kk * 2048, lbo, 1024);
          if (leader) tc_mma_ss(tmem_base + 256 + g * 128, da, db, idesc_o, (j | kk) != 0);
        }
        if (leader) tc_commit(&pv_done[g]);
      };
      mbar_wait(q_full, 0);
      wait_tile(0);
      issue_s(0, 0);
      issue_s(1, 0);
      free_tile(0);
      for (int j = 0; j < nblk; ++j) {
        const int tV = 2 * j + 1, tK = 2 * j + 2;
        const bool more = j + 1 < nblk;
        mbar_wait(&p_full[0], j & 1);   // group a consumed S_a(j) and wrote P_a(j)
        wait_tile(tV);
        issue_pv(0, tV, j);
        if (more) { wait_tile(tK); issue_s(0, tK); }
        mbar_wait(&p_full[1], j & 1);
        tc_fence_after();
        issue_pv(1, tV, j);
        free_tile(tV);
        if (more) { issue_s(1, tK); free_tile(tK); }
      }
    }
  } else {
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + g * 128 + lane_off, tO = tmem_base + 256 + g * 128 + lane_off;
    uint8_t* p_row = sP + g * C::kPBytes + row * 128;
    float mb = 0.f, l_run = 0.f;  // mb: the row's reference exponent (log2 units), see flash_attn4_kernel
    float pend_alpha = 1.0f;      // rescale decided at the end of the previous block
    bool pend = false;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&s_full[g], j & 1);
      tc_fence_after();
      const int n_valid = min(BN, p.N - j * BN);
      if (j == 0) {  // the first block fixes the reference: one extra scan of its scores for the maximum
        float mx = -INFINITY;
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t hv[32];
          tmem_ld32(tS + c0, hv);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < n_valid) mx = fmaxf(mx, __uint_as_float(hv[i]));
        }
        mb = mx * p.scale_log2;
      } else {
        // P_g and O_g belong to PV_g(j - 1) until it has retired
        mbar_wait(&pv_done[g], (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, pend)) {  // O rows are thread-local here: a warp rescales on its own
#pragma unroll
          for (int c0 = 0; c0 < (C::kOnesCol ? D + 32 : D); c0 += 32) {  // + the row-sum column(s) at D
            uint32_t v[32];
            tmem_ld32(tO + c0, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * pend_alpha);
            tmem_st32(tO + c0, v);
          }
          tmem_wait_st();
          l_run *= pend_alpha;
        }
      }
      // ONE pass over the row (flash_attn4_kernel has the argument): exponentials against the running reference; the
      // block's row sum decides whether the NEXT block starts from a new one
      const float nmb = -mb;
      float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t sv[32];
        tmem_ld32(tS + c0, sv);
        tmem_wait_ld();
        if (n_valid < BN) {  // ragged last block only
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i >= n_valid) sv[i] = 0xff800000u;  // -inf
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float pr[8];
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            float t0, t1;
            ffma2(t0, t1, __uint_as_float(sv[q * 8 + i]), __uint_as_float(sv[q * 8 + i + 1]), p.scale_log2, nmb);
            if ((PMASK >> (i >> 1)) & 1) {
              exp2_poly2(t0, t1);  // (-inf -> 2^-125: negligible against the row sum, and its V row is zero)
              pr[i] = t0; pr[i + 1] = t1;
            } else {
              pr[i] = exp2f(t0);  // exp2(-inf) = 0
              pr[i + 1] = exp2f(t1);
            }
            fadd2(ls0, ls1, pr[i], pr[i + 1]);
          }
          uint4 u;
          u.x = pack_bf16x2(pr[0], pr[1]); u.y = pack_bf16x2(pr[2], pr[3]);
          u.z = pack_bf16x2(pr[4], pr[5]); u.w = pack_bf16x2(pr[6], pr[7]);
          const int ch = (c0 >> 3) + q;  // 16-byte chunk of the 256-byte row: two 64-key panels
          *reinterpret_cast<uint4*>(p_row + (ch >> 3) * (128 * 128) + (((ch & 7) ^ (row & 7)) << 4)) = u;
        }
      }
      const float ls = ls0 + ls1;
      l_run += ls;
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&p_full[g]);
      pend = ls > kAttnRebase;
      const float lg = pend ? log2f(ls) : 0.f;
      pend_alpha = exp2f(-lg);
      mb += lg;
    }
    // epilogue: O / l
    mbar_wait(&pv_done[g], (nblk - 1) & 1);
    tc_fence_after();
    const int m = m0 + g * 128 + row;
    if (C::kOnesCol) {  // l = the ones column of O
      uint32_t lv[16];
      tmem_ld16(tO + D, lv);
      tmem_wait_ld();
      l_run = __uint_as_float(lv[0]);
    }
    const float inv = 1.0f / l_run;
    bf16* o = p.out + static_cast<long long>(b) * p.o_batch + static_cast<long long>(m) * p.o_row +
              static_cast<long long>(h) * p.o_head;
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tO + c0, v);
      tmem_wait_ld();
      if (m < p.M) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]) * inv, __uint_as_float(v[q * 8 + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]) * inv, __uint_as_float(v[q * 8 + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]) * inv, __uint_as_float(v[q * 8 + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]) * inv, __uint_as_float(v[q * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(o + c0 + q * 8) = u;
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================================
// v5 (end of round 2): v3 with the S accumulator handed back as soon as a row is in REGISTERS.
//
// ncu on v3 (profiles/r02/ncu_r02_attn3_final.txt): the softmax warps wait for S 40 % of the time while the MMA issuer
// waits ~900 cycles per block for P - each group is one serial chain S MMA -> softmax -> P -> PV MMA -> next S, and two
// chains do not fill the tensor pipe (60.8 % active). Here a softmax thread first copies its whole row (128 fp32
// scores) out of TMEM and arrives on s_free; the issuer then runs S_g(j+1) UNDER the softmax arithmetic of block j, so
// the chain per block is softmax -> PV only. 128 score registers + the rest need ~200 registers per softmax thread:
// the roles sit in whole warpgroups (producer + issuer + two idle warps | two softmax groups) and setmaxnreg moves the
// first warpgroup's registers (40 each) to the softmax warps (232 each).
// Tile order on the ring is the order of use: K0 | K1 V0 | K2 V1 | ...
// ================================================================================================================
template <int D, int PMASK>
__global__ void __launch_bounds__(384, 1)
flash_attn5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using C = Attn3Cfg<D>;
  constexpr int BN = C::kBN, NS = C::kSlots;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                             // [group][d panel][128 rows x 128 B]
  uint8_t* sKV = sQ + C::kQBytes;                 // [slot][d panel][128 keys x 128 B]
  uint8_t* sP = sKV + NS * C::kTileBytes;         // [group][key panel][128 rows x 128 B]
  uint8_t* sOnes = sP + 2 * C::kPBytes;           // [128 keys x 128 B] of bf16 1.0 (d = 64 only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + C::kOnesBytes);
  uint64_t* q_full = bars;              // 1
  uint64_t* kv_full = q_full + 1;       // [NS]
  uint64_t* kv_empty = kv_full + NS;    // [NS]
  uint64_t* s_full = kv_empty + NS;     // [2 groups]
  uint64_t* p_full = s_full + 2;        // [2]
  uint64_t* pv_done = p_full + 2;       // [2]
  uint64_t* s_free = pv_done + 2;       // [2] 128 arrivals: the group holds its scores in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // uniform: see the MMA issuer
  const int m0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int nblk = (p.N + BN - 1) / BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1); mbar_init(&p_full[g], 128); mbar_init(&pv_done[g], 1); mbar_init(&s_free[g], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (C::kOnesCol) {
    for (int i = threadIdx.x; i < C::kOnesBytes / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_launch_dependents();
  pdl_wait();  // nothing above touches global memory

  if (warp < 4) {
  setmaxnreg_dec<40>();
  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer: Q (both groups) once, then K0 | K1 V0 | K2 V1 | ... through the ring ----
      mbar_expect_tx(q_full, C::kQBytes);
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int pnl = 0; pnl < D / 64; ++pnl)
          tma_load_4d(sQ + (g * (D / 64) + pnl) * (128 * 128), &tmQ, q_full, pnl * 64, m0 + g * 128, h, b);
      int slot = 0;
      uint32_t phase = 0;
      auto load = [&](const CUtensorMap* tm, int blk) {
        mbar_wait_relaxed(&kv_empty[slot], phase ^ 1);
        uint8_t* dst = sKV + slot * C::kTileBytes;
        mbar_expect_tx(&kv_full[slot], C::kTileBytes);
#pragma unroll
        for (int pnl = 0; pnl < D / 64; ++pnl)
          tma_load_4d(dst + pnl * (BN * 128), tm, &kv_full[slot], pnl * 64, blk * BN, h, b);
        if (++slot == NS) { slot = 0; phase ^= 1; }
      };
      load(&tmK, 0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) load(&tmK, j + 1);
        load(&tmV, j);
      }
    }
  } else if (warp == 1) {
    {  // whole warp on warp-uniform values, one elected lane issues (gemm_conv_tc.cu has the argument)
      const bool leader = elect_one();
      // ---- MMA issuer ----
      const uint32_t idesc_s = umma_idesc_bf16(128, BN, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, C::kOnesCol ? D + 16 : D, 0, 1);
      int slot = 0;     // ring position of the next tile to consume (same order as the producer)
      uint32_t phase = 0;
      auto next_tile = [&]() {
        mbar_wait(&kv_full[slot], phase);
        tc_fence_after();
        const int s0 = slot;
        if (++slot == NS) { slot = 0; phase ^= 1; }
        return s0;
      };
      auto tile_addr = [&](int s0) { return smem_u32(sKV + s0 * C::kTileBytes); };
      auto free_tile = [&](int s0) { if (leader) tc_commit(&kv_empty[s0]); };
      auto issue_s = [&](int g, int t) {
        const uint32_t aQ = smem_u32(sQ + g * (D / 64) * (128 * 128));
        const uint32_t aK = tile_addr(t);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint64_t da = umma_desc_sw128(aQ + (kk >> 2) * (128 * 128) + (kk & 3) * 32, 16, 1024);
          const uint64_t db = umma_desc_sw128(aK + (kk >> 2) * (BN * 128) + (kk & 3) * 32, 16, 1024);
          if (leader) tc_mma_ss(tmem_base + g * 128, da, db, idesc_s, kk != 0);
        }
        if (leader) tc_commit(&s_full[g]);
      };
      auto issue_pv = [&](int g, int t, int j) {
        const uint32_t aP = smem_u32(sP + g * C::kPBytes);
        const uint32_t aV = tile_addr(t);
        // second 64-wide N atom: the next d panel (d = 128) or the all-ones panel (d = 64)
        const uint32_t lbo = C::kOnesCol ? smem_u32(sOnes) - aV : BN * 128;
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
          const uint64_t da = umma_desc_sw128(aP + (kk >> 2) * (128 * 128) + (kk & 3) * 32, 16, 1024);
          const uint64_t db = umma_desc_sw128(aV + kk * 2048, lbo, 1024);
          if (leader) tc_mma_ss(tmem_base + 256 + g * 128, da, db, idesc_o, (j | kk) != 0);
        }
        if (leader) tc_commit(&pv_done[g]);
      };
      mbar_wait(q_full, 0);
      {
        const int s0 = next_tile();  // K(0)
        issue_s(0, s0);
        issue_s(1, s0);
        free_tile(s0);
      }
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) {
          const int sk = next_tile();  // K(j + 1)
          for (int g = 0; g < 2; ++g) {
            mbar_wait(&s_free[g], j & 1);  // group g holds S_g(j) in registers: its TMEM columns may be overwritten
            tc_fence_after();
            issue_s(g, sk);
          }
          free_tile(sk);
        }
        const int sv = next_tile();    // V(j)
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&p_full[g], j & 1);
          tc_fence_after();
          issue_pv(g, sv, j);
        }
        free_tile(sv);
      }
    }
  }
  } else {
    setmaxnreg_inc<232>();
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + g * 128 + lane_off, tO = tmem_base + 256 + g * 128 + lane_off;
    uint8_t* p_row = sP + g * C::kPBytes + row * 128;
    float mb = 0.f, l_run = 0.f;  // mb: the row's reference exponent (log2 units), see flash_attn4_kernel
    float pend_alpha = 1.0f;      // rescale decided at the end of the previous block
    bool pend = false;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&s_full[g], j & 1);
      tc_fence_after();
      const int n_valid = min(BN, p.N - j * BN);
      // the whole row leaves TMEM first: S_g may then be overwritten by the next block's MMAs while we do the arithmetic
      uint32_t sv[128];
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) tmem_ld32(tS + c0, *reinterpret_cast<uint32_t(*)[32]>(&sv[c0]));
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&s_free[g]);
      if (n_valid < BN) {  // ragged last block only
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= n_valid) sv[i] = 0xff800000u;  // -inf
      }
      if (j == 0) {  // the first block fixes the reference at its maximum
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 128; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
        mb = mx * p.scale_log2;
      } else {
        // P_g and O_g belong to PV_g(j - 1) until it has retired
        mbar_wait(&pv_done[g], (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, pend)) {  // O rows are thread-local here: a warp rescales on its own
#pragma unroll
          for (int c0 = 0; c0 < (C::kOnesCol ? D + 32 : D); c0 += 32) {  // + the row-sum column(s) at D
            uint32_t v[32];
            tmem_ld32(tO + c0, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * pend_alpha);
            tmem_st32(tO + c0, v);
          }
          tmem_wait_st();
          l_run *= pend_alpha;
        }
      }
      // ONE pass over the row (flash_attn4_kernel has the argument): exponentials against the running reference; the
      // block's row sum decides whether the NEXT block starts from a new one
      const float nmb = -mb;
      float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {  // 16-byte chunk (8 keys) of the 256-byte P row: two 64-key panels
        float pr[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          float t0, t1;
          ffma2(t0, t1, __uint_as_float(sv[ch * 8 + i]), __uint_as_float(sv[ch * 8 + i + 1]), p.scale_log2, nmb);
          if ((PMASK >> (i >> 1)) & 1) {
            exp2_poly2(t0, t1);  // (-inf -> 2^-125: negligible against the row sum, and its V row is zero)
            pr[i] = t0; pr[i + 1] = t1;
          } else {
            pr[i] = exp2f(t0);  // exp2(-inf) = 0
            pr[i + 1] = exp2f(t1);
          }
          fadd2(ls0, ls1, pr[i], pr[i + 1]);
        }
        uint4 u;
        u.x = pack_bf16x2(pr[0], pr[1]); u.y = pack_bf16x2(pr[2], pr[3]);
        u.z = pack_bf16x2(pr[4], pr[5]); u.w = pack_bf16x2(pr[6], pr[7]);
        *reinterpret_cast<uint4*>(p_row + (ch >> 3) * (128 * 128) + (((ch & 7) ^ (row & 7)) << 4)) = u;
      }
      const float ls = ls0 + ls1;
      l_run += ls;
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&p_full[g]);
      pend = ls > kAttnRebase;
      const float lg = pend ? log2f(ls) : 0.f;
      pend_alpha = exp2f(-lg);
      mb += lg;
    }
    // epilogue: O / l
    mbar_wait(&pv_done[g], (nblk - 1) & 1);
    tc_fence_after();
    const int m = m0 + g * 128 + row;
    if (C::kOnesCol) {  // l = the ones column of O
      uint32_t lv[16];
      tmem_ld16(tO + D, lv);
      tmem_wait_ld();
      l_run = __uint_as_float(lv[0]);
    }
    const float inv = 1.0f / l_run;
    bf16* o = p.out + static_cast<long long>(b) * p.o_batch + static_cast<long long>(m) * p.o_row +
              static_cast<long long>(h) * p.o_head;
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tO + c0, v);
      tmem_wait_ld();
      if (m < p.M) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]) * inv, __uint_as_float(v[q * 8 + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]) * inv, __uint_as_float(v[q * 8 + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]) * inv, __uint_as_float(v[q * 8 + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]) * inv, __uint_as_float(v[q * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(o + c0 + q * 8) = u;
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int D>
static int launch_attn3(const void* q, const void* k, const void* v, void* out, int B, int H, int M, int N,
                        const long long* qs, const long long* ks, const long long* vs, const long long* os, float scale,
                        cudaStream_t st) {
  using C = Attn3Cfg<D>;
  // fraction of the exponentials on the FMA pipe: 0 none, 1 one pair in four, 2 every other pair (OMGSR_ATTN_POLY)
  static const int poly = getenv("OMGSR_ATTN_POLY") ? atoi(getenv("OMGSR_ATTN_POLY")) : (D == 64 ? kAttnPolyDefault : 1);
  static unsigned long long attr_done[3] = {0, 0, 0};  // one bit per device
  const void* kfn = poly == 2 ? reinterpret_cast<const void*>(&flash_attn3_kernel<D, 0xA>)
                  : poly == 1 ? reinterpret_cast<const void*>(&flash_attn3_kernel<D, 0x8>)
                              : reinterpret_cast<const void*>(&flash_attn3_kernel<D, 0>);
  if (int rc = ensure_smem_attr(kfn, C::kSmemBytes, &attr_done[poly == 2 ? 2 : (poly == 1 ? 1 : 0)], "flash_attn3")) return rc;
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t eb = 2;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(M), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    uint64_t str[3] = {qs[0] * eb, qs[1] * eb, qs[2] * eb};
    uint32_t box[4] = {64, 128, 1, 1};
    int rc = make_tmap_bf16(&tmQ, q, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(N), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    uint32_t box[4] = {64, static_cast<uint32_t>(C::kBN), 1, 1};
    uint64_t strk[3] = {ks[0] * eb, ks[1] * eb, ks[2] * eb};
    int rc = make_tmap_bf16(&tmK, k, 4, dims, strk, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    uint64_t strv[3] = {vs[0] * eb, vs[1] * eb, vs[2] * eb};
    rc = make_tmap_bf16(&tmV, v, 4, dims, strv, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  AttnParams p;
  p.M = M; p.N = N; p.H = H; p.B = B;
  p.o_row = os[0]; p.o_head = os[1]; p.o_batch = os[2];
  p.out = static_cast<bf16*>(out);
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((M + 255) / 256, H, B);
  if (poly == 2) launch_pdl(flash_attn3_kernel<D, 0xA>, dim3(grid), dim3(320), C::kSmemBytes, st, tmQ, tmK, tmV, p);
  else if (poly == 1) launch_pdl(flash_attn3_kernel<D, 0x8>, dim3(grid), dim3(320), C::kSmemBytes, st, tmQ, tmK, tmV, p);
  else launch_pdl(flash_attn3_kernel<D, 0>, dim3(grid), dim3(320), C::kSmemBytes, st, tmQ, tmK, tmV, p);
  return check_launch("flash_attn3_kernel");
}
template <int D>
static int launch_attn5(const void* q, const void* k, const void* v, void* out, int B, int H, int M, int N,
                        const long long* qs, const long long* ks, const long long* vs, const long long* os, float scale,
                        cudaStream_t st) {
  using C = Attn3Cfg<D>;
  // fraction of the exponentials on the FMA pipe: 0 none, 1 one pair in four, 2 every other pair (OMGSR_ATTN_POLY)
  static const int poly = getenv("OMGSR_ATTN_POLY") ? atoi(getenv("OMGSR_ATTN_POLY")) : (D == 64 ? kAttnPolyDefault : 1);
  static unsigned long long attr_done[3] = {0, 0, 0};  // one bit per device
  const void* kfn = poly == 2 ? reinterpret_cast<const void*>(&flash_attn5_kernel<D, 0xA>)
                  : poly == 1 ? reinterpret_cast<const void*>(&flash_attn5_kernel<D, 0x8>)
                              : reinterpret_cast<const void*>(&flash_attn5_kernel<D, 0>);
  if (int rc = ensure_smem_attr(kfn, C::kSmemBytes, &attr_done[poly == 2 ? 2 : (poly == 1 ? 1 : 0)], "flash_attn5")) return rc;
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t eb = 2;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(M), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    uint64_t str[3] = {qs[0] * eb, qs[1] * eb, qs[2] * eb};
    uint32_t box[4] = {64, 128, 1, 1};
    int rc = make_tmap_bf16(&tmQ, q, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(N), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    uint32_t box[4] = {64, static_cast<uint32_t>(C::kBN), 1, 1};
    uint64_t strk[3] = {ks[0] * eb, ks[1] * eb, ks[2] * eb};
    int rc = make_tmap_bf16(&tmK, k, 4, dims, strk, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    uint64_t strv[3] = {vs[0] * eb, vs[1] * eb, vs[2] * eb};
    rc = make_tmap_bf16(&tmV, v, 4, dims, strv, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  AttnParams p;
  p.M = M; p.N = N; p.H = H; p.B = B;
  p.o_row = os[0]; p.o_head = os[1]; p.o_batch = os[2];
  p.out = static_cast<bf16*>(out);
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((M + 255) / 256, H, B);
  if (poly == 2) launch_pdl(flash_attn5_kernel<D, 0xA>, dim3(grid), dim3(384), C::kSmemBytes, st, tmQ, tmK, tmV, p);
  else if (poly == 1) launch_pdl(flash_attn5_kernel<D, 0x8>, dim3(grid), dim3(384), C::kSmemBytes, st, tmQ, tmK, tmV, p);
  else launch_pdl(flash_attn5_kernel<D, 0>, dim3(grid), dim3(384), C::kSmemBytes, st, tmQ, tmK, tmV, p);
  return check_launch("flash_attn5_kernel");
}

// ================================================================================================================
// v4 (round 2, d = 128): v3 with the PV product TRANSPOSED so that it, too, is made of N = 256 instructions.
//
// (Kept as OMGSR_ATTN_VARIANT=4, no longer the default: it was built on the wrong "154 cycles per N <= 128 instruction"
// figure of the single-lane issue loop, DESIGN.md section 4. At the real costs - 64 cycles per N = 128, 128 per N = 256
// instruction - the transposed product saves no tensor-pipe time and v3 is faster: 0.895 vs 0.977 ms on the FLUX shape.)
// The accumulator is O^T [d = 128 lanes, 256 query columns (both groups)]: A = V^T (the V tile read as an MN-major A
// operand), B = P of BOTH groups ([256 query rows, 16 keys], K-major: the two groups' P tiles are adjacent in shared
// memory), one N = 256 instruction per 16 keys serves 256 query rows.
//   * TMEM: S_a @0, S_b @128 (single buffers), O^T @256 (256 columns: group a's query rows are columns 0..127).
//   * tensor-pipe order: S_a(j+1), S_b(j+1), PV^T(j). A group reads its S, exponentiates into REGISTERS (bf16-packed,
//     64 registers), releases S at once (s_free) so that the next S can be issued, and stores P only when PV^T(j-1) has
//     retired (pv_done): no second P buffer (there is no shared memory for one: Q 64 + K/V ring 96 + P 64 KB).
//   * lazy rescale per group: O^T columns belong to ONE group, so a group rescales its own 128 columns (each warp its
//     lane quarter) with the per-row factors passed through shared memory; the "does any row of the group need it"
//     vote is a named-barrier reduction of the group's 128 threads.
//   * epilogue: O^T / l is transposed through shared memory (the dead Q tile of the group) and leaves as coalesced
//     256-byte rows.
// ================================================================================================================
struct Attn4Cfg {
  static constexpr int D = 128;
  static constexpr int kBN = 128;
  static constexpr int kSlots = 3;
  static constexpr int kQBytes = 2 * 128 * D * 2;
  static constexpr int kTileBytes = kBN * D * 2;
  static constexpr int kPBytes = 2 * 256 * 128;      // [key panel][group][128 rows x 128 B]
  static constexpr int kMiscBytes = 256 + 2 * 128 * 4;  // barriers, alpha[2][128] (reused for 1 / l in the epilogue)
  static constexpr int kSmemBytes = kQBytes + kSlots * kTileBytes + kPBytes + kMiscBytes + 1024;
};

__device__ __forceinline__ bool group_any(int barrier_id, bool pred) {  // OR over the 128 threads of a softmax group
  uint32_t out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "bar.red.or.pred p, %1, 128, q;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}\n"
      : "=r"(out)
      : "r"(barrier_id), "r"(static_cast<uint32_t>(pred))
      : "memory");
  return out != 0;
}
__device__ __forceinline__ void group_sync(int barrier_id) {
  asm volatile("bar.sync %0, 128;" ::"r"(barrier_id) : "memory");
}

__global__ void __launch_bounds__(320, 1)
flash_attn4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using C = Attn4Cfg;
  constexpr int D = C::D, BN = C::kBN, NS = C::kSlots;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                             // [group][d panel][128 rows x 128 B]; epilogue: transposed O tile
  uint8_t* sKV = sQ + C::kQBytes;                 // [slot][d panel][128 keys x 128 B]
  uint8_t* sP = sKV + NS * C::kTileBytes;         // [key panel][group][128 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + C::kPBytes);
  uint64_t* q_full = bars;              // 1
  uint64_t* kv_full = q_full + 1;       // [NS]
  uint64_t* kv_empty = kv_full + NS;    // [NS]
  uint64_t* s_full = kv_empty + NS;     // [2]
  uint64_t* s_free = s_full + 2;        // [2] 128 arrivals: the group holds its scores in registers
  uint64_t* p_full = s_free + 2;        // [2] 128 arrivals
  uint64_t* pv_done = p_full + 2;       // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);
  float* s_alpha = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2][128]

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // uniform: see the MMA issuer
  const int m0 = blockIdx.x * 256, h = blockIdx.y, b = blockIdx.z;
  const int nblk = (p.N + BN - 1) / BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int g = 0; g < 2; ++g) { mbar_init(&s_full[g], 1); mbar_init(&s_free[g], 128); mbar_init(&p_full[g], 128); }
    mbar_init(pv_done, 1);
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

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer: Q once, then the K / V tiles in consumption order K0 | K1 V0 | K2 V1 | ... ----
      mbar_expect_tx(q_full, C::kQBytes);
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int pnl = 0; pnl < D / 64; ++pnl)
          tma_load_4d(sQ + (g * (D / 64) + pnl) * (128 * 128), &tmQ, q_full, pnl * 64, m0 + g * 128, h, b);
      int slot = 0;
      uint32_t phase = 0;
      auto load = [&](const CUtensorMap* tm, int blk) {
        mbar_wait_relaxed(&kv_empty[slot], phase ^ 1);
        uint8_t* dst = sKV + slot * C::kTileBytes;
        mbar_expect_tx(&kv_full[slot], C::kTileBytes);
#pragma unroll
        for (int pnl = 0; pnl < D / 64; ++pnl)
          tma_load_4d(dst + pnl * (BN * 128), tm, &kv_full[slot], pnl * 64, blk * BN, h, b);
        if (++slot == NS) { slot = 0; phase ^= 1; }
      };
      load(&tmK, 0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) load(&tmK, j + 1);
        load(&tmV, j);
      }
    }
  } else if (warp == 1) {
    {  // whole warp on warp-uniform values, one elected lane issues (gemm_conv_tc.cu has the argument)
      const bool leader = elect_one();
      // ---- MMA issuer ----
      const uint32_t idesc_s = umma_idesc_bf16(128, BN, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, 256, 1, 0);  // A = V^T (MN-major), B = P (K-major), N = 256 rows
      int slot = 0;     // ring position of the next tile to consume (same order as the producer)
      uint32_t phase = 0;
      auto next_tile = [&]() {
        mbar_wait(&kv_full[slot], phase);
        tc_fence_after();
        const int s0 = slot;
        if (++slot == NS) { slot = 0; phase ^= 1; }
        return s0;
      };
      auto issue_s = [&](int g, int s0) {
        const uint32_t aQ = smem_u32(sQ + g * (D / 64) * (128 * 128));
        const uint32_t aK = smem_u32(sKV + s0 * C::kTileBytes);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint64_t da = umma_desc_sw128(aQ + (kk >> 2) * (128 * 128) + (kk & 3) * 32, 16, 1024);
          const uint64_t db = umma_desc_sw128(aK + (kk >> 2) * (BN * 128) + (kk & 3) * 32, 16, 1024);
          if (leader) tc_mma_ss(tmem_base + g * 128, da, db, idesc_s, kk != 0);
        }
        if (leader) tc_commit(&s_full[g]);
      };
      mbar_wait(q_full, 0);
      {
        const int s0 = next_tile();  // K(0)
        issue_s(0, s0);
        issue_s(1, s0);
        if (leader) tc_commit(&kv_empty[s0]);
      }
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) {
          const int s0 = next_tile();  // K(j + 1)
          for (int g = 0; g < 2; ++g) {
            mbar_wait(&s_free[g], j & 1);  // group g holds S_g(j) in registers: its TMEM columns may be overwritten
            tc_fence_after();
            issue_s(g, s0);
          }
          if (leader) tc_commit(&kv_empty[s0]);
        }
        const int sv = next_tile();    // V(j)
        mbar_wait(&p_full[0], j & 1);
        mbar_wait(&p_full[1], j & 1);
        tc_fence_after();
        const uint32_t aV = smem_u32(sKV + sv * C::kTileBytes);
        const uint32_t aP = smem_u32(sP);
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
          // A = V^T: 64-wide d atoms are BN*128 bytes apart (LBO), 8-key groups 1024 B apart (SBO)
          const uint64_t da = umma_desc_sw128(aV + kk * 2048, BN * 128, 1024);
          const uint64_t db = umma_desc_sw128(aP + (kk >> 2) * (256 * 128) + (kk & 3) * 32, 16, 1024);
          if (leader) tc_mma_ss(tmem_base + 256, da, db, idesc_o, (j | kk) != 0);
        }
        if (leader) tc_commit(pv_done);
        if (leader) tc_commit(&kv_empty[sv]);
      }
    }
  } else {
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // softmax role: query row of the group; O^T role: d index (TMEM lane)
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + g * 128 + lane_off;
    const uint32_t tOT = tmem_base + 256 + g * 128 + lane_off;  // this group's 128 query columns, my 32 d lanes
    uint8_t* p_row = sP + g * (128 * 128) + row * 128;
    float* alpha_g = s_alpha + g * 128;
    float mb = 0.f, l_run = 0.f;  // mb: the row's reference exponent (log2 units); p = 2^(s * scale_log2 - mb)
    float pend_alpha = 1.0f;   // rescale decided at the end of the previous block, applied once PV^T of that block retired
    bool pend_any = false;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&s_full[g], j & 1);
      tc_fence_after();
      const int n_valid = min(BN, p.N - j * BN);
      if (j == 0) {  // the first block fixes the reference at its maximum: one extra scan of its scores
        float mx = -INFINITY;
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t hv[32];
          tmem_ld32(tS + c0, hv);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < n_valid) mx = fmaxf(mx, __uint_as_float(hv[i]));
        }
        mb = mx * p.scale_log2;
      }
      // ONE pass over the row: exponentials against the running reference straight into bf16-packed registers. The
      // reference need not be the row maximum: any value within a bounded distance of it keeps p inside fp32 / bf16
      // range, and l and O carry the same factor, which cancels in O / l. So the maximum is not tracked at all; the
      // block's row SUM (needed anyway) is the trigger: sum <= 2^10 bounds every p of the block by 2^10; a larger sum
      // moves the reference up by log2(sum) for the NEXT block (then that block's p would have summed to 1, its largest
      // lying in [1/128, 1]) and l / O^T are rescaled by 1 / sum one block later. Per pair of scores: FFMA2, 2 MUFU.EX2,
      // FADD2, one bf16x2 pack = 2.5 issued instructions per score (4.5 with the maximum tracked in scalar fp32, 7.4 in
      // the two-pass form): with two softmax warps per scheduler the instruction stream of this loop, not the tensor
      // pipe, set the pace of the earlier versions of this kernel.
      const float nmb = -mb;
      uint32_t pk[64];
      float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t sv[32];
        tmem_ld32(tS + c0, sv);
        tmem_wait_ld();
        if (c0 == 96) {  // the whole row has left TMEM: the next S may overwrite it
          tc_fence_before();
          mbar_arrive(&s_free[g]);
        }
        if (n_valid < BN) {  // ragged last block only
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i >= n_valid) sv[i] = 0xff800000u;  // -inf
        }
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float e0, e1;
          ffma2(e0, e1, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]), p.scale_log2, nmb);
          e0 = exp2f(e0);  // exp2(-inf) = 0
          e1 = exp2f(e1);
          fadd2(ls0, ls1, e0, e1);
          pk[(c0 + i) >> 1] = pack_bf16x2(e0, e1);
        }
      }
      // the P tile (and O^T) belong to PV^T(j - 1) until it has retired
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
      }
      if (pend_any) {  // rescale this group's 128 columns of O^T (everything up to block j - 1) to the new reference
        group_sync(1 + g);  // every row's factor is in shared memory
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tOT + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha_g[c0 + i]);
          tmem_st32(tOT + c0, v);
        }
        tmem_wait_st();
        l_run *= pend_alpha;
      }
      const float ls = ls0 + ls1;
      l_run += ls;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        uint4 u = make_uint4(pk[q * 4 + 0], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
        *reinterpret_cast<uint4*>(p_row + (q >> 3) * (256 * 128) + (((q & 7) ^ (row & 7)) << 4)) = u;
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&p_full[g]);
      // does the next block start from a new reference? (the vote also orders this block's reads of alpha_g before
      // the next writes)
      const bool need = ls > kAttnRebase;
      const float lg = need ? log2f(ls) : 0.f;
      pend_alpha = exp2f(-lg);
      mb += lg;
      pend_any = (j + 1 < nblk) && group_any(1 + g, need);
      if (pend_any) alpha_g[row] = pend_alpha;
    }
    // a rescale decided in the last block never happened: l and O^T are both still in the old reference
    // ---- epilogue: O^T / l -> shared memory (transposed) -> coalesced rows ----
    mbar_wait(pv_done, (nblk - 1) & 1);
    tc_fence_after();
    float* invl_g = alpha_g;  // the rescale factors are dead now
    invl_g[row] = 1.0f / l_run;
    group_sync(1 + g);
    bf16* tile = reinterpret_cast<bf16*>(sQ + g * (2 * 128 * 128));  // [128 query rows][128 d], the group's dead Q tile
#pragma unroll
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tOT + c0, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        tile[(c0 + i) * 128 + row] = __float2bfloat16(__uint_as_float(v[i]) * invl_g[c0 + i]);
    }
    tc_fence_before();
    group_sync(1 + g);
    const int tid = (warp - 2 - 4 * g) * 32 + lane;  // 0..127 inside the group
    for (int idx = tid; idx < 128 * 16; idx += 128) {
      const int r = idx >> 4, ch = idx & 15;
      const int m = m0 + g * 128 + r;
      if (m < p.M) {
        bf16* o = p.out + static_cast<long long>(b) * p.o_batch + static_cast<long long>(m) * p.o_row +
                  static_cast<long long>(h) * p.o_head;
        *reinterpret_cast<uint4*>(o + ch * 8) = *reinterpret_cast<const uint4*>(tile + r * 128 + ch * 8);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int launch_attn4(const void* q, const void* k, const void* v, void* out, int B, int H, int M, int N,
                        const long long* qs, const long long* ks, const long long* vs, const long long* os, float scale,
                        cudaStream_t st) {
  using C = Attn4Cfg;
  constexpr int D = 128;
  static unsigned long long attr_done = 0;  // one bit per device
  if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(&flash_attn4_kernel), C::kSmemBytes, &attr_done, "flash_attn4")) return rc;
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t eb = 2;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(M), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    uint64_t str[3] = {qs[0] * eb, qs[1] * eb, qs[2] * eb};
    uint32_t box[4] = {64, 128, 1, 1};
    int rc = make_tmap_bf16(&tmQ, q, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(N), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    uint32_t box[4] = {64, static_cast<uint32_t>(C::kBN), 1, 1};
    uint64_t strk[3] = {ks[0] * eb, ks[1] * eb, ks[2] * eb};
    int rc = make_tmap_bf16(&tmK, k, 4, dims, strk, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    uint64_t strv[3] = {vs[0] * eb, vs[1] * eb, vs[2] * eb};
    rc = make_tmap_bf16(&tmV, v, 4, dims, strv, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  AttnParams p;
  p.M = M; p.N = N; p.H = H; p.B = B;
  p.o_row = os[0]; p.o_head = os[1]; p.o_batch = os[2];
  p.out = static_cast<bf16*>(out);
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((M + 255) / 256, H, B);
  launch_pdl(flash_attn4_kernel, dim3(grid), dim3(320), C::kSmemBytes, st, tmQ, tmK, tmV, p);
  return check_launch("flash_attn4_kernel");
}

template <int D>
static int launch_attn2(const void* q, const void* k, const void* v, void* out, int B, int H, int M, int N,
                        const long long* qs, const long long* ks, const long long* vs, const long long* os, float scale,
                        cudaStream_t st) {
  using C = Attn2Cfg<D>;
  static unsigned long long attr_done = 0;  // one bit per device
  if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(&flash_attn2_kernel<D>), C::kSmemBytes, &attr_done, "flash_attn2")) return rc;
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t eb = 2;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(M), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    uint64_t str[3] = {qs[0] * eb, qs[1] * eb, qs[2] * eb};
    uint32_t box[4] = {64, 128, 1, 1};
    int rc = make_tmap_bf16(&tmQ, q, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(N), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    uint32_t box[4] = {64, static_cast<uint32_t>(C::kBN), 1, 1};
    uint64_t strk[3] = {ks[0] * eb, ks[1] * eb, ks[2] * eb};
    int rc = make_tmap_bf16(&tmK, k, 4, dims, strk, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    uint64_t strv[3] = {vs[0] * eb, vs[1] * eb, vs[2] * eb};
    rc = make_tmap_bf16(&tmV, v, 4, dims, strv, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  AttnParams p;
  p.M = M; p.N = N; p.H = H; p.B = B;
  p.o_row = os[0]; p.o_head = os[1]; p.o_batch = os[2];
  p.out = static_cast<bf16*>(out);
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((M + 255) / 256, H, B);
  launch_pdl(flash_attn2_kernel<D>, dim3(grid), dim3(320), C::kSmemBytes, st, tmQ, tmK, tmV, p);
  return check_launch("flash_attn2_kernel");
}

template <int D>
static int launch_attn(const void* q, const void* k, const void* v, void* out, int B, int H, int M, int N,
                       const long long* qs, const long long* ks, const long long* vs, const long long* os, float scale,
                       cudaStream_t st) {
  using C = AttnCfg<D>;
  static unsigned long long attr_done = 0;  // one bit per device
  if (int rc = ensure_smem_attr(reinterpret_cast<const void*>(&flash_attn_kernel<D>), C::kSmemBytes, &attr_done, "flash_attn")) return rc;
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t eb = 2;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(M), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    uint64_t str[3] = {qs[0] * eb, qs[1] * eb, qs[2] * eb};
    uint32_t box[4] = {64, static_cast<uint32_t>(C::kBM), 1, 1};
    int rc = make_tmap_bf16(&tmQ, q, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(N), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    uint32_t box[4] = {64, static_cast<uint32_t>(C::kBN), 1, 1};
    uint64_t strk[3] = {ks[0] * eb, ks[1] * eb, ks[2] * eb};
    int rc = make_tmap_bf16(&tmK, k, 4, dims, strk, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    uint64_t strv[3] = {vs[0] * eb, vs[1] * eb, vs[2] * eb};
    rc = make_tmap_bf16(&tmV, v, 4, dims, strv, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  AttnParams p;
  p.M = M; p.N = N; p.H = H; p.B = B;
  p.o_row = os[0]; p.o_head = os[1]; p.o_batch = os[2];
  p.out = static_cast<bf16*>(out);
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((M + C::kBM - 1) / C::kBM, H, B);
  launch_pdl(flash_attn_kernel<D>, dim3(grid), dim3(192), C::kSmemBytes, st, tmQ, tmK, tmV, p);
  return check_launch("flash_attn_kernel");
}

}  // namespace omgsr

using namespace omgsr;

#include <cstdlib>
static int g_attn_variant = getenv("OMGSR_ATTN_VARIANT") ? atoi(getenv("OMGSR_ATTN_VARIANT")) : 0;
/* bring-up / A-B timing switch: 0 = automatic (v5: 128 keys, O in natural layout, S released from registers), 1 = always
 * the single-tile kernel, 2 = the 64-key ping-pong kernel (v2) wherever a 128-key kernel would be chosen, 3 = v3 (v5
 * without the early S release), 4 = v4 (transposed O, N = 256 PV instructions) at d = 128, 5 = v5 explicitly. v4 was the
 * default while every kernel issued MMAs from a single-lane branch (N = 128 instructions then cost 154 cycles, N = 256
 * 172); with the warp-uniform issue loop an N = 128 instruction costs its 64 tensor-pipe cycles and the natural layout
 * is the faster one (profiles/r02/attn_lean_issue.txt, attn_v5_ab.txt). */
extern "C" int omgsr_flash_attn_set_variant(int variant) { g_attn_variant = variant; return OMGSR_OK; }

extern "C" int omgsr_flash_attn_bf16(const void* q, const void* k, const void* v, void* out, int B, int H, int M,
                                     int N, int D, long long q_row, long long q_head, long long q_batch,
                                     long long k_row, long long k_head, long long k_batch, long long v_row,
                                     long long v_head, long long v_batch, long long o_row, long long o_head,
                                     long long o_batch, float scale, void* stream) {
  if (B <= 0 || H <= 0 || M <= 0 || N <= 0) return set_error(OMGSR_ERR_SHAPE, "flash_attn: empty problem");
  if (H > 65535 || B > 65535) return set_error(OMGSR_ERR_SHAPE, "flash_attn: H/B exceed grid limits");
  const long long all = q_row | q_head | q_batch | k_row | k_head | k_batch | v_row | v_head | v_batch | o_row |
                        o_head | o_batch;
  if (all & 7) return set_error(OMGSR_ERR_ALIGN, "flash_attn: strides must be multiples of 8 elements");
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
       reinterpret_cast<uintptr_t>(out)) & 15)
    return set_error(OMGSR_ERR_ALIGN, "flash_attn: pointers must be 16-byte aligned");
  const long long qs[3] = {q_row, q_head, q_batch}, ks[3] = {k_row, k_head, k_batch}, vs[3] = {v_row, v_head, v_batch},
                  os[3] = {o_row, o_head, o_batch};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // long query sequences: ping-pong kernel (2 x 128 rows per CTA, O in TMEM); short ones (M <= 128, e.g. the 8x8
  // UNet level) keep the 128-row kernel so half of every CTA is not padding.
  const bool v2 = (M > 128) && g_attn_variant != 1;
  // long key sequences: 128 keys per block (v3: every S instruction is N = 128, half the instruction count of v2)
  if (v2 && N >= 128 && g_attn_variant != 2) {
    if (g_attn_variant == 5 || (g_attn_variant == 0 && kAttnV5Default)) {
      if (D == 64) return launch_attn5<64>(q, k, v, out, B, H, M, N, qs, ks, vs, os, scale, st);
      if (D == 128) return launch_attn5<128>(q, k, v, out, B, H, M, N, qs, ks, vs, os, scale, st);
    }
    if (D == 64) return launch_attn3<64>(q, k, v, out, B, H, M, N, qs, ks, vs, os, scale, st);
    if (D == 128 && g_attn_variant == 4 && o_head == 128)  // (v4's coalesced row store assumes head-contiguous output)
      return launch_attn4(q, k, v, out, B, H, M, N, qs, ks, vs, os, scale, st);
    if (D == 128) return launch_attn3<128>(q, k, v, out, B, H, M, N, qs, ks, vs, os, scale, st);
  }
  if (v2) {
    if (D == 64) return launch_attn2<64>(q, k, v, out, B, H, M, N, qs, ks, vs, os, scale, st);
    if (D == 128) return launch_attn2<128>(q, k, v, out, B, H, M, N, qs, ks, vs, os, scale, st);
  }
  if (D == 64) return launch_attn<64>(q, k, v, out, B, H, M, N, qs, ks, vs, os, scale, st);
  if (D == 128) return launch_attn<128>(q, k, v, out, B, H, M, N, qs, ks, vs, os, scale, st);
  return set_error(OMGSR_ERR_SHAPE, "flash_attn: head dim %d unsupported (64 or 128)", D);
}
