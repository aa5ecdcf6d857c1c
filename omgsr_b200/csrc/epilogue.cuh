// Shared accumulator epilogue of the tcgen05 kernels: 32 fp32 accumulator columns of one row (one thread)
// -> +bias -> activation -> *gate -> +residual -> bf16|fp32 store.
// The option set of a launch selects ONE specialised per-row routine (epilogue_row<MODE>, dispatched per accumulator
// row, never inside the 32-column chunk loop); bias comes from a per-warp shared-memory slice, the residual is
// prefetched one chunk ahead, outputs leave as 256-bit stores. History (each step from an ncu capture): per-element
// branches cost ~2400 SASS instructions per chunk (12 % tensor-pipe activity on the K=1152 layers); per-chunk branches
// still issued ~330 (mostly predicated-off) instructions; the specialised bodies issue ~95.
#pragma once
#include "common.cuh"
#include "../../include/omgsr_b200.h"

namespace omgsr {

struct EpiArgs {
  const float* bias;      // [Cout] or null
  const float* gate_row;  // gate vector of this row's batch entry, or null
  const bf16* res_row;    // residual row, or null
  void* out;              // output base
  long long o_off;        // element offset of this row
  int Cout;               // accumulator columns that exist (bias length)
  int cout_out;           // output channels (Cout, or Cout/2 for GEGLU)
  int act, act_c0, act_c1;
  int out_f32;
  int vec32;       // bf16 output rows are 32-byte aligned at every 16-channel granule pair: one 256-bit store each
  int res_vec32;   // same for the residual rows: 256-bit loads
  int res_f32;     // res_row points at fp32 data (fp32 residual stream: EPI_F32S / EPI_F32S_GATE only)
  float* gn_part;  // optional: 16 per-thread partials of this 32-column chunk, (sum, sumsq) per 4 channels, filled
                   // by epilogue_store32 for the fused GroupNorm statistics (see gn_chunk_accumulate)
};

__device__ __forceinline__ float warp_sum_all(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int ACT>
__device__ __forceinline__ void act32(float (&f)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (ACT == OMGSR_ACT_SILU) f[j] = silu_f(f[j]);
    else if (ACT == OMGSR_ACT_GELU_ERF) f[j] = gelu_erf_f(f[j]);
    else if (ACT == OMGSR_ACT_GELU_TANH) f[j] = gelu_tanh_f(f[j]);
  }
}

__device__ __forceinline__ void add_bias32(float (&f)[32], const float* bias, int cv, int Cout) {
  if (cv + 32 <= Cout) {
    const float4* bp = reinterpret_cast<const float4*>(bias + cv);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = __ldg(bp + q);
      f[4 * q + 0] += b.x; f[4 * q + 1] += b.y; f[4 * q + 2] += b.z; f[4 * q + 3] += b.w;
    }
  } else {  // ragged tail (Cout not a multiple of 32): clamp the index, the extra columns are never stored
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] += __ldg(bias + min(cv + j, Cout - 1));
  }
}

// Bias staged in shared memory: each epilogue warp keeps the BN bias values of its current n-tile in a private slice,
// refilled (before the accumulator wait, so the latency is hidden) only when the n-tile changes. Reading the bias
// with __ldg per chunk cost an L2 round trip per chunk: with > 200 KB of the SM's L1/shared carve-out given to the
// operand ring, the few KB of L1 left are thrashed by the residual / output streams.
template <int N>
__device__ __forceinline__ void stage_bias(float* slice, const float* bias, int c_first, int Cout, int lane) {
  __syncwarp();
  float v[(N + 31) / 32];
#pragma unroll
  for (int i = 0; i < (N + 31) / 32; ++i) v[i] = __ldg(bias + min(c_first + lane + 32 * i, Cout - 1));  // all in flight
#pragma unroll
  for (int i = 0; i < (N + 31) / 32; ++i)
    if (lane + 32 * i < N) slice[lane + 32 * i] = v[i];
  __syncwarp();
}
// Split form for the persistent kernels: the loads of the NEXT tile's bias are issued before the current tile's chunk
// loop and committed to the slice after it, so their L2 latency never sits on the epilogue's critical path.
template <int N>
__device__ __forceinline__ void bias_prefetch(float (&v)[(N + 31) / 32], const float* bias, int c_first, int Cout,
                                              int lane) {
#pragma unroll
  for (int i = 0; i < (N + 31) / 32; ++i) v[i] = __ldg(bias + min(c_first + lane + 32 * i, Cout - 1));
}
template <int N>
__device__ __forceinline__ void bias_commit(float* slice, const float (&v)[(N + 31) / 32], int lane) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < (N + 31) / 32; ++i)
    if (lane + 32 * i < N) slice[lane + 32 * i] = v[i];
  __syncwarp();
}
__device__ __forceinline__ void add_bias32_s(float (&f)[32], const float* slice) {
  const float4* bp = reinterpret_cast<const float4*>(slice);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = bp[q];
    f[4 * q + 0] += b.x; f[4 * q + 1] += b.y; f[4 * q + 2] += b.z; f[4 * q + 3] += b.w;
  }
}

// Residual operand of one 32-column chunk (4 x 16 B of this thread's row). The kernels issue these loads one chunk
// AHEAD of their use (and before the tcgen05.ld wait): inside epilogue_store32 they sat behind the previous
// granule's store, which the compiler must assume may alias, so every 16-byte load exposed a full HBM round trip
// (measured: 128->128 @512^2 conv 2.11 ms with a residual vs 1.31 ms without).
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256). A thread of the epilogue owns one output row, so a 16-byte
// store covers HALF a 32-byte sector and the other half arrives with a later instruction: every sector reached L2 as
// two partial writes (measured: K = 64 GEMM, i.e. pure epilogue, 2.4 TB/s of output). 32 bytes per thread writes
// whole sectors.
__device__ __forceinline__ void st_global_256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z),
               "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void ld_global_256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p)
               : "memory");
}

__device__ __forceinline__ void load_res32(uint4 (&rr)[4], const EpiArgs& e, int oc0, bool valid) {
  if (e.res_row && valid) {
    if (e.res_vec32 && oc0 + 32 <= e.cout_out) {
      ld_global_256(e.res_row + oc0, rr[0], rr[1]);
      ld_global_256(e.res_row + oc0 + 16, rr[2], rr[3]);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (oc0 + q * 8 < e.cout_out) rr[q] = *reinterpret_cast<const uint4*>(e.res_row + oc0 + q * 8);
    }
  }
}

// pull this thread's residual row (n_ch channels from oc0) into L2 ahead of the accumulator wait
__device__ __forceinline__ void prefetch_res_row(const EpiArgs& e, int oc0, int n_ch, bool valid) {
  if (e.res_row && valid) {
    const char* p = reinterpret_cast<const char*>(e.res_row + oc0);
    const int bytes = min(n_ch, e.cout_out - oc0) * 2;
    for (int o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
  }
}

// f: 32 values for accumulator columns [cv, cv+32) of this row; oc0: first OUTPUT channel they map to;
// rr: the residual of this chunk as loaded by load_res32 (ignored when there is no residual).
__device__ __forceinline__ void epilogue_store32(float (&f)[32], int cv, int oc0, bool valid, const EpiArgs& e,
                                                 const uint4 (&rr4)[4]) {
  if (e.act != OMGSR_ACT_NONE && e.act != OMGSR_ACT_GEGLU && cv >= e.act_c0 && cv < e.act_c1) {
    switch (e.act) {
      case OMGSR_ACT_SILU: act32<OMGSR_ACT_SILU>(f); break;
      case OMGSR_ACT_GELU_ERF: act32<OMGSR_ACT_GELU_ERF>(f); break;
      default: act32<OMGSR_ACT_GELU_TANH>(f); break;
    }
  }
  // NOTE: every lane of the warp runs this loop (the GroupNorm reduction below is warp-collective);
  // rows outside the tensor only skip their loads / stores and contribute zeros.
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int oc = oc0 + q * 8;
    if (oc >= e.cout_out) break;  // warp-uniform
    float* ff = f + q * 8;
    if (e.gate_row) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(e.gate_row + oc));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(e.gate_row + oc + 4));
      ff[0] *= g0.x; ff[1] *= g0.y; ff[2] *= g0.z; ff[3] *= g0.w;
      ff[4] *= g1.x; ff[5] *= g1.y; ff[6] *= g1.z; ff[7] *= g1.w;
    }
    if (e.res_row && valid) {
      const uint4 rr = rr4[q];
      const float2 r0 = unpack_bf16x2(rr.x), r1 = unpack_bf16x2(rr.y), r2 = unpack_bf16x2(rr.z),
                   r3 = unpack_bf16x2(rr.w);
      ff[0] += r0.x; ff[1] += r0.y; ff[2] += r1.x; ff[3] += r1.y;
      ff[4] += r2.x; ff[5] += r2.y; ff[6] += r3.x; ff[7] += r3.y;
    }
    if (e.out_f32) {
      if (valid) {
        float* o = static_cast<float*>(e.out) + e.o_off + oc;
        *reinterpret_cast<float4*>(o) = make_float4(ff[0], ff[1], ff[2], ff[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(ff[4], ff[5], ff[6], ff[7]);
      }
    } else {
      uint4 o;
      o.x = pack_bf16x2(ff[0], ff[1]); o.y = pack_bf16x2(ff[2], ff[3]);
      o.z = pack_bf16x2(ff[4], ff[5]); o.w = pack_bf16x2(ff[6], ff[7]);
      if (valid) *reinterpret_cast<uint4*>(static_cast<bf16*>(e.out) + e.o_off + oc) = o;
      if (e.gn_part) {
        // partial statistics of the values as stored (bf16-rounded): channels 0-3 and 4-7 of this granule
        const float2 v0 = unpack_bf16x2(o.x), v1 = unpack_bf16x2(o.y), v2 = unpack_bf16x2(o.z), v3 = unpack_bf16x2(o.w);
        e.gn_part[4 * q + 0] = valid ? (v0.x + v0.y) + (v1.x + v1.y) : 0.f;
        e.gn_part[4 * q + 1] = valid ? (v0.x * v0.x + v0.y * v0.y) + (v1.x * v1.x + v1.y * v1.y) : 0.f;
        e.gn_part[4 * q + 2] = valid ? (v2.x + v2.y) + (v3.x + v3.y) : 0.f;
        e.gn_part[4 * q + 3] = valid ? (v2.x * v2.x + v2.y * v2.y) + (v3.x * v3.x + v3.y * v3.y) : 0.f;
      }
    }
  }
}

// Compile-time specialisations of epilogue_store32 for the option sets the two pipelines actually use. The generic
// version above is if-converted by the compiler into ~330 issued (mostly predicated-off) instructions per chunk;
// with one epilogue warp per scheduler that, not the tensor pipe, bounded every K <= ~1500 layer (ncu source view:
// 58 % of all samples inside the chunk body). The specialised bodies issue ~100.
template <int ACT, bool GATE, bool RES, bool GN>
__device__ __forceinline__ void epilogue_store32_t(float (&f)[32], int cv, int oc0, bool valid, const EpiArgs& e,
                                                   const uint4 (&rr4)[4]) {
  if (ACT != OMGSR_ACT_NONE && cv >= e.act_c0 && cv < e.act_c1) act32<ACT>(f);
  uint4 o[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int oc = oc0 + q * 8;
    float* ff = f + q * 8;
    if (GATE && oc < e.cout_out) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(e.gate_row + oc));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(e.gate_row + oc + 4));
      ff[0] *= g0.x; ff[1] *= g0.y; ff[2] *= g0.z; ff[3] *= g0.w;
      ff[4] *= g1.x; ff[5] *= g1.y; ff[6] *= g1.z; ff[7] *= g1.w;
    }
    if (RES) {  // rows / granules that are not stored may add stale registers: harmless, they are never written
      const uint4 rr = rr4[q];
      const float2 r0 = unpack_bf16x2(rr.x), r1 = unpack_bf16x2(rr.y), r2 = unpack_bf16x2(rr.z),
                   r3 = unpack_bf16x2(rr.w);
      ff[0] += r0.x; ff[1] += r0.y; ff[2] += r1.x; ff[3] += r1.y;
      ff[4] += r2.x; ff[5] += r2.y; ff[6] += r3.x; ff[7] += r3.y;
    }
    o[q].x = pack_bf16x2(ff[0], ff[1]); o[q].y = pack_bf16x2(ff[2], ff[3]);
    o[q].z = pack_bf16x2(ff[4], ff[5]); o[q].w = pack_bf16x2(ff[6], ff[7]);
  }
  bf16* orow = static_cast<bf16*>(e.out) + e.o_off + oc0;
  if (e.vec32 && oc0 + 32 <= e.cout_out) {  // warp-uniform
    if (valid) {
      st_global_256(orow, o[0], o[1]);
      st_global_256(orow + 16, o[2], o[3]);
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (valid && oc0 + q * 8 < e.cout_out) *reinterpret_cast<uint4*>(orow + q * 8) = o[q];
  }
  if (GN) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const bool on = valid && oc0 + q * 8 < e.cout_out;
      const float2 v0 = unpack_bf16x2(o[q].x), v1 = unpack_bf16x2(o[q].y), v2 = unpack_bf16x2(o[q].z),
                   v3 = unpack_bf16x2(o[q].w);
      e.gn_part[4 * q + 0] = on ? (v0.x + v0.y) + (v1.x + v1.y) : 0.f;
      e.gn_part[4 * q + 1] = on ? (v0.x * v0.x + v0.y * v0.y) + (v1.x * v1.x + v1.y * v1.y) : 0.f;
      e.gn_part[4 * q + 2] = on ? (v2.x + v2.y) + (v3.x + v3.y) : 0.f;
      e.gn_part[4 * q + 3] = on ? (v2.x * v2.x + v2.y * v2.y) + (v3.x * v3.x + v3.y * v3.y) : 0.f;
    }
  }
}

// option set -> specialisation id (0 = generic). Computed once per thread, warp-uniform.
enum : int { EPI_GENERIC = 0, EPI_PLAIN, EPI_RES, EPI_GN, EPI_RES_GN, EPI_GELU_TANH, EPI_GATE_RES, EPI_SILU,
              EPI_F32S, EPI_F32S_GATE };
__host__ __device__ __forceinline__ int epilogue_mode(int act, bool gate, bool res, bool out_f32, bool gn, bool res_f32 = false) {
  if (res_f32) return gate ? EPI_F32S_GATE : EPI_F32S;  // host side guarantees: out_f32, res, no activation, no gn
  if (out_f32) return EPI_GENERIC;
  if (act == OMGSR_ACT_NONE || act == OMGSR_ACT_GEGLU) {  // GEGLU: the kernels fold the gate half before the store
    if (gate) return (res && !gn) ? EPI_GATE_RES : EPI_GENERIC;
    return gn ? (res ? EPI_RES_GN : EPI_GN) : (res ? EPI_RES : EPI_PLAIN);
  }
  if (gate || res || gn) return EPI_GENERIC;
  if (act == OMGSR_ACT_GELU_TANH) return EPI_GELU_TANH;
  if (act == OMGSR_ACT_SILU) return EPI_SILU;
  return EPI_GENERIC;
}
__device__ __forceinline__ void epilogue_store32_m(int mode, float (&f)[32], int cv, int oc0, bool valid,
                                                   const EpiArgs& e, const uint4 (&rr)[4]) {
  switch (mode) {
    case EPI_PLAIN: epilogue_store32_t<OMGSR_ACT_NONE, false, false, false>(f, cv, oc0, valid, e, rr); break;
    case EPI_RES: epilogue_store32_t<OMGSR_ACT_NONE, false, true, false>(f, cv, oc0, valid, e, rr); break;
    case EPI_GN: epilogue_store32_t<OMGSR_ACT_NONE, false, false, true>(f, cv, oc0, valid, e, rr); break;
    case EPI_RES_GN: epilogue_store32_t<OMGSR_ACT_NONE, false, true, true>(f, cv, oc0, valid, e, rr); break;
    case EPI_GELU_TANH: epilogue_store32_t<OMGSR_ACT_GELU_TANH, false, false, false>(f, cv, oc0, valid, e, rr); break;
    case EPI_GATE_RES: epilogue_store32_t<OMGSR_ACT_NONE, true, true, false>(f, cv, oc0, valid, e, rr); break;
    case EPI_SILU: epilogue_store32_t<OMGSR_ACT_SILU, false, false, false>(f, cv, oc0, valid, e, rr); break;
    default: epilogue_store32(f, cv, oc0, valid, e, rr); break;
  }
}

// ---- fused GroupNorm statistics ---------------------------------------------------------------------------
// Reduce NV per-thread values across the 32 lanes (= 32 pixels) of a warp with a halving butterfly:
// NV-1 + (5 - log2 NV) shuffles instead of 5*NV. Afterwards lane L with (L & ((32 >> LOG2) - 1)) == 0 holds the
// total of element L >> (5 - LOG2).
template <int NV, int LOG2>
__device__ __forceinline__ float warp_butterfly(float (&v)[NV], int lane) {
  int step = 16;
#pragma unroll
  for (int half = NV / 2; half >= 1; half >>= 1, step >>= 1) {
    const bool upper = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  float r = v[0];
#pragma unroll
  for (int o = (16 >> LOG2); o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}

// part[16]: (sum, sumsq) of channel quads 0..7 of a 32-channel chunk for this thread's pixel. Adds the warp totals
// per GroupNorm group into acc (this warp's private shared-memory slice, laid out [group][2] for the chunk's
// groups). cpg in {4, 8, 16}.
__device__ __forceinline__ void gn_chunk_accumulate(float (&part)[16], int cpg, float* acc, int lane) {
  if (cpg == 4) {
    const float r = warp_butterfly<16, 4>(part, lane);
    if ((lane & 1) == 0) acc[lane >> 1] += r;
  } else if (cpg == 8) {
    float v[8];
#pragma unroll
    for (int g = 0; g < 4; ++g) { v[2 * g] = part[4 * g] + part[4 * g + 2]; v[2 * g + 1] = part[4 * g + 1] + part[4 * g + 3]; }
    const float r = warp_butterfly<8, 3>(v, lane);
    if ((lane & 3) == 0) acc[lane >> 2] += r;
  } else {  // 16
    float v[4];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      v[2 * g] = (part[8 * g] + part[8 * g + 2]) + (part[8 * g + 4] + part[8 * g + 6]);
      v[2 * g + 1] = (part[8 * g + 1] + part[8 * g + 3]) + (part[8 * g + 5] + part[8 * g + 7]);
    }
    const float r = warp_butterfly<4, 2>(v, lane);
    if ((lane & 7) == 0) acc[lane >> 3] += r;
  }
}

// Deterministic statistics: each epilogue warp owns ONE slot per image in a scratch buffer laid out
// [image][slot = CTA * (epilogue warps per CTA: 4 or 8) + warp][groups * 2] and writes its partial sums there with plain stores (a warp leaves an image
// at most once: the persistent tile loops visit the images in increasing order). Slots nobody writes stay zero (the
// host zeroes the scratch), and gn_finalize_kernel (glue.cu) adds the slots of an image in a fixed order. The first
// version used fp32 atomicAdd into one [groups][2] cell per image: the arrival order changed from run to run and with
// it the rounding (1-3 dB of PSNR noise on the amplifying random-init nets).
__device__ __forceinline__ float* gn_slot(float* scratch, int image, int n, int slot, int slots) {
  return scratch + (static_cast<long long>(image) * slots + slot) * n;  // slot = CTA * warps-per-CTA + warp < slots
}
__device__ __forceinline__ void gn_flush(float* acc, int n, float* slot, int lane) {
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    slot[i] = acc[i];
    acc[i] = 0.f;
  }
  __syncwarp();
}


// ---- one accumulator row segment, specialised per option set ------------------------------------------------
// n_cols output columns of this thread's row, read from TMEM at t_row in 32-column chunks.
//   cv0      first accumulator column (absolute; activation range and GroupNorm group index are relative to it)
//   oc0      first output channel
//   bias_s   this warp's shared-memory bias slice for columns [cv0, ...), or null
//   gate_off GEGLU: TMEM column / bias-slice offset of the gate half
//   rr       residual of the FIRST chunk, loaded by the caller before it waited for the accumulator; on return it
//            holds the first chunk of next_res (if any), so the residual stream stays one chunk ahead across rows
// The option set is a template parameter so that the chunk loop contains no dispatch at all: with the switch inside
// the loop every chunk paid an indirect branch plus ~20 register moves (ncu source view, K = 64 GEMM).
template <int MODE>
struct EpiTraits {
  static constexpr bool kGeneric = MODE == EPI_GENERIC;
  static constexpr bool kRes = MODE == EPI_RES || MODE == EPI_RES_GN || MODE == EPI_GATE_RES;
  static constexpr bool kGn = MODE == EPI_GN || MODE == EPI_RES_GN;
  static constexpr bool kGate = MODE == EPI_GATE_RES;
  static constexpr int kAct = MODE == EPI_GELU_TANH ? OMGSR_ACT_GELU_TANH : (MODE == EPI_SILU ? OMGSR_ACT_SILU : OMGSR_ACT_NONE);
};

template <int MODE, bool GEGLU>
__device__ __forceinline__ void epilogue_row(uint32_t t_row, int n_cols, int gate_off, int cv0, int oc0, bool valid,
                                             EpiArgs& e, const bf16* next_res, int next_oc0, bool next_valid,
                                             const float* bias_s, uint4 (&rr)[4], float* gn_slice, int gn_cpg,
                                             int lane) {
  using T = EpiTraits<MODE>;
  const bool has_res = T::kGeneric ? (e.res_row != nullptr) : T::kRes;
  const bool has_gn = T::kGeneric ? (gn_slice != nullptr && e.gn_part != nullptr) : T::kGn;
  float gn_part[16];
  float* const gn_keep = e.gn_part;
  e.gn_part = has_gn ? gn_part : nullptr;
  uint4 rn[4] = {};
#pragma unroll 1
  for (int c0 = 0; c0 < n_cols; c0 += 32) {
    uint32_t v[32];
    float f[32];
    tmem_ld32(t_row + c0, v);
    if (has_res) {
      if (c0 + 32 < n_cols) {
        load_res32(rn, e, oc0 + c0 + 32, valid);
      } else if (next_res) {
        const bf16* keep = e.res_row;
        e.res_row = next_res;
        load_res32(rn, e, next_oc0, next_valid);
        e.res_row = keep;
      }
    }
    if (GEGLU) {
      uint32_t g[32];
      float gf[32];
      tmem_ld32(t_row + gate_off + c0, g);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) { f[j] = __uint_as_float(v[j]); gf[j] = __uint_as_float(g[j]); }
      if (bias_s) {
        add_bias32_s(f, bias_s + c0);              // value columns (row-permuted weight)
        add_bias32_s(gf, bias_s + gate_off + c0);  // gate columns
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] *= gelu_erf_f(gf[j]);
    } else {
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      if (bias_s) add_bias32_s(f, bias_s + c0);
    }
    if (T::kGeneric) epilogue_store32(f, cv0 + c0, oc0 + c0, valid, e, rr);
    else epilogue_store32_t<T::kAct, T::kGate, T::kRes, T::kGn>(f, cv0 + c0, oc0 + c0, valid, e, rr);
    if (has_res) {
#pragma unroll
      for (int q = 0; q < 4; ++q) rr[q] = rn[q];
    }
    if (has_gn) gn_chunk_accumulate(gn_part, gn_cpg, gn_slice + ((cv0 + c0) / gn_cpg) * 2, lane);
  }
  e.gn_part = gn_keep;
}

// fp32 residual stream (FLUX DiT): out = res + gate * (acc + bias), residual and output both fp32 rows. The DiT adds
// 2 x 57 gated updates to one stream; rounding the stream to bf16 after each of them was the largest single error term
// of the full-size FLUX parity test (DiT output 2.2e-2 max-relative vs the fp32 oracle). The residual of the next
// chunk is loaded before the current one is consumed; n_cols % 32 == 0.
__device__ __forceinline__ void st_global_f32x8(float* p, const float (&f)[32], int o) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(f[o]), "f"(f[o + 1]),
               "f"(f[o + 2]), "f"(f[o + 3]), "f"(f[o + 4]), "f"(f[o + 5]), "f"(f[o + 6]), "f"(f[o + 7])
               : "memory");
}
template <bool GATE>
__device__ __forceinline__ void epilogue_row_f32stream(uint32_t t_row, int n_cols, int oc0, bool valid, const EpiArgs& e,
                                                       const float* bias_s) {
  const float* res = reinterpret_cast<const float*>(e.res_row);
  float* out = static_cast<float*>(e.out) + e.o_off;
  float4 rn[8];
  auto load = [&](int oc) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      rn[q] = valid ? *reinterpret_cast<const float4*>(res + oc + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  load(oc0);
#pragma unroll 1
  for (int c0 = 0; c0 < n_cols; c0 += 32) {
    uint32_t v[32];
    float f[32];
    tmem_ld32(t_row + c0, v);
    float4 rc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) rc[q] = rn[q];
    if (c0 + 32 < n_cols) load(oc0 + c0 + 32);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (bias_s) add_bias32_s(f, bias_s + c0);
    const int oc = oc0 + c0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (GATE) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(e.gate_row + oc + 4 * q));
        f[4 * q] *= g.x; f[4 * q + 1] *= g.y; f[4 * q + 2] *= g.z; f[4 * q + 3] *= g.w;
      }
      f[4 * q] += rc[q].x; f[4 * q + 1] += rc[q].y; f[4 * q + 2] += rc[q].z; f[4 * q + 3] += rc[q].w;
    }
    if (valid) {
#pragma unroll
      for (int q = 0; q < 4; ++q) st_global_f32x8(out + oc + 8 * q, f, 8 * q);
    }
  }
}

// element offset -> residual row pointer for either residual type (EpiArgs carries it as a bf16 pointer)
__device__ __forceinline__ const bf16* res_row_ptr(const bf16* res, long long off, int res_f32) {
  return res_f32 ? reinterpret_cast<const bf16*>(reinterpret_cast<const float*>(res) + off) : res + off;
}

// ---- FLUX q / k heads: per-head RMSNorm + RoPE applied to the fp32 accumulator -------------------------------
// (diffusers: q = to_q(x); q = RMSNorm_128(q) * w; q = apply_rope(q) inside the attention processor that
// infer/omgsr_f_infer_model.py:191-200 reaches.) A thread owns one token row, a head is 128 consecutive accumulator
// columns of that row, so the mean square is thread-local: pass 1 reads the head from TMEM for the sum of squares,
// pass 2 reads it again, scales, rotates pairs (x0, x1) -> (x0 c - x1 s, x1 c + x0 s) and stores bf16. Round 1 ran this
// as a separate in-place kernel over the bf16 projection (114 launches and 450 MB of HBM traffic each per F-1024
// step); fused, the rounding to bf16 happens once, after the rotation.
struct RopeArgs {
  const float* wq;     // [128] RMSNorm weight of the q heads
  const float* wk;     // [128] of the k heads
  const float* cos_t;  // [positions, 128] (each frequency repeated twice), null = no rope mode
  const float* sin_t;
  int q0, q1, k0, k1;  // accumulator column ranges (multiples of 128) holding q / k heads
  int pos_offset;      // position of row 0 of every batch entry
  float eps;
};

__device__ __forceinline__ void epilogue_head_rope(uint32_t t_head, const float* bias_s, const float* __restrict__ w,
                                                   const float* __restrict__ cs, const float* __restrict__ sn,
                                                   bf16* orow, bool valid, float eps) {
  float ssq = 0.f;
#pragma unroll 1
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    float f[32];
    tmem_ld32(t_head + c0, v);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (bias_s) add_bias32_s(f, bias_s + c0);
#pragma unroll
    for (int j = 0; j < 32; ++j) ssq = fmaf(f[j], f[j], ssq);
  }
  const float rs = rsqrtf(ssq * (1.0f / 128.f) + eps);
#pragma unroll 1
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    float f[32];
    tmem_ld32(t_head + c0, v);
    float4 cc[8], ss[8];
    if (valid) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        cc[q] = __ldg(reinterpret_cast<const float4*>(cs + c0) + q);
        ss[q] = __ldg(reinterpret_cast<const float4*>(sn + c0) + q);
      }
    }
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (bias_s) add_bias32_s(f, bias_s + c0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + c0) + q);
      f[4 * q + 0] *= rs * wv.x; f[4 * q + 1] *= rs * wv.y; f[4 * q + 2] *= rs * wv.z; f[4 * q + 3] *= rs * wv.w;
    }
    if (valid) {
      uint4 o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 c0v = cc[2 * q], c1v = cc[2 * q + 1], s0v = ss[2 * q], s1v = ss[2 * q + 1];
        const float* x = f + 8 * q;
        o[q].x = pack_bf16x2(x[0] * c0v.x - x[1] * s0v.x, x[1] * c0v.y + x[0] * s0v.y);
        o[q].y = pack_bf16x2(x[2] * c0v.z - x[3] * s0v.z, x[3] * c0v.w + x[2] * s0v.w);
        o[q].z = pack_bf16x2(x[4] * c1v.x - x[5] * s1v.x, x[5] * c1v.y + x[4] * s1v.y);
        o[q].w = pack_bf16x2(x[6] * c1v.z - x[7] * s1v.z, x[7] * c1v.w + x[6] * s1v.w);
      }
      st_global_256(orow + c0, o[0], o[1]);
      st_global_256(orow + c0 + 16, o[2], o[3]);
    }
  }
}

// 64 accumulator columns that are ALREADY in registers (the rolling-row convolution copies a row out of TMEM, frees the
// TMEM slot at once and only then does the arithmetic and the stores): the chunk body of epilogue_row, twice. Both
// 32-column halves of the residual row were loaded by the caller before it waited for the accumulator.
template <int MODE>
__device__ __forceinline__ void epilogue_regs64(const uint32_t (&v0)[32], const uint32_t (&v1)[32], int cv0, int oc0,
                                                bool valid, EpiArgs& e, const float* bias_s, const uint4 (&rr0)[4],
                                                const uint4 (&rr1)[4], float* gn_slice, int gn_cpg, int lane) {
  using T = EpiTraits<MODE>;
  const bool has_gn = T::kGeneric ? (gn_slice != nullptr && e.gn_part != nullptr) : T::kGn;
  float gn_part[16];
  float* const gn_keep = e.gn_part;
  e.gn_part = has_gn ? gn_part : nullptr;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c0 = 32 * h;
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(h == 0 ? v0[j] : v1[j]);
    if (bias_s) add_bias32_s(f, bias_s + c0);
    if (T::kGeneric) epilogue_store32(f, cv0 + c0, oc0 + c0, valid, e, h == 0 ? rr0 : rr1);
    else epilogue_store32_t<T::kAct, T::kGate, T::kRes, T::kGn>(f, cv0 + c0, oc0 + c0, valid, e, h == 0 ? rr0 : rr1);
    if (has_gn) gn_chunk_accumulate(gn_part, gn_cpg, gn_slice + ((cv0 + c0) / gn_cpg) * 2, lane);
  }
  e.gn_part = gn_keep;
}

// run-time option set -> specialised row routine (once per accumulator row, outside the chunk loop)
template <bool GEGLU>
__device__ __forceinline__ void epilogue_row_m(int mode, uint32_t t_row, int n_cols, int gate_off, int cv0, int oc0,
                                               bool valid, EpiArgs& e, const bf16* next_res, int next_oc0,
                                               bool next_valid, const float* bias_s, uint4 (&rr)[4], float* gn_slice,
                                               int gn_cpg, int lane) {
#define OMGSR_EPI_ROW(M)                                                                                              \
  epilogue_row<M, GEGLU>(t_row, n_cols, gate_off, cv0, oc0, valid, e, next_res, next_oc0, next_valid, bias_s, rr,     \
                         gn_slice, gn_cpg, lane)
  if (GEGLU) {  // only the plain and the generic set occur with GEGLU
    if (mode == EPI_PLAIN) OMGSR_EPI_ROW(EPI_PLAIN);
    else OMGSR_EPI_ROW(EPI_GENERIC);
    return;
  }
  if (mode == EPI_F32S || mode == EPI_F32S_GATE) {
    if (mode == EPI_F32S_GATE) epilogue_row_f32stream<true>(t_row, n_cols, oc0, valid, e, bias_s);
    else epilogue_row_f32stream<false>(t_row, n_cols, oc0, valid, e, bias_s);
    return;
  }
  switch (mode) {
    case EPI_PLAIN: OMGSR_EPI_ROW(EPI_PLAIN); break;
    case EPI_RES: OMGSR_EPI_ROW(EPI_RES); break;
    case EPI_GN: OMGSR_EPI_ROW(EPI_GN); break;
    case EPI_RES_GN: OMGSR_EPI_ROW(EPI_RES_GN); break;
    case EPI_GELU_TANH: OMGSR_EPI_ROW(EPI_GELU_TANH); break;
    case EPI_GATE_RES: OMGSR_EPI_ROW(EPI_GATE_RES); break;
    case EPI_SILU: OMGSR_EPI_ROW(EPI_SILU); break;
    default: OMGSR_EPI_ROW(EPI_GENERIC); break;
  }
#undef OMGSR_EPI_ROW
}

}  // namespace omgsr
