// Shared accumulator epilogue of the tcgen05 kernels: 32 fp32 accumulator columns of one row (one thread)
// -> +bias -> activation -> *gate -> +residual -> bf16|fp32 store.
// All per-launch options are branched on ONCE per 32-column chunk (never per element) and bias / gate / residual
// are read with 16-byte loads: the first version branched per element and spent ~2400 SASS instructions per
// chunk, which made the epilogue (not the tensor pipe) the bottleneck of the K=1152 layers (ncu: 12 % tensor).
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
__device__ __forceinline__ void load_res32(uint4 (&rr)[4], const EpiArgs& e, int oc0, bool valid) {
  if (e.res_row && valid) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (oc0 + q * 8 < e.cout_out) rr[q] = *reinterpret_cast<const uint4*>(e.res_row + oc0 + q * 8);
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
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int oc = oc0 + q * 8;
    if (oc >= e.cout_out) break;  // warp-uniform
    float* ff = f + q * 8;
    if (GATE) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(e.gate_row + oc));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(e.gate_row + oc + 4));
      ff[0] *= g0.x; ff[1] *= g0.y; ff[2] *= g0.z; ff[3] *= g0.w;
      ff[4] *= g1.x; ff[5] *= g1.y; ff[6] *= g1.z; ff[7] *= g1.w;
    }
    if (RES) {  // rows that are not stored may add stale registers: harmless, they are never written
      const uint4 rr = rr4[q];
      const float2 r0 = unpack_bf16x2(rr.x), r1 = unpack_bf16x2(rr.y), r2 = unpack_bf16x2(rr.z),
                   r3 = unpack_bf16x2(rr.w);
      ff[0] += r0.x; ff[1] += r0.y; ff[2] += r1.x; ff[3] += r1.y;
      ff[4] += r2.x; ff[5] += r2.y; ff[6] += r3.x; ff[7] += r3.y;
    }
    uint4 o;
    o.x = pack_bf16x2(ff[0], ff[1]); o.y = pack_bf16x2(ff[2], ff[3]);
    o.z = pack_bf16x2(ff[4], ff[5]); o.w = pack_bf16x2(ff[6], ff[7]);
    if (valid) *reinterpret_cast<uint4*>(static_cast<bf16*>(e.out) + e.o_off + oc) = o;
    if (GN) {
      const float2 v0 = unpack_bf16x2(o.x), v1 = unpack_bf16x2(o.y), v2 = unpack_bf16x2(o.z), v3 = unpack_bf16x2(o.w);
      e.gn_part[4 * q + 0] = valid ? (v0.x + v0.y) + (v1.x + v1.y) : 0.f;
      e.gn_part[4 * q + 1] = valid ? (v0.x * v0.x + v0.y * v0.y) + (v1.x * v1.x + v1.y * v1.y) : 0.f;
      e.gn_part[4 * q + 2] = valid ? (v2.x + v2.y) + (v3.x + v3.y) : 0.f;
      e.gn_part[4 * q + 3] = valid ? (v2.x * v2.x + v2.y * v2.y) + (v3.x * v3.x + v3.y * v3.y) : 0.f;
    }
  }
}

// option set -> specialisation id (0 = generic). Computed once per thread, warp-uniform.
enum : int { EPI_GENERIC = 0, EPI_PLAIN, EPI_RES, EPI_GN, EPI_RES_GN, EPI_GELU_TANH, EPI_GATE_RES, EPI_SILU };
__device__ __forceinline__ int epilogue_mode(int act, bool gate, bool res, bool out_f32, bool gn) {
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

// flush this warp's slice (n floats) into the global [groups][2] accumulator of one image and clear it
__device__ __forceinline__ void gn_flush(float* acc, int n, float* dst, int lane) {
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    const float v = acc[i];
    if (v != 0.f) atomicAdd(dst + i, v);
    acc[i] = 0.f;
  }
  __syncwarp();
}

}  // namespace omgsr
