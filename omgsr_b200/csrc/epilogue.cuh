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
};

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

// f: 32 values for accumulator columns [cv, cv+32) of this row; oc0: first OUTPUT channel they map to.
__device__ __forceinline__ void epilogue_store32(float (&f)[32], int cv, int oc0, bool valid, const EpiArgs& e) {
  if (e.act != OMGSR_ACT_NONE && e.act != OMGSR_ACT_GEGLU && cv >= e.act_c0 && cv < e.act_c1) {
    switch (e.act) {
      case OMGSR_ACT_SILU: act32<OMGSR_ACT_SILU>(f); break;
      case OMGSR_ACT_GELU_ERF: act32<OMGSR_ACT_GELU_ERF>(f); break;
      default: act32<OMGSR_ACT_GELU_TANH>(f); break;
    }
  }
  if (!valid) return;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int oc = oc0 + q * 8;
    if (oc >= e.cout_out) break;
    float* ff = f + q * 8;
    if (e.gate_row) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(e.gate_row + oc));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(e.gate_row + oc + 4));
      ff[0] *= g0.x; ff[1] *= g0.y; ff[2] *= g0.z; ff[3] *= g0.w;
      ff[4] *= g1.x; ff[5] *= g1.y; ff[6] *= g1.z; ff[7] *= g1.w;
    }
    if (e.res_row) {
      const uint4 rr = *reinterpret_cast<const uint4*>(e.res_row + oc);
      const float2 r0 = unpack_bf16x2(rr.x), r1 = unpack_bf16x2(rr.y), r2 = unpack_bf16x2(rr.z),
                   r3 = unpack_bf16x2(rr.w);
      ff[0] += r0.x; ff[1] += r0.y; ff[2] += r1.x; ff[3] += r1.y;
      ff[4] += r2.x; ff[5] += r2.y; ff[6] += r3.x; ff[7] += r3.y;
    }
    if (e.out_f32) {
      float* o = static_cast<float*>(e.out) + e.o_off + oc;
      *reinterpret_cast<float4*>(o) = make_float4(ff[0], ff[1], ff[2], ff[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(ff[4], ff[5], ff[6], ff[7]);
    } else {
      uint4 o;
      o.x = pack_bf16x2(ff[0], ff[1]); o.y = pack_bf16x2(ff[2], ff[3]);
      o.z = pack_bf16x2(ff[4], ff[5]); o.w = pack_bf16x2(ff[6], ff[7]);
      *reinterpret_cast<uint4*>(static_cast<bf16*>(e.out) + e.o_off + oc) = o;
    }
  }
}

}  // namespace omgsr
