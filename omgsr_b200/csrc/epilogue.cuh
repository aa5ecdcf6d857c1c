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
  float* gn_sums;  // optional [groups][2] (sum, sumsq) accumulator of THIS row's image: GroupNorm statistics of
  int gn_cpg;      // the stored tensor are produced here instead of by a separate pass over HBM
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

// f: 32 values for accumulator columns [cv, cv+32) of this row; oc0: first OUTPUT channel they map to.
__device__ __forceinline__ void epilogue_store32(float (&f)[32], int cv, int oc0, bool valid, const EpiArgs& e) {
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
      const uint4 rr = *reinterpret_cast<const uint4*>(e.res_row + oc);
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
      if (e.gn_sums) {
        // statistics of the values as stored (bf16-rounded), 8 channels = one or two groups (cpg 4 | >= 8)
        const float2 v0 = unpack_bf16x2(o.x), v1 = unpack_bf16x2(o.y), v2 = unpack_bf16x2(o.z), v3 = unpack_bf16x2(o.w);
        float s0 = valid ? (v0.x + v0.y) + (v1.x + v1.y) : 0.f;
        float q0 = valid ? (v0.x * v0.x + v0.y * v0.y) + (v1.x * v1.x + v1.y * v1.y) : 0.f;
        float s1 = valid ? (v2.x + v2.y) + (v3.x + v3.y) : 0.f;
        float q1 = valid ? (v2.x * v2.x + v2.y * v2.y) + (v3.x * v3.x + v3.y * v3.y) : 0.f;
        if (e.gn_cpg == 4) {
          s0 = warp_sum_all(s0); q0 = warp_sum_all(q0); s1 = warp_sum_all(s1); q1 = warp_sum_all(q1);
          if ((threadIdx.x & 31) == 0) {
            float* g = e.gn_sums + (oc >> 2) * 2;
            atomicAdd(g, s0); atomicAdd(g + 1, q0); atomicAdd(g + 2, s1); atomicAdd(g + 3, q1);
          }
        } else {
          s0 = warp_sum_all(s0 + s1); q0 = warp_sum_all(q0 + q1);
          if ((threadIdx.x & 31) == 0) {
            float* g = e.gn_sums + (oc / e.gn_cpg) * 2;
            atomicAdd(g, s0); atomicAdd(g + 1, q0);
          }
        }
      }
    }
  }
}

}  // namespace omgsr
