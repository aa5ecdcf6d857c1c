// HBM-bound glue kernels of the OMGSR inference path: GroupNorm(+SiLU), LayerNorm(+AdaLN modulation),
// row softmax, FLUX RMSNorm+RoPE, thin convolutions at the NCHW module boundary, latent-sized
// elementwise steps, 2x2 pack/unpack, Gaussian tile blend, layout changes.
// All of them are vectorised (16-byte accesses), warp-reduced and sized in multiples of the SM count.
#include "common.cuh"
#include "../../include/omgsr_b200.h"

namespace omgsr {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
  return o;
}
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  const uint4 u = ld_stream(p);
  return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
}

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics: stats[b, g] = (sum, sumsq). Thread owns one 8-channel vector column and walks
// pixels. Bit-reproducible: the per-thread channel sums are folded per group in a FIXED order inside the block
// (shared-memory table, no atomics), every block writes its (sum, sumsq) per group to its own scratch slot
// [image][block][2G], and gn_finalize_kernel adds the slots in a fixed order.
// ------------------------------------------------------------------------------------------------
__global__ void gn_stats_kernel(const bf16* __restrict__ x, long long ldx, float* __restrict__ scratch, long long HW,
                                int C, int G, int vpr, int rows_per_iter, int blocks_per_img) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float tab[];  // [2][rows_per_iter][C]: per-(row lane, channel) sum / sum of squares
  const int b = blockIdx.x / blocks_per_img;
  const int blk = blockIdx.x % blocks_per_img;
  const int v = threadIdx.x % vpr;
  const int r0 = threadIdx.x / vpr;
  if (threadIdx.x < vpr * rows_per_iter) {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    const bf16* base = x + static_cast<long long>(b) * HW * ldx + v * 8;
    const long long step = static_cast<long long>(blocks_per_img) * rows_per_iter;
    long long r = static_cast<long long>(blk) * rows_per_iter + r0;
    for (; r + 3 * step < HW; r += 4 * step) {  // four independent 16-byte loads in flight per thread
      float f0[8], f1[8], f2[8], f3[8];
      const uint4 u0 = ld_stream(base + r * ldx), u1 = ld_stream(base + (r + step) * ldx),
                  u2 = ld_stream(base + (r + 2 * step) * ldx), u3 = ld_stream(base + (r + 3 * step) * ldx);
      unpack8(u0, f0); unpack8(u1, f1); unpack8(u2, f2); unpack8(u3, f3);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += (f0[j] + f1[j]) + (f2[j] + f3[j]);
        q[j] += (f0[j] * f0[j] + f1[j] * f1[j]) + (f2[j] * f2[j] + f3[j] * f3[j]);
      }
    }
    for (; r < HW; r += step) {
      float f[8];
      unpack8(ld_stream(base + r * ldx), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] += f[j] * f[j]; }
    }
    float* ts = tab + (static_cast<long long>(r0) * C + v * 8);
    float* tq = ts + static_cast<long long>(rows_per_iter) * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) { ts[j] = s[j]; tq[j] = q[j]; }
  }
  __syncthreads();
  const int cpg = C / G;
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) {
    const int g = i >> 1;
    const float* t = tab + (i & 1) * static_cast<long long>(rows_per_iter) * C + g * cpg;
    float acc = 0.f;
    for (int r = 0; r < rows_per_iter; ++r)
      for (int c = 0; c < cpg; ++c) acc += t[r * C + c];
    scratch[(static_cast<long long>(b) * blocks_per_img + blk) * 2 * G + i] = acc;
  }
}

// out[b][i] = sum over (launch l, slot s) of scratch[((l * B + b) * slots + s) * n + i]. Grid (image, value group): a
// block owns V of the n values of one image, the (l, s) range is split over blockDim.x / V thread groups in a fixed
// interleaving, each thread adds its share in a fixed order (four independent chains) and the groups are combined in
// index order: same bits every run. (One block per image read 300 KB of mostly-zero slots with 16 blocks in flight:
// 20 us per call, 3 % of the S-512 step; 8 x more blocks and a quarter of the chain length bring it to ~3 us.)
__global__ void gn_finalize_kernel(const float* __restrict__ scratch, int launches, int B, int slots, int n, int V,
                                   float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float part[];  // [P][V]
  const int b = blockIdx.x, v0 = blockIdx.y * V;
  const int P = blockDim.x / V;
  const int i = threadIdx.x % V, pid = threadIdx.x / V;
  const long long total = static_cast<long long>(launches) * slots;
  if (pid < P) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    auto at = [&](long long k) {
      const long long l = k / slots, sl = k % slots;
      return __ldg(scratch + ((l * B + b) * slots + sl) * n + v0 + i);
    };
    long long k = pid;
    for (; k + 3LL * P < total; k += 4LL * P) {
      const float x0 = at(k), x1 = at(k + P), x2 = at(k + 2LL * P), x3 = at(k + 3LL * P);
      a0 += x0; a1 += x1; a2 += x2; a3 += x3;
    }
    for (; k < total; k += P) a0 += at(k);
    part[pid * V + i] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
  // fixed-order tree over the P partial sums of each value (P is a power of two: blockDim.x and V are)
  for (int stride = P >> 1; stride > 0; stride >>= 1) {
    if (pid < stride) part[pid * V + i] += part[(pid + stride) * V + i];
    __syncthreads();
  }
  if (threadIdx.x < V) out[static_cast<long long>(b) * n + v0 + threadIdx.x] = part[threadIdx.x];
}

int launch_gn_finalize(const float* scratch, int launches, int B, int slots, int n, float* out, cudaStream_t st) {
  if (n <= 0 || n > 1024 || (n & 1)) return set_error(OMGSR_ERR_SHAPE, "gn_finalize: 2 * groups = %d out of range", n);
  const int V = (n % 8 == 0) ? 8 : 2;
  const int threads = 512;
  launch_pdl(gn_finalize_kernel, dim3(dim3(B, n / V)), dim3(threads), static_cast<size_t>(threads) * sizeof(float), st, scratch, launches, B,
                                                                                                  slots, n, V, out);
  return check_launch("gn_finalize_kernel");
}

// GroupNorm apply, stage 1: per-(image, channel) affine coefficients  y = x * a + b  with
//   a = rstd_g * gamma_c,  b = beta_c - mean_g * a       (tiny: B*C entries)
// stats_mode 0: stats[b] = (sum, sumsq) of this very tensor; 1: stats[b % stats_mod] = (mean, var) shared by the
// tiles of one image (cross-tile GroupNorm of the tiled VAE).
__global__ void gn_coef_kernel(const float* __restrict__ stats, const float* __restrict__ gamma,
                               const float* __restrict__ beta, float2* __restrict__ coef, int B, long long HW, int C,
                               int G, float eps, int stats_mode, int stats_mod) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i % C;
  const int cpg = C / G, g = c / cpg;
  const float* st = stats + static_cast<long long>(stats_mode ? b % stats_mod : b) * 2 * G;
  float mean, var;
  if (stats_mode) {
    mean = st[2 * g];
    var = st[2 * g + 1];
  } else {
    const float inv_n = 1.0f / (static_cast<float>(HW) * cpg);
    mean = st[2 * g] * inv_n;
    var = fmaxf(st[2 * g + 1] * inv_n - mean * mean, 0.f);
  }
  const float a = rsqrtf(var + eps) * gamma[c];
  coef[i] = make_float2(a, beta[c] - mean * a);
}

// stage 2: streaming y = silu?(x * a + b). A thread owns one 8-channel column for its whole pixel range, so its
// 16 coefficients are loaded once; per element the loop is one FMA (+ SiLU).
template <bool SILU>
__global__ void gn_apply_kernel(const bf16* __restrict__ x, long long ldx, bf16* __restrict__ y, long long ldy,
                                const float2* __restrict__ coef, long long HW, int C, int vpr, int rows_per_iter,
                                int blocks_per_img) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x / blocks_per_img;
  const int blk = blockIdx.x % blocks_per_img;
  if (threadIdx.x >= vpr * rows_per_iter) return;
  const int v = threadIdx.x % vpr;
  const int r0 = threadIdx.x / vpr;
  float ca[8], cb[8];
  const float2* cf = coef + static_cast<long long>(b) * C + v * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 t = __ldg(cf + j);
    ca[j] = SILU ? 0.5f * t.x : t.x;  // SiLU: halved affine, silu(2h) = h + h tanh(h)
    cb[j] = SILU ? 0.5f * t.y : t.y;
  }
  const bf16* xb = x + static_cast<long long>(b) * HW * ldx + v * 8;
  bf16* yb = y + static_cast<long long>(b) * HW * ldy + v * 8;
  const long long step = static_cast<long long>(blocks_per_img) * rows_per_iter;
  long long r = static_cast<long long>(blk) * rows_per_iter + r0;
  // four pixels in flight per thread (ncu: 72 % of the HBM peak with two, long-scoreboard bound)
  for (; r + 3 * step < HW; r += 4 * step) {
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = ld_stream(xb + (r + i * step) * ldx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float f[8];
      unpack8(u[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float o = fmaf(f[j], ca[j], cb[j]);
        f[j] = SILU ? silu_half(o) : o;
      }
      *reinterpret_cast<uint4*>(yb + (r + i * step) * ldy) = pack8(f);
    }
  }
  for (; r < HW; r += step) {
    float f0[8];
    unpack8(ld_stream(xb + r * ldx), f0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float o0 = fmaf(f0[j], ca[j], cb[j]);
      f0[j] = SILU ? silu_half(o0) : o0;
    }
    *reinterpret_cast<uint4*>(yb + r * ldy) = pack8(f0);
  }
}

// cross-tile GroupNorm statistics of the tiled VAE (GroupNormParam.summary, infer/vaehook.py:489-508):
// per entry e = (tile, image b = e % n_img): mean_e, biased var_e; result[b] = sum_e w_e * (mean_e, var_e)
// with w_e = pixels_e / sum of pixels over the tiles of image b.
__global__ void gn_combine_kernel(const float* __restrict__ sums, const float* __restrict__ pixels, int entries,
                                  int n_img, int G, int cpg, float* __restrict__ out, float* __restrict__ wsum_out) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * G) return;
  const int b = i / G, g = i % G;
  float wsum = 0.f, m = 0.f, v = 0.f;
  for (int e = b; e < entries; e += n_img) {
    if (!(pixels[e] > 0.f)) continue;  // zero-weight entry (a shadow rank's duplicate tile): contributes nothing
    const float n = pixels[e] * cpg;
    const float mean = sums[(static_cast<long long>(e) * G + g) * 2] / n;
    const float var = fmaxf(sums[(static_cast<long long>(e) * G + g) * 2 + 1] / n - mean * mean, 0.f);
    wsum += pixels[e];
    m += pixels[e] * mean;
    v += pixels[e] * var;
  }
  if (wsum_out) {  // raw numerators: the caller sums them over ranks (tile-parallel VAE) and divides afterwards
    out[2 * i] = m;
    out[2 * i + 1] = v;
    if (g == 0) wsum_out[b] = wsum;
  } else {
    out[2 * i] = m / wsum;
    out[2 * i + 1] = v / wsum;
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row; the row is cached in registers (C <= 4096).
// ------------------------------------------------------------------------------------------------
// 8 consecutive input elements as floats: bf16 rows (16 bytes) or fp32 rows (the FLUX fp32 residual stream, 32 bytes)
__device__ __forceinline__ void ld8_stream(const bf16* p, float (&f)[8]) { unpack8(ld_stream(p), f); }
__device__ __forceinline__ void ld8_stream(const float* p, float (&f)[8]) {
  const float4 a = ld_stream_f4(reinterpret_cast<const float4*>(p)), b = ld_stream_f4(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

template <int VPL, typename TIn>  // 8-element vectors per lane
__global__ void layernorm_kernel(const TIn* __restrict__ x, bf16* __restrict__ y, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, long long rows, int C, float eps,
                                 long long rows_per_batch, long long gb_stride, long long ldx, long long x_batch,
                                 int rms) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int nvec = C / 8;
  for (long long row = warp_global; row < rows; row += nwarps) {
    // input rows may be batch-strided (a token range inside a longer joint sequence); output is dense
    const TIn* xr = (rows_per_batch > 0 && x_batch > 0)
                        ? x + (row / rows_per_batch) * x_batch + (row % rows_per_batch) * ldx
                        : x + row * ldx;
    float f[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        ld8_stream(xr + v * 8, f[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += f[i][j];
      }
    }
    const float mean = rms ? 0.f : warp_sum(s) / C;  // rms (T5LayerNorm): no mean subtraction, variance about zero
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = f[i][j] - mean; q += d * d; }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
    const long long gb = (rows_per_batch > 0 ? (row / rows_per_batch) : 0) * gb_stride;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        float gm[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f}, bt[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (gamma) {  // 16-byte loads of the affine / modulation vectors (was one scalar load per element)
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + gb + v * 8));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + gb + v * 8 + 4));
          gm[0] = g0.x; gm[1] = g0.y; gm[2] = g0.z; gm[3] = g0.w; gm[4] = g1.x; gm[5] = g1.y; gm[6] = g1.z; gm[7] = g1.w;
        }
        if (beta) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + gb + v * 8));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + gb + v * 8 + 4));
          bt[0] = b0.x; bt[1] = b0.y; bt[2] = b0.z; bt[3] = b0.w; bt[4] = b1.x; bt[5] = b1.y; bt[6] = b1.z; bt[7] = b1.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) f[i][j] = fmaf((f[i][j] - mean) * rstd, gm[j], bt[j]);
        *reinterpret_cast<uint4*>(y + row * C + v * 8) = pack8(f[i]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Row softmax of fp32 scores -> bf16 probabilities (one block per row, scores re-read from L2).
// ------------------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(const float* __restrict__ S, long long ld_s, bf16* __restrict__ P, long long ld_p,
                                    int n, int n_pad, float scale_log2e, int causal_m, const float* __restrict__ bias,
                                    long long bias_rows, long long ld_b) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[32];
  const float* s = S + blockIdx.x * ld_s;
  bf16* p = P + blockIdx.x * ld_p;
  // additive bias (T5 relative position bias [heads * queries, keys], shared by the batch), in natural-log units
  const float* bz = bias ? bias + (blockIdx.x % bias_rows) * ld_b : nullptr;
  // causal (text encoders): rows are [.., M queries]; query i sees keys 0..i, the rest of the row is written as zeros
  if (causal_m > 0) n = min(n, static_cast<int>(blockIdx.x % causal_m) + 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  auto val = [&](int j) { return bz ? fmaf(s[j], scale_log2e, bz[j] * 1.4426950408889634f) : s[j] * scale_log2e; };
  float m = -INFINITY;
  for (int j = threadIdx.x; j < n; j += blockDim.x) m = fmaxf(m, val(j));
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = (lane < nw) ? red[lane] : -INFINITY;
  m = warp_max(m);
  __syncthreads();
  float sum = 0.f;
  for (int j = threadIdx.x; j < n; j += blockDim.x) sum += exp2f(val(j) - m);
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = (lane < nw) ? red[lane] : 0.f;
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int j = threadIdx.x; j < n_pad; j += blockDim.x)
    p[j] = __float2bfloat16(j < n ? exp2f(val(j) - m) * inv : 0.f);
}

// Register-resident variant for rows of up to 256 * 4 * V4 scores (V4 <= 32: 32768): the row is read from HBM exactly once (V4 float4
// per thread, all in flight), reduced, exponentiated and written as bf16 (8-byte stores). The three-pass kernel above
// ran at 1.3 TB/s on the VAE mid-block attention (4096 x 4096 scores per image).
template <int V4>
__global__ void __launch_bounds__(256) softmax_rows_reg_kernel(const float* __restrict__ S, long long ld_s,
                                                               bf16* __restrict__ P, long long ld_p, int n, int n_pad,
                                                               float scale_log2e) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[2][8];
  const float4* s4 = reinterpret_cast<const float4*>(S + blockIdx.x * ld_s);
  uint2* p2 = reinterpret_cast<uint2*>(P + blockIdx.x * ld_p);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n4 = n_pad >> 2;
  float4 v[V4];
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const int j = threadIdx.x + 256 * i;
    v[i] = j < n4 ? ld_stream_f4(s4 + j) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  }
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const int c = (threadIdx.x + 256 * i) * 4;
    if (c + 0 >= n) v[i].x = -INFINITY;
    if (c + 1 >= n) v[i].y = -INFINITY;
    if (c + 2 >= n) v[i].z = -INFINITY;
    if (c + 3 >= n) v[i].w = -INFINITY;
    m = fmaxf(m, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
  }
  m = warp_max(m);
  if (lane == 0) red[0][warp] = m;
  __syncthreads();
  m = red[0][lane & 7];
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
  const float mo = m * scale_log2e;
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    v[i].x = exp2f(fmaf(v[i].x, scale_log2e, -mo)); v[i].y = exp2f(fmaf(v[i].y, scale_log2e, -mo));
    v[i].z = exp2f(fmaf(v[i].z, scale_log2e, -mo)); v[i].w = exp2f(fmaf(v[i].w, scale_log2e, -mo));
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  sum = warp_sum(sum);
  if (lane == 0) red[1][warp] = sum;
  __syncthreads();
  sum = red[1][lane & 7];
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  sum += __shfl_xor_sync(0xffffffffu, sum, 4);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < n4) p2[j] = make_uint2(pack_bf16x2(v[i].x * inv, v[i].y * inv), pack_bf16x2(v[i].z * inv, v[i].w * inv));
  }
}

// ------------------------------------------------------------------------------------------------
// FLUX: per-head RMSNorm (d = 128) + rotary embedding on q and k, in place. One warp per (row, head):
// lane owns 4 consecutive channels = 2 rotation pairs. Heads [0, heads_per_w) use weight[0] (q),
// the next heads_per_w use weight[1] (k), so q and k of a fused qkv row are handled by one launch.
// ------------------------------------------------------------------------------------------------
__global__ void rmsnorm_rope_kernel(bf16* __restrict__ x, long long ld, long long batch_stride,
                                    const float* __restrict__ weight, const float* __restrict__ cos_t,
                                    const float* __restrict__ sin_t, long long rows_per_batch, int batch, int heads,
                                    int heads_per_w, int pos_offset, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long wg = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nw = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const long long total = rows_per_batch * batch * heads;
  for (long long it = wg; it < total; it += nw) {
    const int h = static_cast<int>(it % heads);
    const long long br = it / heads;
    const long long r = br % rows_per_batch, b = br / rows_per_batch;
    bf16* px = x + b * batch_stride + r * ld + h * 128 + lane * 4;
    const uint2 u = *reinterpret_cast<const uint2*>(px);
    const float2 a = unpack_bf16x2(u.x), c2 = unpack_bf16x2(u.y);
    float ss = a.x * a.x + a.y * a.y + c2.x * c2.x + c2.y * c2.y;
    ss = warp_sum(ss);
    const float rs = rsqrtf(ss * (1.0f / 128.f) + eps);
    const float4 w = *reinterpret_cast<const float4*>(weight + (h / heads_per_w) * 128 + lane * 4);
    const float v0 = a.x * rs * w.x, v1 = a.y * rs * w.y, v2 = c2.x * rs * w.z, v3 = c2.y * rs * w.w;
    const long long pos = pos_offset + r;
    const float4 c = *reinterpret_cast<const float4*>(cos_t + pos * 128 + lane * 4);
    const float4 s = *reinterpret_cast<const float4*>(sin_t + pos * 128 + lane * 4);
    uint2 o;
    o.x = pack_bf16x2(v0 * c.x - v1 * s.x, v1 * c.y + v0 * s.y);
    o.y = pack_bf16x2(v2 * c.z - v3 * s.z, v3 * c.w + v2 * s.w);
    *reinterpret_cast<uint2*>(px) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// conv_in: NCHW (bf16|f32) -> NHWC bf16, Cin <= 16, 1x1 or 3x3 'same'. Thread = (pixel, 8 out channels).
// ------------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void conv_in_kernel(const TIn* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                               bf16* __restrict__ y, int B, int Cin, int H, int W, int Cout, int ks, long long ldy,
                               int cg, long long total) {
  pdl_launch_dependents();
  pdl_wait();
  const int pad = ks / 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % cg);
    const long long pix = i / cg;
    const int xx = static_cast<int>(pix % W);
    const int yy = static_cast<int>((pix / W) % H);
    const int b = static_cast<int>(pix / (static_cast<long long>(W) * H));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias ? __ldg(bias + g * 8 + j) : 0.f;
    for (int r = 0; r < ks; ++r) {
      const int iy = yy + r - pad;
      if (iy < 0 || iy >= H) continue;
      for (int s = 0; s < ks; ++s) {
        const int ix = xx + s - pad;
        if (ix < 0 || ix >= W) continue;
        for (int ci = 0; ci < Cin; ++ci) {
          const float v = static_cast<float>(x[((static_cast<long long>(b) * Cin + ci) * H + iy) * W + ix]);
          const float* wp = w + (static_cast<long long>((r * ks + s) * Cin + ci)) * Cout + g * 8;
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
          acc[0] += v * w0.x; acc[1] += v * w0.y; acc[2] += v * w0.z; acc[3] += v * w0.w;
          acc[4] += v * w1.x; acc[5] += v * w1.y; acc[6] += v * w1.z; acc[7] += v * w1.w;
        }
      }
    }
    *reinterpret_cast<uint4*>(y + pix * ldy + g * 8) = pack8(acc);
  }
}

// im2col for the thin-input 3x3 convs (Cin <= 16): NCHW (bf16|f32) -> patch rows [B*H*W, Kpad] bf16 with
// k = (r*3 + s)*Cin + ci, zero padded to Kpad (multiple of 8; the GEMM zero-fills its k-blocks beyond Kpad).
// One thread writes one 16-byte chunk (8 k values) of one pixel.
template <typename TIn>
__global__ void im2col3x3_kernel(const TIn* __restrict__ x, bf16* __restrict__ y, int Cin, int H, int W, int Kpad,
                                 long long total) {
  pdl_launch_dependents();
  pdl_wait();
  const int chunks = Kpad / 8;
  const int K = 9 * Cin;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % chunks);
    const long long pix = i / chunks;
    const int xx = static_cast<int>(pix % W);
    const int yy = static_cast<int>((pix / W) % H);
    const long long b = pix / (static_cast<long long>(W) * H);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = ch * 8 + j;
      float v = 0.f;
      if (k < K) {
        const int tap = k / Cin, ci = k - tap * Cin;
        const int iy = yy + tap / 3 - 1, ix = xx + tap % 3 - 1;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = static_cast<float>(x[((b * Cin + ci) * H + iy) * W + ix]);
      }
      f[j] = v;
    }
    *reinterpret_cast<uint4*>(y + pix * Kpad + ch * 8) = pack8(f);
  }
}

// conv_out: NHWC bf16 -> NCHW, Cout <= 4, 3x3 p1. Weights (bf16 [Cout, 9*Cin]) staged in smem as fp32.
template <int COUT, typename TOut>
__global__ void conv_out_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const float* __restrict__ bias,
                                TOut* __restrict__ y, int B, int H, int W, int Cin, int clamp, long long ldx,
                                long long total) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float ws[];  // [COUT][9*Cin]
  const int K = 9 * Cin;
  for (int i = threadIdx.x; i < COUT * K; i += blockDim.x) ws[i] = __bfloat162float(w[i]);
  __syncthreads();
  for (long long pix = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; pix < total;
       pix += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xx = static_cast<int>(pix % W);
    const int yy = static_cast<int>((pix / W) % H);
    const int b = static_cast<int>(pix / (static_cast<long long>(W) * H));
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = bias ? __ldg(bias + co) : 0.f;
    for (int r = 0; r < 3; ++r) {
      const int iy = yy + r - 1;
      if (iy < 0 || iy >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int ix = xx + s - 1;
        if (ix < 0 || ix >= W) continue;
        const bf16* px = x + ((static_cast<long long>(b) * H + iy) * W + ix) * ldx;
        const float* wk = ws + (r * 3 + s) * Cin;
        for (int c = 0; c < Cin; c += 8) {
          float f[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(px + c)), f);
#pragma unroll
          for (int co = 0; co < COUT; ++co) {
            const float4 w0 = *reinterpret_cast<const float4*>(wk + co * K + c);
            const float4 w1 = *reinterpret_cast<const float4*>(wk + co * K + c + 4);
            acc[co] += f[0] * w0.x + f[1] * w0.y + f[2] * w0.z + f[3] * w0.w + f[4] * w1.x + f[5] * w1.y +
                       f[6] * w1.z + f[7] * w1.w;
          }
        }
      }
    }
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      float v = acc[co];
      if (clamp) v = fminf(fmaxf(v, -1.f), 1.f);
      y[((static_cast<long long>(b) * COUT + co) * H + yy) * W + xx] = static_cast<TOut>(v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// latent-sized elementwise steps
// ------------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void posterior_kernel(const float* __restrict__ mom, long long ldm, const float* __restrict__ eps,
                                 TOut* __restrict__ z, int Cz, long long hw, long long total, float shift,
                                 float scale) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i % hw;
    const int c = static_cast<int>((i / hw) % Cz);
    const long long b = i / (hw * Cz);
    const float* m = mom + (b * hw + p) * ldm;
    const float mean = m[c];
    const float logvar = fminf(fmaxf(m[Cz + c], -30.f), 20.f);
    const float v = mean + __expf(0.5f * logvar) * eps[i];
    z[i] = static_cast<TOut>((v - shift) * scale);
  }
}

template <typename T>
__global__ void axpby_kernel(const T* __restrict__ x, const T* __restrict__ y, T* __restrict__ out, long long n,
                             float a, float b, float c) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v = a * static_cast<float>(x[i]) + c;
    if (y) v += b * static_cast<float>(y[i]);
    out[i] = static_cast<T>(v);
  }
}

template <typename TZ>
__global__ void x0_unscale_kernel(const TZ* __restrict__ z, const float* __restrict__ eh, long long lde,
                                  bf16* __restrict__ out, const float* __restrict__ Wpq, const float* __restrict__ bpq,
                                  int Cz, long long hw, long long total_pix, float s1, float inv_s2, float inv_scale,
                                  float shift) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total_pix;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / hw, p = i % hw;
    float z0[16];
    for (int c = 0; c < Cz; ++c) {
      const float zz = static_cast<float>(z[(b * Cz + c) * hw + p]);
      z0[c] = (zz - s1 * eh[i * lde + c]) * inv_s2 * inv_scale + shift;
    }
    for (int co = 0; co < Cz; ++co) {
      float v;
      if (Wpq) {
        v = bpq ? bpq[co] : 0.f;
        for (int c = 0; c < Cz; ++c) v += Wpq[co * Cz + c] * z0[c];
      } else {
        v = z0[co];
      }
      out[(b * Cz + co) * hw + p] = __float2bfloat16(v);
    }
  }
}

// FLUX 2x2 pack: tokens[b, (y/2)(w/2) + x/2, c*4 + dy*2 + dx] = z[b, c, y, x]
__global__ void pack_kernel(const bf16* __restrict__ z, bf16* __restrict__ tok, int C, int h, int w, long long ld,
                            long long total, int unpack) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % w);
    const int y = static_cast<int>((i / w) % h);
    const int c = static_cast<int>((i / (static_cast<long long>(w) * h)) % C);
    const long long b = i / (static_cast<long long>(w) * h * C);
    const long long t = b * (h / 2) * (w / 2) + static_cast<long long>(y / 2) * (w / 2) + x / 2;
    bf16* tp = tok + t * ld + c * 4 + (y & 1) * 2 + (x & 1);
    if (unpack) const_cast<bf16*>(z)[i] = *tp; else *tp = z[i];
  }
}

// out[b,c,y,x] = sum_t w[y-oy_t, x-ox_t] * tile_t[b,c,y-oy_t,x-ox_t] / sum_t w[...]   (generic strides)
struct BlendStrides {
  long long t_t, t_b, t_c, t_y, t_x;  // tile tensor strides (elements)
  long long o_b, o_c, o_y, o_x;       // output strides
};
template <typename T>
__global__ void tile_blend_kernel(const T* __restrict__ tiles, const float* __restrict__ wts,
                                  const int* __restrict__ ofs, int ntiles, float* __restrict__ out, int B, int C, int h,
                                  int w, int T_, long long total, BlendStrides st) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int x = static_cast<int>((i / C) % w);
    const int y = static_cast<int>((i / (static_cast<long long>(C) * w)) % h);
    const int b = static_cast<int>(i / (static_cast<long long>(C) * w * h));
    float num = 0.f, den = 0.f;
    for (int t = 0; t < ntiles; ++t) {
      const int tx = x - ofs[2 * t], ty = y - ofs[2 * t + 1];
      if (tx < 0 || ty < 0 || tx >= T_ || ty >= T_) continue;
      const float wv = wts[ty * T_ + tx];
      num += wv * static_cast<float>(tiles[t * st.t_t + b * st.t_b + c * st.t_c + ty * st.t_y + tx * st.t_x]);
      den += wv;
    }
    out[b * st.o_b + c * st.o_c + y * st.o_y + x * st.o_x] = num / den;
  }
}

// layout changes through a 32x32 smem tile
template <typename TIn>
__global__ void nchw_to_nhwc_kernel(const TIn* __restrict__ x, bf16* __restrict__ y, int C, long long HW, long long ldy) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float tile[32][33];
  const long long b = blockIdx.z;
  const long long p0 = blockIdx.x * 32LL;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j;
    const long long p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < HW) ? static_cast<float>(x[(b * C + c) * HW + p]) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long p = p0 + j;
    const int c = c0 + threadIdx.x;
    if (c < C && p < HW) y[(b * HW + p) * ldy + c] = __float2bfloat16(tile[threadIdx.x][j]);
  }
}
template <typename TOut>
__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ x, long long ldx, TOut* __restrict__ y, int C,
                                    long long HW, int clamp) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float tile[32][33];
  const long long b = blockIdx.z;
  const long long p0 = blockIdx.x * 32LL;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long p = p0 + j;
    const int c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < HW) ? __bfloat162float(x[(b * HW + p) * ldx + c]) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j;
    const long long p = p0 + threadIdx.x;
    if (c < C && p < HW) {
      float v = tile[threadIdx.x][j];
      if (clamp) v = fminf(fmaxf(v, -1.f), 1.f);
      y[(b * C + c) * HW + p] = static_cast<TOut>(v);
    }
  }
}

static int grid_for(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace omgsr

using namespace omgsr;

extern "C" long long omgsr_gn_scratch_floats(int B, int G, int launches) {
  if (B <= 0 || G <= 0 || launches <= 0) return 0;
  const long long conv = static_cast<long long>(launches) * B * gn_slots_per_image() * 2 * G;
  const long long stats = (static_cast<long long>(sm_count()) * 8 + B) * 2 * G;
  return conv > stats ? conv : stats;
}

extern "C" int omgsr_groupnorm_stats(const void* x, long long ldx, float* stats, int B, long long HW, int C, int G,
                                     float* scratch, long long scratch_floats, void* stream) {
  if (C % 8 || C % G || G > 512 || C > 4096) return set_error(OMGSR_ERR_SHAPE, "groupnorm_stats: C=%d G=%d", C, G);
  if (ldx % 8) return set_error(OMGSR_ERR_ALIGN, "groupnorm_stats: ldx %% 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vpr = C / 8;
  int rows_per_iter = 256 / vpr;
  if (rows_per_iter < 1) rows_per_iter = 1;
  const int threads = ((vpr * rows_per_iter + 31) / 32) * 32;
  long long bpi = (static_cast<long long>(sm_count()) * 8 + B - 1) / B;
  const long long max_bpi = (HW + rows_per_iter - 1) / rows_per_iter;
  if (bpi > max_bpi) bpi = max_bpi;
  if (bpi < 1) bpi = 1;
  if (!scratch || scratch_floats < bpi * B * 2 * G)
    return set_error(OMGSR_ERR_SHAPE, "groupnorm_stats: scratch needs %lld floats (omgsr_gn_scratch_floats), got %lld",
                     bpi * B * 2 * G, scratch ? scratch_floats : 0LL);
  const size_t smem = 2ull * rows_per_iter * C * sizeof(float);  // <= 32 KB
  launch_pdl(gn_stats_kernel, dim3(static_cast<unsigned>(bpi * B)), dim3(threads), smem, st, 
      static_cast<const bf16*>(x), ldx, scratch, HW, C, G, vpr, rows_per_iter, static_cast<int>(bpi));
  int rc = check_launch("gn_stats_kernel");
  if (rc) return rc;
  return launch_gn_finalize(scratch, 1, B, static_cast<int>(bpi), 2 * G, stats, st);
}

extern "C" int omgsr_groupnorm_apply(const void* x, long long ldx, void* y, long long ldy, const float* stats,
                                     const float* gamma, const float* beta, void* coef_scratch, int B, long long HW,
                                     int C, int G, float eps, int silu, int stats_mode, int stats_mod, void* stream) {
  if (C % 8 || C % G || C / 8 > 1024) return set_error(OMGSR_ERR_SHAPE, "groupnorm_apply: C=%d G=%d", C, G);
  if ((ldx | ldy) % 8) return set_error(OMGSR_ERR_ALIGN, "groupnorm_apply: ld %% 8");
  if (stats_mode && stats_mod <= 0) return set_error(OMGSR_ERR_SHAPE, "groupnorm_apply: stats_mod must be > 0");
  if (!coef_scratch) return set_error(OMGSR_ERR_SHAPE, "groupnorm_apply: coef_scratch (B*C float2) is required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float2* coef = static_cast<float2*>(coef_scratch);
  launch_pdl(gn_coef_kernel, dim3((B * C + 255) / 256), dim3(256), 0, st, stats, gamma, beta, coef, B, HW, C, G, eps, stats_mode, stats_mod);
  int rc = check_launch("gn_coef_kernel");
  if (rc) return rc;
  const int vpr = C / 8;
  int rows_per_iter = 256 / vpr;
  if (rows_per_iter < 1) rows_per_iter = 1;
  const int threads = ((vpr * rows_per_iter + 31) / 32) * 32;
  long long bpi = (static_cast<long long>(sm_count()) * 16 + B - 1) / B;
  const long long max_bpi = (HW + 2LL * rows_per_iter - 1) / (2LL * rows_per_iter);
  if (bpi > max_bpi) bpi = max_bpi;
  if (bpi < 1) bpi = 1;
  const unsigned grid = static_cast<unsigned>(bpi * B);
  if (silu)
    launch_pdl(gn_apply_kernel<true>, dim3(grid), dim3(threads), 0, st, static_cast<const bf16*>(x), ldx, static_cast<bf16*>(y), ldy, coef,
                                                    HW, C, vpr, rows_per_iter, static_cast<int>(bpi));
  else
    launch_pdl(gn_apply_kernel<false>, dim3(grid), dim3(threads), 0, st, static_cast<const bf16*>(x), ldx, static_cast<bf16*>(y), ldy, coef,
                                                     HW, C, vpr, rows_per_iter, static_cast<int>(bpi));
  return check_launch("gn_apply_kernel");
}

extern "C" int omgsr_groupnorm_coef(const float* stats, const float* gamma, const float* beta, void* coef, int B,
                                   long long HW, int C, int G, float eps, int stats_mode, int stats_mod, void* stream) {
  if (C % 8 || C % G) return set_error(OMGSR_ERR_SHAPE, "groupnorm_coef: C=%d G=%d", C, G);
  if (stats_mode && stats_mod <= 0) return set_error(OMGSR_ERR_SHAPE, "groupnorm_coef: stats_mod must be > 0");
  if (!coef || (reinterpret_cast<uintptr_t>(coef) & 15)) return set_error(OMGSR_ERR_ALIGN, "groupnorm_coef: coef must be 16-byte aligned");
  launch_pdl(gn_coef_kernel, dim3((B * C + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      stats, gamma, beta, static_cast<float2*>(coef), B, HW, C, G, eps, stats_mode, stats_mod);
  return check_launch("gn_coef_kernel");
}

extern "C" int omgsr_groupnorm_combine(const float* sums, const float* pixels, int entries, int n_img, int C, int G,
                                       float* out, float* wsum_out, void* stream) {
  if (entries <= 0 || n_img <= 0 || entries % n_img || C % G)
    return set_error(OMGSR_ERR_SHAPE, "groupnorm_combine: entries=%d n_img=%d", entries, n_img);
  const int n = n_img * G;
  launch_pdl(gn_combine_kernel, dim3((n + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream), sums, pixels, entries, n_img, G,
                                                                                 C / G, out, wsum_out);
  return check_launch("gn_combine_kernel");
}

static int layernorm_impl(const void* x, int x_f32, void* y, const float* gamma, const float* beta, long long rows, int C,
                          float eps, long long rows_per_batch, long long gb_stride, long long ldx,
                          long long x_batch_stride, int rms, void* stream);
extern "C" int omgsr_layernorm(const void* x, int x_f32, void* y, const float* gamma, const float* beta, long long rows,
                               int C, float eps, long long rows_per_batch, long long gb_stride, long long ldx,
                               long long x_batch_stride, void* stream) {
  return layernorm_impl(x, x_f32, y, gamma, beta, rows, C, eps, rows_per_batch, gb_stride, ldx, x_batch_stride, 0, stream);
}
/* T5LayerNorm: y = x * rsqrt(mean(x^2) + eps) * weight */
extern "C" int omgsr_rmsnorm(const void* x, int x_f32, void* y, const float* weight, long long rows, int C, float eps,
                             long long ldx, void* stream) {
  return layernorm_impl(x, x_f32, y, weight, nullptr, rows, C, eps, 0, 0, ldx, 0, 1, stream);
}
static int layernorm_impl(const void* x, int x_f32, void* y, const float* gamma, const float* beta, long long rows, int C,
                          float eps, long long rows_per_batch, long long gb_stride, long long ldx,
                          long long x_batch_stride, int rms, void* stream) {
  if (C % 8 || C > 4096) return set_error(OMGSR_ERR_SHAPE, "layernorm: C=%d", C);
  if (ldx <= 0) ldx = C;
  if ((ldx | x_batch_stride) % 8) return set_error(OMGSR_ERR_ALIGN, "layernorm: strides %% 8");
  if ((gb_stride % 4) || (reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15)
    return set_error(OMGSR_ERR_ALIGN, "layernorm: gamma/beta must be 16-byte aligned with gb_stride %% 4 == 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = 256;
  const int grid = grid_for(rows * 32, threads);
  const int vpl = (C / 8 + 31) / 32;
  bf16* yo = static_cast<bf16*>(y);
#define LN_CASE(V)                                                                                                  \
  do {                                                                                                              \
    if (x_f32)                                                                                                      \
      launch_pdl(layernorm_kernel<V, float>, dim3(grid), dim3(threads), 0, st, static_cast<const float*>(x), yo, gamma, beta, rows, C,  \
                                                           eps, rows_per_batch, gb_stride, ldx, x_batch_stride, rms);    \
    else                                                                                                            \
      launch_pdl(layernorm_kernel<V, bf16>, dim3(grid), dim3(threads), 0, st, static_cast<const bf16*>(x), yo, gamma, beta, rows, C,    \
                                                          eps, rows_per_batch, gb_stride, ldx, x_batch_stride, rms);     \
  } while (0)
  if (vpl <= 1) LN_CASE(1);
  else if (vpl <= 2) LN_CASE(2);
  else if (vpl <= 4) LN_CASE(4);
  else if (vpl <= 5) LN_CASE(5);
  else if (vpl <= 8) LN_CASE(8);
  else if (vpl <= 12) LN_CASE(12);
  else LN_CASE(16);
#undef LN_CASE
  return check_launch("layernorm_kernel");
}

extern "C" int omgsr_softmax_rows(const float* S, long long ld_s, void* P, long long ld_p, long long rows, int n,
                                  int n_pad, float scale, void* stream) {
  if (rows <= 0 || n <= 0 || n_pad < n) return set_error(OMGSR_ERR_SHAPE, "softmax_rows: rows=%lld n=%d", rows, n);
  if (rows > 0x7fffffffLL) return set_error(OMGSR_ERR_SHAPE, "softmax_rows: too many rows");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec_ok = (n_pad % 4 == 0) && (ld_s % 4 == 0) && (ld_p % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(S) & 15) == 0) && ((reinterpret_cast<uintptr_t>(P) & 7) == 0);
  if (vec_ok && n_pad >= 1024 && n_pad <= 32768) {
    const unsigned g = static_cast<unsigned>(rows);
    bf16* Pb = static_cast<bf16*>(P);
    const float sl = scale * 1.4426950408889634f;
    if (n_pad <= 2048) launch_pdl(softmax_rows_reg_kernel<2>, dim3(g), dim3(256), 0, st, S, ld_s, Pb, ld_p, n, n_pad, sl);
    else if (n_pad <= 4096) launch_pdl(softmax_rows_reg_kernel<4>, dim3(g), dim3(256), 0, st, S, ld_s, Pb, ld_p, n, n_pad, sl);
    else if (n_pad <= 8192) launch_pdl(softmax_rows_reg_kernel<8>, dim3(g), dim3(256), 0, st, S, ld_s, Pb, ld_p, n, n_pad, sl);
    // 1024^2 images: 16384 tokens in the VAE mid block (the three-pass kernel took 2 ms per launch there, 7x its
    // HBM time; profiles/r02/launches_f1024_b4_r02_summary.txt)
    else if (n_pad <= 16384) launch_pdl(softmax_rows_reg_kernel<16>, dim3(g), dim3(256), 0, st, S, ld_s, Pb, ld_p, n, n_pad, sl);
    else launch_pdl(softmax_rows_reg_kernel<32>, dim3(g), dim3(256), 0, st, S, ld_s, Pb, ld_p, n, n_pad, sl);
    return check_launch("softmax_rows_reg_kernel");
  }
  const int threads = n >= 2048 ? 256 : (n >= 512 ? 128 : 64);
  launch_pdl(softmax_rows_kernel, dim3(static_cast<unsigned>(rows)), dim3(threads), 0, static_cast<cudaStream_t>(stream), 
      S, ld_s, static_cast<bf16*>(P), ld_p, n, n_pad, scale * 1.4426950408889634f, 0, static_cast<const float*>(nullptr), 1LL, 0LL);
  return check_launch("softmax_rows_kernel");
}

extern "C" int omgsr_softmax_rows_causal(const float* S, long long ld_s, void* P, long long ld_p, long long rows, int n,
                                         int n_pad, float scale, int queries, void* stream) {
  if (rows <= 0 || n <= 0 || n_pad < n || queries <= 0 || rows % queries)
    return set_error(OMGSR_ERR_SHAPE, "softmax_rows_causal: rows=%lld n=%d queries=%d", rows, n, queries);
  if (rows > 0x7fffffffLL) return set_error(OMGSR_ERR_SHAPE, "softmax_rows_causal: too many rows");
  const int threads = n >= 2048 ? 256 : (n >= 512 ? 128 : 64);
  launch_pdl(softmax_rows_kernel, dim3(static_cast<unsigned>(rows)), dim3(threads), 0, static_cast<cudaStream_t>(stream),
      S, ld_s, static_cast<bf16*>(P), ld_p, n, n_pad, scale * 1.4426950408889634f, queries, static_cast<const float*>(nullptr), 1LL, 0LL);
  return check_launch("softmax_rows_kernel(causal)");
}

extern "C" int omgsr_softmax_rows_bias(const float* S, long long ld_s, void* P, long long ld_p, long long rows, int n,
                                       int n_pad, float scale, const float* bias, long long bias_rows, long long ld_b,
                                       void* stream) {
  if (rows <= 0 || n <= 0 || n_pad < n || !bias || bias_rows <= 0 || ld_b < n)
    return set_error(OMGSR_ERR_SHAPE, "softmax_rows_bias: rows=%lld n=%d bias_rows=%lld", rows, n, bias_rows);
  if (rows > 0x7fffffffLL) return set_error(OMGSR_ERR_SHAPE, "softmax_rows_bias: too many rows");
  const int threads = n >= 2048 ? 256 : (n >= 512 ? 128 : 64);
  launch_pdl(softmax_rows_kernel, dim3(static_cast<unsigned>(rows)), dim3(threads), 0, static_cast<cudaStream_t>(stream),
      S, ld_s, static_cast<bf16*>(P), ld_p, n, n_pad, scale * 1.4426950408889634f, 0, bias, bias_rows, ld_b);
  return check_launch("softmax_rows_kernel(bias)");
}

// elementwise product of two dense bf16 tensors (T5 gated feed-forward: gelu_new(wi_0 x) * wi_1 x)
__global__ void mul_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, long long n8) {
  pdl_launch_dependents();
  pdl_wait();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;  // grid_for caps the grid
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8; i += stride) {
  const uint4 x = reinterpret_cast<const uint4*>(a)[i], y = reinterpret_cast<const uint4*>(b)[i];
  const float2 x0 = unpack_bf16x2(x.x), x1 = unpack_bf16x2(x.y), x2 = unpack_bf16x2(x.z), x3 = unpack_bf16x2(x.w);
  const float2 y0 = unpack_bf16x2(y.x), y1 = unpack_bf16x2(y.y), y2 = unpack_bf16x2(y.z), y3 = unpack_bf16x2(y.w);
  uint4 o;
  o.x = pack_bf16x2(x0.x * y0.x, x0.y * y0.y); o.y = pack_bf16x2(x1.x * y1.x, x1.y * y1.y);
  o.z = pack_bf16x2(x2.x * y2.x, x2.y * y2.y); o.w = pack_bf16x2(x3.x * y3.x, x3.y * y3.y);
  reinterpret_cast<uint4*>(out)[i] = o;
  }
}

extern "C" int omgsr_mul_bf16(const void* a, const void* b, void* out, long long n, void* stream) {
  if (n <= 0 || (n % 8)) return set_error(OMGSR_ERR_SHAPE, "mul: n=%lld (must be a positive multiple of 8)", n);
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15)
    return set_error(OMGSR_ERR_ALIGN, "mul: 16-byte alignment");
  launch_pdl(mul_kernel, dim3(grid_for(n / 8, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
             static_cast<const bf16*>(a), static_cast<const bf16*>(b), static_cast<bf16*>(out), n / 8);
  return check_launch("mul_kernel");
}

// ------------------------------------------------------------------------------------------------
// Token + position embedding of the text encoders: out[r, :] = tok[ids[r], :] + pos[r % T, :] (fp32 add, bf16 out)
// ------------------------------------------------------------------------------------------------
__global__ void embed_tokens_kernel(const int* __restrict__ ids, const bf16* __restrict__ tok, const bf16* __restrict__ pos,
                                    bf16* __restrict__ out, long long rows, int T, int C, int vocab) {
  pdl_launch_dependents();
  pdl_wait();
  const int cg = C >> 3;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;  // grid_for caps the grid: grid-stride loop
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < rows * cg; idx += stride) {
  const long long r = idx / cg;
  const int c = static_cast<int>(idx - r * cg) * 8;
  int id = ids[r];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const uint4 a = *reinterpret_cast<const uint4*>(tok + static_cast<long long>(id) * C + c);
  uint4 o = a;
  if (pos) {
    const uint4 b = *reinterpret_cast<const uint4*>(pos + static_cast<long long>(r % T) * C + c);
    const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
    const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y), b2 = unpack_bf16x2(b.z), b3 = unpack_bf16x2(b.w);
    o.x = pack_bf16x2(a0.x + b0.x, a0.y + b0.y); o.y = pack_bf16x2(a1.x + b1.x, a1.y + b1.y);
    o.z = pack_bf16x2(a2.x + b2.x, a2.y + b2.y); o.w = pack_bf16x2(a3.x + b3.x, a3.y + b3.y);
  }
  *reinterpret_cast<uint4*>(out + r * C + c) = o;
  }
}

extern "C" int omgsr_embed_tokens(const int* ids, const void* tok, const void* pos, void* out, long long rows, int T,
                                  int C, int vocab, void* stream) {
  if (rows <= 0 || T <= 0 || C <= 0 || (C % 8) || vocab <= 0) return set_error(OMGSR_ERR_SHAPE, "embed_tokens: rows=%lld T=%d C=%d", rows, T, C);
  const long long total = rows * (C / 8);
  launch_pdl(embed_tokens_kernel, dim3(grid_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), ids,
             static_cast<const bf16*>(tok), static_cast<const bf16*>(pos), static_cast<bf16*>(out), rows, T, C, vocab);
  return check_launch("embed_tokens_kernel");
}

extern "C" int omgsr_rmsnorm_rope(void* x, long long ld, long long batch_stride, const float* weight,
                                  const float* cos_t, const float* sin_t, long long rows_per_batch, int batch,
                                  int heads, int heads_per_weight, int pos_offset, float eps, void* stream) {
  if ((ld | batch_stride) % 4) return set_error(OMGSR_ERR_ALIGN, "rmsnorm_rope: strides %% 4");
  if (heads <= 0 || heads_per_weight <= 0) return set_error(OMGSR_ERR_SHAPE, "rmsnorm_rope: heads");
  launch_pdl(rmsnorm_rope_kernel, dim3(grid_for(rows_per_batch * batch * heads * 32, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<bf16*>(x), ld, batch_stride, weight, cos_t, sin_t, rows_per_batch, batch, heads, heads_per_weight,
      pos_offset, eps);
  return check_launch("rmsnorm_rope_kernel");
}

extern "C" int omgsr_conv_in_nchw(const void* x, int x_f32, const float* w, const float* bias, void* y, int B, int Cin,
                                  int H, int W, int Cout, int ksize, long long ldy, void* stream) {
  if (Cout % 8 || (ksize != 1 && ksize != 3) || Cin > 64) return set_error(OMGSR_ERR_SHAPE, "conv_in: Cin=%d Cout=%d k=%d", Cin, Cout, ksize);
  if (ldy % 8) return set_error(OMGSR_ERR_ALIGN, "conv_in: ldy %% 8");
  const int cg = Cout / 8;
  const long long total = static_cast<long long>(B) * H * W * cg;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (x_f32)
    launch_pdl(conv_in_kernel<float>, dim3(grid_for(total, 256) * 4), dim3(256), 0, st, static_cast<const float*>(x), w, bias,
                                                                   static_cast<bf16*>(y), B, Cin, H, W, Cout, ksize,
                                                                   ldy, cg, total);
  else
    launch_pdl(conv_in_kernel<bf16>, dim3(grid_for(total, 256) * 4), dim3(256), 0, st, static_cast<const bf16*>(x), w, bias,
                                                                  static_cast<bf16*>(y), B, Cin, H, W, Cout, ksize,
                                                                  ldy, cg, total);
  return check_launch("conv_in_kernel");
}

extern "C" int omgsr_im2col3x3_nchw(const void* x, int x_f32, void* y, int B, int Cin, int H, int W, int Kpad,
                                    void* stream) {
  if (Cin <= 0 || Cin > 16 || Kpad % 8 || Kpad < 9 * Cin) return set_error(OMGSR_ERR_SHAPE, "im2col3x3: Cin=%d Kpad=%d", Cin, Kpad);
  const long long total = static_cast<long long>(B) * H * W * (Kpad / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (x_f32)
    launch_pdl(im2col3x3_kernel<float>, dim3(grid_for(total, 256) * 2), dim3(256), 0, st, static_cast<const float*>(x), static_cast<bf16*>(y),
                                                                      Cin, H, W, Kpad, total);
  else
    launch_pdl(im2col3x3_kernel<bf16>, dim3(grid_for(total, 256) * 2), dim3(256), 0, st, static_cast<const bf16*>(x), static_cast<bf16*>(y),
                                                                     Cin, H, W, Kpad, total);
  return check_launch("im2col3x3_kernel");
}

template <int COUT, typename TOut>
static int launch_conv_out(const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int Cin,
                           int clamp, long long ldx, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(COUT) * 9 * Cin * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(conv_out_kernel<COUT, TOut>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return set_error(OMGSR_ERR_CUDA, "conv_out smem: %s", cudaGetErrorString(e));
  }
  const long long total = static_cast<long long>(B) * H * W;
  launch_pdl(conv_out_kernel<COUT, TOut>, dim3(grid_for(total, 128)), dim3(128), smem, st, 
      static_cast<const bf16*>(x), static_cast<const bf16*>(w), bias, static_cast<TOut*>(y), B, H, W, Cin, clamp, ldx,
      total);
  return check_launch("conv_out_kernel");
}

extern "C" int omgsr_conv_out_nchw(const void* x, const void* w, const float* bias, void* y, int y_f32, int B, int H,
                                   int W, int Cin, int Cout, int clamp, long long ldx, void* stream) {
  if (Cin % 8 || ldx % 8) return set_error(OMGSR_ERR_SHAPE, "conv_out: Cin=%d ldx=%lld", Cin, ldx);
  if (static_cast<size_t>(Cout) * 9 * Cin * 4 > 200 * 1024) return set_error(OMGSR_ERR_SHAPE, "conv_out: weights exceed smem");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (Cout == 3)
    return y_f32 ? launch_conv_out<3, float>(x, w, bias, y, B, H, W, Cin, clamp, ldx, st)
                 : launch_conv_out<3, bf16>(x, w, bias, y, B, H, W, Cin, clamp, ldx, st);
  if (Cout == 4)
    return y_f32 ? launch_conv_out<4, float>(x, w, bias, y, B, H, W, Cin, clamp, ldx, st)
                 : launch_conv_out<4, bf16>(x, w, bias, y, B, H, W, Cin, clamp, ldx, st);
  return set_error(OMGSR_ERR_SHAPE, "conv_out: Cout=%d unsupported (3 or 4)", Cout);
}

extern "C" int omgsr_posterior_sample(const float* moments, long long ldm, const float* eps, void* z, int z_f32, int B,
                                      int Cz, int h, int w, float shift, float scale, void* stream) {
  const long long hw = static_cast<long long>(h) * w, total = hw * Cz * B;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (z_f32)
    launch_pdl(posterior_kernel<float>, dim3(grid_for(total, 256)), dim3(256), 0, st, moments, ldm, eps, static_cast<float*>(z), Cz, hw,
                                                                  total, shift, scale);
  else
    launch_pdl(posterior_kernel<bf16>, dim3(grid_for(total, 256)), dim3(256), 0, st, moments, ldm, eps, static_cast<bf16*>(z), Cz, hw,
                                                                 total, shift, scale);
  return check_launch("posterior_kernel");
}

extern "C" int omgsr_latent_axpby(const void* x, const void* y, void* out, int is_f32, long long n, float a, float b,
                                  float c, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (is_f32)
    launch_pdl(axpby_kernel<float>, dim3(grid_for(n, 256)), dim3(256), 0, st, static_cast<const float*>(x), static_cast<const float*>(y),
                                                          static_cast<float*>(out), n, a, b, c);
  else
    launch_pdl(axpby_kernel<bf16>, dim3(grid_for(n, 256)), dim3(256), 0, st, static_cast<const bf16*>(x), static_cast<const bf16*>(y),
                                                         static_cast<bf16*>(out), n, a, b, c);
  return check_launch("axpby_kernel");
}

extern "C" int omgsr_x0_unscale(const void* z, int z_f32, const float* eps_hat, long long lde, void* out,
                                const float* Wpq, const float* bpq, int B, int Cz, int h, int w, float s1, float s2,
                                float scale, float shift, void* stream) {
  if (Cz > 16) return set_error(OMGSR_ERR_SHAPE, "x0_unscale: Cz=%d > 16", Cz);
  const long long hw = static_cast<long long>(h) * w, total = hw * B;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (z_f32)
    launch_pdl(x0_unscale_kernel<float>, dim3(grid_for(total, 128)), dim3(128), 0, st, static_cast<const float*>(z), eps_hat, lde,
                                                                   static_cast<bf16*>(out), Wpq, bpq, Cz, hw, total, s1,
                                                                   1.0f / s2, 1.0f / scale, shift);
  else
    launch_pdl(x0_unscale_kernel<bf16>, dim3(grid_for(total, 128)), dim3(128), 0, st, static_cast<const bf16*>(z), eps_hat, lde,
                                                                  static_cast<bf16*>(out), Wpq, bpq, Cz, hw, total, s1,
                                                                  1.0f / s2, 1.0f / scale, shift);
  return check_launch("x0_unscale_kernel");
}

extern "C" int omgsr_pack_latents(const void* z, void* tokens, int B, int C, int h, int w, long long ld_tok,
                                  void* stream) {
  if ((h | w) & 1) return set_error(OMGSR_ERR_SHAPE, "pack_latents: odd latent size");
  const long long total = static_cast<long long>(B) * C * h * w;
  launch_pdl(pack_kernel, dim3(grid_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const bf16*>(z), static_cast<bf16*>(tokens), C, h, w, ld_tok, total, 0);
  return check_launch("pack_kernel");
}
extern "C" int omgsr_unpack_latents(const void* tokens, void* z, int B, int C, int h, int w, long long ld_tok,
                                    void* stream) {
  if ((h | w) & 1) return set_error(OMGSR_ERR_SHAPE, "unpack_latents: odd latent size");
  const long long total = static_cast<long long>(B) * C * h * w;
  launch_pdl(pack_kernel, dim3(grid_for(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const bf16*>(z), const_cast<bf16*>(static_cast<const bf16*>(tokens)), C, h, w, ld_tok, total, 1);
  return check_launch("pack_kernel");
}

extern "C" int omgsr_tile_blend(const void* tiles, int tiles_f32, const float* weights, const int* ofs_xy_dev,
                                int ntiles, float* out, int B, int C, int h, int w, int T, const long long* tile_strides,
                                const long long* out_strides, void* stream) {
  const long long total = static_cast<long long>(B) * C * h * w;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BlendStrides bs;
  bs.t_t = tile_strides[0]; bs.t_b = tile_strides[1]; bs.t_c = tile_strides[2]; bs.t_y = tile_strides[3];
  bs.t_x = tile_strides[4];
  bs.o_b = out_strides[0]; bs.o_c = out_strides[1]; bs.o_y = out_strides[2]; bs.o_x = out_strides[3];
  if (tiles_f32)
    launch_pdl(tile_blend_kernel<float>, dim3(grid_for(total, 256)), dim3(256), 0, st, static_cast<const float*>(tiles), weights,
                                                                   ofs_xy_dev, ntiles, out, B, C, h, w, T, total, bs);
  else
    launch_pdl(tile_blend_kernel<bf16>, dim3(grid_for(total, 256)), dim3(256), 0, st, static_cast<const bf16*>(tiles), weights, ofs_xy_dev,
                                                                  ntiles, out, B, C, h, w, T, total, bs);
  return check_launch("tile_blend_kernel");
}

extern "C" int omgsr_nchw_to_nhwc(const void* x, int x_f32, void* y, int B, int C, int H, int W, long long ldy,
                                  void* stream) {
  const long long HW = static_cast<long long>(H) * W;
  dim3 grid(static_cast<unsigned>((HW + 31) / 32), (C + 31) / 32, B), block(32, 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (x_f32)
    launch_pdl(nchw_to_nhwc_kernel<float>, dim3(grid), dim3(block), 0, st, static_cast<const float*>(x), static_cast<bf16*>(y), C, HW, ldy);
  else
    launch_pdl(nchw_to_nhwc_kernel<bf16>, dim3(grid), dim3(block), 0, st, static_cast<const bf16*>(x), static_cast<bf16*>(y), C, HW, ldy);
  return check_launch("nchw_to_nhwc_kernel");
}
extern "C" int omgsr_nhwc_to_nchw(const void* x, long long ldx, void* y, int y_f32, int B, int C, int H, int W,
                                  int clamp, void* stream) {
  const long long HW = static_cast<long long>(H) * W;
  dim3 grid(static_cast<unsigned>((HW + 31) / 32), (C + 31) / 32, B), block(32, 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (y_f32)
    launch_pdl(nhwc_to_nchw_kernel<float>, dim3(grid), dim3(block), 0, st, static_cast<const bf16*>(x), ldx, static_cast<float*>(y), C, HW,
                                                       clamp);
  else
    launch_pdl(nhwc_to_nchw_kernel<bf16>, dim3(grid), dim3(block), 0, st, static_cast<const bf16*>(x), ldx, static_cast<bf16*>(y), C, HW,
                                                      clamp);
  return check_launch("nhwc_to_nchw_kernel");
}
