// Colour alignment of the SR output to the (upsampled) LQ input, on device.
// Follows infer/wavelet_color_fix.py: wavelet_reconstruction (:114-125) = high-frequency of the content
// plus 5-level a-trous low-pass of the style; adaptive_instance_normalization (:60-74).
#include "common.cuh"
#include "../../include/omgsr_b200.h"

namespace omgsr {

// one a-trous stage: out = blur(in) with the 3x3 binomial kernel at dilation r, replicate padding
// (wavelet_blur, infer/wavelet_color_fix.py:77-96). first: in = style - content. last: out += content.
__global__ void atrous_kernel(const float* __restrict__ in, const float* __restrict__ in2, const float* __restrict__ add,
                              float* __restrict__ out, int H, int W, int r, long long total) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const int y = static_cast<int>((i / W) % H);
    const long long plane = i - (static_cast<long long>(y) * W + x);
    const int xs[3] = {max(x - r, 0), x, min(x + r, W - 1)};
    const int ys[3] = {max(y - r, 0), y, min(y + r, H - 1)};
    const float kw[3] = {0.25f, 0.5f, 0.25f};
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float row = 0.f;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const long long idx = plane + static_cast<long long>(ys[a]) * W + xs[b];
        float v = in[idx];
        if (in2) v -= in2[idx];
        row += kw[b] * v;
      }
      acc += kw[a] * row;
    }
    if (add) acc += add[i];
    out[i] = acc;
  }
}

__global__ void plane_moments_kernel(const float* __restrict__ a, const float* __restrict__ b, double* __restrict__ acc,
                                     long long HW, int chunks) {
  // acc[plane*4 + {0,1,2,3}] = sum a, sum a^2, sum b, sum b^2
  __shared__ double red[4][32];
  const long long plane = blockIdx.x / chunks;
  const int chunk = blockIdx.x % chunks;
  const float* pa = a + plane * HW;
  const float* pb = b + plane * HW;
  double s[4] = {0, 0, 0, 0};
  for (long long i = static_cast<long long>(chunk) * blockDim.x + threadIdx.x; i < HW;
       i += static_cast<long long>(chunks) * blockDim.x) {
    const double va = pa[i], vb = pb[i];
    s[0] += va; s[1] += va * va; s[2] += vb; s[3] += vb * vb;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    if (lane == 0) red[k][warp] = s[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double v = lane < (blockDim.x >> 5) ? red[k][lane] : 0.0;
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) atomicAdd(&acc[plane * 4 + k], v);
    }
  }
}

__global__ void adain_apply_kernel(const float* __restrict__ content, float* __restrict__ out,
                                   const double* __restrict__ acc, long long HW, long long total) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long plane = i / HW;
    const double n = static_cast<double>(HW);
    const double mc = acc[plane * 4] / n, ms = acc[plane * 4 + 2] / n;
    // unbiased variance + 1e-5 (calc_mean_std, infer/wavelet_color_fix.py:44-57)
    const double vc = (acc[plane * 4 + 1] - n * mc * mc) / (n - 1.0) + 1e-5;
    const double vs = (acc[plane * 4 + 3] - n * ms * ms) / (n - 1.0) + 1e-5;
    const float sc = static_cast<float>(sqrt(vc)), ss = static_cast<float>(sqrt(vs));
    out[i] = (content[i] - static_cast<float>(mc)) / sc * ss + static_cast<float>(ms);
  }
}

static int grid_cf(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  return static_cast<int>(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace omgsr

using namespace omgsr;

extern "C" int omgsr_wavelet_fix(const float* content, const float* style, float* out, float* tmp0, float* tmp1, int B,
                                 int C, int H, int W, void* stream) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return set_error(OMGSR_ERR_SHAPE, "wavelet_fix: empty");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(B) * C * H * W;
  const int grid = grid_cf(total, 256);
  // LP5(style - content): radius 1, 2, 4, 8, 16; the last stage adds the content back
  atrous_kernel<<<grid, 256, 0, st>>>(style, content, nullptr, tmp0, H, W, 1, total);
  int rc = check_launch("atrous_kernel");
  if (rc) return rc;
  float* src = tmp0;
  float* dst = tmp1;
  for (int lvl = 1; lvl < 5; ++lvl) {
    const bool last = (lvl == 4);
    atrous_kernel<<<grid, 256, 0, st>>>(src, nullptr, last ? content : nullptr, last ? out : dst, H, W, 1 << lvl, total);
    rc = check_launch("atrous_kernel");
    if (rc) return rc;
    float* t = src; src = dst; dst = t;
  }
  return OMGSR_OK;
}

extern "C" int omgsr_atrous_blur(const float* in, float* out, int B, int C, int H, int W, int radius, void* stream) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || radius <= 0) return set_error(OMGSR_ERR_SHAPE, "atrous_blur: bad shape");
  const long long total = static_cast<long long>(B) * C * H * W;
  atrous_kernel<<<grid_cf(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, nullptr, nullptr, out, H, W,
                                                                                  radius, total);
  return check_launch("atrous_kernel");
}

extern "C" int omgsr_adain_fix(const float* content, const float* style, float* out, void* scratch, int B, int C, int H,
                               int W, void* stream) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return set_error(OMGSR_ERR_SHAPE, "adain_fix: empty");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long HW = static_cast<long long>(H) * W, planes = static_cast<long long>(B) * C;
  double* acc = static_cast<double*>(scratch);
  cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double) * 4 * planes, st);
  if (e != cudaSuccess) return set_error(OMGSR_ERR_CUDA, "adain memset: %s", cudaGetErrorString(e));
  int chunks = static_cast<int>((HW + 256 * 16 - 1) / (256 * 16));
  if (chunks < 1) chunks = 1;
  if (chunks > 1024) chunks = 1024;
  plane_moments_kernel<<<static_cast<unsigned>(planes * chunks), 256, 0, st>>>(content, style, acc, HW, chunks);
  int rc = check_launch("plane_moments_kernel");
  if (rc) return rc;
  adain_apply_kernel<<<grid_cf(planes * HW, 256), 256, 0, st>>>(content, out, acc, HW, planes * HW);
  return check_launch("adain_apply_kernel");
}
