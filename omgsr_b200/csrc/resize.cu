// Driver-side image pre-processing on the GPU: the reference's PIL chain
//   img.resize((w * upscale, h * upscale))            (Pillow default filter: BICUBIC)
//   img.resize((w - w % 8, h - h % 8), Image.LANCZOS)
//   to_tensor(img).to(weight_dtype) * 2 - 1
// (infer/infer_omgsr_s.py:76-90) restated bit-exactly for 8-bit RGB. Pillow's resampler is pure integer arithmetic
// once the coefficient tables exist (src/libImaging/Resample.c, Pillow 12.2: precompute_coeffs, normalize_coeffs_8bpc,
// ImagingResampleHorizontal_8bpc / Vertical_8bpc): per axis a table of (first tap, tap count) bounds and taps in
// 22-bit fixed point; a horizontal pass to an 8-bit intermediate, then a vertical pass, each
//   out = clip8((2^21 + sum_k in[first + k] * coef[k]) >> 22).
// The tables are built on the host in double precision exactly as Pillow does (same expression order), the two
// passes run as byte kernels. Integer work: the parity bar is bit equality with PIL (tests/test_resize_*.py).
#include "common.cuh"
#include "../../include/omgsr_b200.h"
#include <cmath>
#include <vector>

namespace omgsr {

constexpr int kPrecisionBits = 32 - 8 - 2;

static double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}
static double sinc_filter(double x) {
  if (x == 0.0) return 1.0;
  x = x * M_PI;
  return sin(x) / x;
}
static double lanczos_filter(double x) {
  if (-3.0 <= x && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3);
  return 0.0;
}

// returns ksize; bounds[2 * out] = (first tap, tap count), coeffs[out * ksize + k] in 22-bit fixed point
static int build_coeffs(int in_size, int out_size, int filter, std::vector<int>& bounds, std::vector<int>& coeffs) {
  double (*f)(double) = filter == OMGSR_RESIZE_LANCZOS ? lanczos_filter : bicubic_filter;
  const double fsupport = filter == OMGSR_RESIZE_LANCZOS ? 3.0 : 2.0;
  const double in0 = 0.0, in1 = in_size;
  double scale, filterscale;
  filterscale = scale = (in1 - in0) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = fsupport * filterscale;
  const int ksize = static_cast<int>(ceil(support)) * 2 + 1;
  bounds.assign(static_cast<size_t>(out_size) * 2, 0);
  coeffs.assign(static_cast<size_t>(out_size) * ksize, 0);
  std::vector<double> k(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = in0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int x;
    for (x = 0; x < xmax; ++x) {
      const double w = f((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (; x < ksize; ++x) k[x] = 0;
    bounds[xx * 2] = xmin;
    bounds[xx * 2 + 1] = xmax;
    for (x = 0; x < ksize; ++x) {
      const double v = k[x] * (1 << kPrecisionBits);
      coeffs[static_cast<size_t>(xx) * ksize + x] = static_cast<int>(k[x] < 0 ? -0.5 + v : 0.5 + v);
    }
  }
  return ksize;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;  // arithmetic shift, as Pillow's lookup on in >> PRECISION_BITS
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// one thread per output pixel (3 interleaved channels). HORIZONTAL: taps run along x of the source row;
// otherwise along y of the source column.
template <bool HORIZONTAL>
__global__ void resample_u8_kernel(const uint8_t* __restrict__ src, int src_w, uint8_t* __restrict__ dst, int dst_h,
                                   int dst_w, const int* __restrict__ bounds, const int* __restrict__ coeffs,
                                   int ksize) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(dst_h) * dst_w) return;
  const int x = static_cast<int>(i % dst_w), y = static_cast<int>(i / dst_w);
  const int o = HORIZONTAL ? x : y;
  const int first = bounds[2 * o], n = bounds[2 * o + 1];
  const int* k = coeffs + static_cast<long long>(o) * ksize;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int t = 0; t < n; ++t) {
    const long long p = HORIZONTAL ? (static_cast<long long>(y) * src_w + first + t) * 3
                                   : (static_cast<long long>(first + t) * src_w + x) * 3;
    const int c = k[t];
    s0 += src[p] * c; s1 += src[p + 1] * c; s2 += src[p + 2] * c;
  }
  uint8_t* d = dst + i * 3;
  d[0] = clip8(s0); d[1] = clip8(s1); d[2] = clip8(s2);
}

// to_tensor(img).to(dtype) * 2 - 1: u8 / 255 in fp32, rounded to bf16, then 2t - 1 evaluated in fp32 and rounded
// again (what the bf16 tensor ops of the reference line do); HWC u8 -> NCHW
template <typename T>
__global__ void u8_to_input_kernel(const uint8_t* __restrict__ src, T* __restrict__ dst, int H, int W) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long hw = static_cast<long long>(H) * W;
  if (i >= hw) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float t = static_cast<float>(src[i * 3 + c]) / 255.0f;
    if (sizeof(T) == 2) {
      const float tb = __bfloat162float(__float2bfloat16(t));
      reinterpret_cast<bf16*>(dst)[c * hw + i] = __float2bfloat16(tb * 2.0f - 1.0f);
    } else {
      reinterpret_cast<float*>(dst)[c * hw + i] = t * 2.0f - 1.0f;
    }
  }
}

// clip(x * 0.5 + 0.5, 0, 1).float() -> ToPILImage (mul(255).byte(): truncation); NCHW -> HWC u8
// (infer/infer_omgsr_s.py:96-98). x * 0.5 + 0.5 is evaluated in the tensor's own dtype first, as the reference does.
template <typename T>
__global__ void output_to_u8_kernel(const T* __restrict__ src, uint8_t* __restrict__ dst, int H, int W) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long hw = static_cast<long long>(H) * W;
  if (i >= hw) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v;
    if (sizeof(T) == 2) {
      const float x = __bfloat162float(reinterpret_cast<const bf16*>(src)[c * hw + i]);
      const float h = __bfloat162float(__float2bfloat16(x * 0.5f));       // bf16 multiply
      v = __bfloat162float(__float2bfloat16(h + 0.5f));                   // bf16 add
    } else {
      v = reinterpret_cast<const float*>(src)[c * hw + i] * 0.5f + 0.5f;
    }
    v = fminf(fmaxf(v, 0.0f), 1.0f);
    dst[i * 3 + c] = static_cast<uint8_t>(v * 255.0f);
  }
}

}  // namespace omgsr

using namespace omgsr;

extern "C" int omgsr_resample_coeffs(int in_size, int out_size, int filter, int* bounds, int* coeffs, int max_ksize,
                                     int* ksize_out) {
  if (in_size <= 0 || out_size <= 0) return set_error(OMGSR_ERR_SHAPE, "resample_coeffs: empty axis");
  if (filter != OMGSR_RESIZE_BICUBIC && filter != OMGSR_RESIZE_LANCZOS)
    return set_error(OMGSR_ERR_SHAPE, "resample_coeffs: filter must be OMGSR_RESIZE_BICUBIC or _LANCZOS");
  std::vector<int> b, c;
  const int ksize = build_coeffs(in_size, out_size, filter, b, c);
  if (ksize_out) *ksize_out = ksize;
  if (!bounds || !coeffs) return OMGSR_OK;  // size query
  if (ksize > max_ksize) return set_error(OMGSR_ERR_SHAPE, "resample_coeffs: ksize %d > max_ksize %d", ksize, max_ksize);
  for (size_t i = 0; i < b.size(); ++i) bounds[i] = b[i];
  for (size_t i = 0; i < c.size(); ++i) coeffs[i] = c[i];
  return OMGSR_OK;
}

extern "C" long long omgsr_resize_u8_workspace(int src_h, int src_w, int dst_h, int dst_w, int filter) {
  std::vector<int> b, c;
  const int kx = build_coeffs(src_w, dst_w, filter, b, c), ky = build_coeffs(src_h, dst_h, filter, b, c);
  long long bytes = static_cast<long long>(src_h) * dst_w * 3;  // horizontal-pass intermediate
  bytes = (bytes + 255) & ~255LL;
  bytes += (2LL * dst_w + static_cast<long long>(dst_w) * kx + 2LL * dst_h + static_cast<long long>(dst_h) * ky) * 4 + 1024;
  return bytes;
}

extern "C" int omgsr_resize_u8(const void* src, int src_h, int src_w, void* dst, int dst_h, int dst_w, int filter,
                               void* workspace, long long workspace_bytes, void* stream) {
  if (src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0) return set_error(OMGSR_ERR_SHAPE, "resize_u8: empty image");
  if (filter != OMGSR_RESIZE_BICUBIC && filter != OMGSR_RESIZE_LANCZOS)
    return set_error(OMGSR_ERR_SHAPE, "resize_u8: filter must be OMGSR_RESIZE_BICUBIC or _LANCZOS");
  if (workspace_bytes < omgsr_resize_u8_workspace(src_h, src_w, dst_h, dst_w, filter))
    return set_error(OMGSR_ERR_SHAPE, "resize_u8: workspace too small (see omgsr_resize_u8_workspace)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::vector<int> bx, cx, by, cy;
  const int kx = build_coeffs(src_w, dst_w, filter, bx, cx), ky = build_coeffs(src_h, dst_h, filter, by, cy);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  uint8_t* tmp = ws;
  long long off = (static_cast<long long>(src_h) * dst_w * 3 + 255) & ~255LL;
  int* d_bx = reinterpret_cast<int*>(ws + off); off += bx.size() * 4;
  int* d_cx = reinterpret_cast<int*>(ws + off); off += cx.size() * 4;
  int* d_by = reinterpret_cast<int*>(ws + off); off += by.size() * 4;
  int* d_cy = reinterpret_cast<int*>(ws + off);
  // pageable sources: cudaMemcpyAsync stages them before returning, so the vectors may die with this frame
  cudaError_t e = cudaMemcpyAsync(d_bx, bx.data(), bx.size() * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_cx, cx.data(), cx.size() * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_by, by.data(), by.size() * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_cy, cy.data(), cy.size() * 4, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return set_error(OMGSR_ERR_CUDA, "resize_u8: table upload: %s", cudaGetErrorString(e));
  // Pillow runs the horizontal pass first (only when the width changes), then the vertical one
  const uint8_t* cur = static_cast<const uint8_t*>(src);
  int cur_w = src_w;
  const bool need_h = dst_w != src_w, need_v = dst_h != src_h;
  if (need_h) {
    uint8_t* out = need_v ? tmp : static_cast<uint8_t*>(dst);
    const long long n = static_cast<long long>(src_h) * dst_w;
    resample_u8_kernel<true><<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(cur, cur_w, out, src_h, dst_w,
                                                                                       d_bx, d_cx, kx);
    int rc = check_launch("resample_u8_kernel<h>");
    if (rc) return rc;
    cur = out;
    cur_w = dst_w;
  }
  if (need_v) {
    const long long n = static_cast<long long>(dst_h) * dst_w;
    resample_u8_kernel<false><<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
        cur, cur_w, static_cast<uint8_t*>(dst), dst_h, dst_w, d_by, d_cy, ky);
    return check_launch("resample_u8_kernel<v>");
  }
  if (!need_h) {
    e = cudaMemcpyAsync(dst, src, static_cast<size_t>(src_h) * src_w * 3, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return set_error(OMGSR_ERR_CUDA, "resize_u8: copy: %s", cudaGetErrorString(e));
  }
  return OMGSR_OK;
}

extern "C" int omgsr_u8_to_input(const void* hwc_u8, void* nchw, int out_f32, int H, int W, void* stream) {
  if (H <= 0 || W <= 0) return set_error(OMGSR_ERR_SHAPE, "u8_to_input: empty image");
  const long long n = static_cast<long long>(H) * W;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>((n + 255) / 256);
  if (out_f32)
    u8_to_input_kernel<float><<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(hwc_u8), static_cast<float*>(nchw), H, W);
  else
    u8_to_input_kernel<bf16><<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(hwc_u8), static_cast<bf16*>(nchw), H, W);
  return check_launch("u8_to_input_kernel");
}

extern "C" int omgsr_output_to_u8(const void* nchw, int in_f32, void* hwc_u8, int H, int W, void* stream) {
  if (H <= 0 || W <= 0) return set_error(OMGSR_ERR_SHAPE, "output_to_u8: empty image");
  const long long n = static_cast<long long>(H) * W;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>((n + 255) / 256);
  if (in_f32)
    output_to_u8_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(nchw), static_cast<uint8_t*>(hwc_u8), H, W);
  else
    output_to_u8_kernel<bf16><<<grid, 256, 0, st>>>(static_cast<const bf16*>(nchw), static_cast<uint8_t*>(hwc_u8), H, W);
  return check_launch("output_to_u8_kernel");
}
