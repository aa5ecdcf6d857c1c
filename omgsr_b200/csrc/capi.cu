#include <cstdlib>
// C-ABI plumbing: error state, tensor-map encoding through the runtime-resolved driver entry point,
// device properties, launch counter.
#include "common.cuh"
#include "../../include/omgsr_b200.h"
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cudaTypedefs.h>

namespace omgsr {
int g_pdl = getenv("OMGSR_PDL") ? atoi(getenv("OMGSR_PDL")) : 1;

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(OMGSR_ERR_CUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return OMGSR_OK;
}

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return set_error(OMGSR_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (box[i] == 0 || box[i] > 256) return set_error(OMGSR_ERR_SHAPE, "tensor map: box[%d]=%u out of range", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gs[i] = strides_bytes[i];
    if (gs[i] % 16) return set_error(OMGSR_ERR_ALIGN, "tensor map: stride[%d]=%llu not 16-byte aligned", i,
                                     static_cast<unsigned long long>(gs[i]));
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gd,
                   gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(OMGSR_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu,%llu] box [%u,%u,%u,%u,%u]",
                     static_cast<int>(r), rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
                     (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0),
                     (unsigned long long)(rank > 4 ? gd[4] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0,
                     rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0);
  }
  return OMGSR_OK;
}

int sm_count() {
  static int n[64] = {};  // per device: a process may drive any device of the box (infer_omgsr_s.py --device cuda:N)
  int dev = 0;
  cudaGetDevice(&dev);
  dev = (dev < 0 || dev >= 64) ? 0 : dev;
  if (!n[dev]) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}

int ensure_smem_attr(const void* func, int bytes, unsigned long long* done_mask, const char* what) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev = (dev < 0 || dev >= 64) ? 0 : dev;
  if ((*done_mask >> dev) & 1ull) return OMGSR_OK;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return set_error(OMGSR_ERR_CUDA, "cudaFuncSetAttribute(%s): %s", what, cudaGetErrorString(e));
  *done_mask |= 1ull << dev;
  return OMGSR_OK;
}

}  // namespace omgsr

extern "C" int omgsr_version(void) { return 100; }
extern "C" const char* omgsr_last_error(void) { return omgsr::g_err; }
extern "C" int omgsr_destroy(void) { return OMGSR_OK; }
extern "C" long long omgsr_launch_count(void) { return omgsr::g_launches.load(); }
namespace omgsr { extern const char* g_trace_variant; }
extern "C" const char* omgsr_last_variant(void) { return omgsr::g_trace_variant; }
