// mma_probe.cu — stand-alone tcgen05.mma issue-rate probe (round-2 design input for the 128-channel halo convolution).
//
// One elected thread per CTA (pair) issues a fixed pattern of tcgen05.mma instructions on whatever is in shared
// memory, commits, waits, and reports clock64 cycles per pattern. All 148 SMs run the same loop, so clocks / power are
// those of a real kernel. Optional interference: extra warps that read-modify-write shared memory (the GroupNorm
// transform of conv_halo_tc2_kernel<2,true>) and a TMA stream into a ring (operand refills).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I omgsr_b200/csrc scripts/mma_probe.cu \
//        omgsr_b200/csrc/capi.cu -o scripts/mma_probe
//   ./scripts/mma_probe            # prints one line per pattern
//
// Patterns (M is 128 rows per CTA in every case):
//   single N=256 / N=128 / N=64     cta_group::1: A 4 KB + B N*32 B of shared-memory operand reads per instruction
//   pair   N=256 / N=128 / N=64     cta_group::2: A 4 KB + B N*16 B per CTA per instruction
//   pair   mix 128,256,256,128      the row-shared schedule: for a fixed tap column the (output row, tap row) pairs that
//                                   read the same input row share one instruction (N = 256 = two accumulators)
//   a_views = k                     the A descriptor cycles through k start addresses 128 B apart (shifted halo views)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"

using namespace omgsr;

namespace omgsr { const char* g_trace_variant = ""; }  // capi.cu (linked for make_tmap_bf16) refers to it

struct ProbeParams {
  int pair;        // 0: cta_group::1, 1: cta_group::2
  int pattern[8];  // N of each instruction of the pattern (0 terminates)
  int a_views;     // number of distinct A start rows cycled through
  int iters;       // pattern repetitions (each instruction is issued 4x: the K = 64 block)
  int xf_warps;    // extra warps doing LDS / FMA / tanh / STS on a 64 KB region (0 = none)
  int tma;         // 1: a producer thread streams 16 KB boxes into a 4-deep ring (not consumed)
  int a_mn, b_mn;  // operand majors: 1 = MN-major (transposed) A / B, as the attention kernels use for V
  int a_tmem;      // single-CTA only: the A operand comes from TMEM (tcgen05.mma [d], [a_tmem], b_desc)
  int commit_each; // 1: tcgen05.commit (multicast in pair mode) to a dummy barrier after EVERY pattern element (4 MMAs)
  int d_mode;      // 0: the 4 K-steps of a pattern element hit one accumulator; 1: consecutive instructions alternate
                   //    between two accumulators (independent chains); 2: every instruction of the run hits ONE accumulator
  long long* cycles;  // [gridDim.x]
};

constexpr int kSmem = 200 * 1024;

template <bool PAIR>
__global__ void __launch_bounds__(256, 1) probe_kernel(const __grid_constant__ CUtensorMap tm, const ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 80 KB: A views
  uint8_t* sB = smem + 80 * 1024;     // 48 KB: B blocks
  uint8_t* sX = smem + 128 * 1024;    // 64 KB: transform / TMA scratch
  __shared__ uint64_t done_bar, ring_bar[4], dummy_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if constexpr (PAIR) rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i * 2654435761u >> 20 & 0x00ff00ffu);  // small bf16 values
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    mbar_init(&dummy_bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&ring_bar[i], 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (warp == 1) {
    if constexpr (PAIR) { tmem_alloc2(&tmem_slot, 512); tmem_relinquish2(); }
    else { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();

  if (warp == 0 && lane == 0) {
    if (!PAIR || rank == 0) {
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      long long t0 = clock64();
      int view = 0;
      for (int it = 0; it < p.iters; ++it) {
        for (int j = 0; j < 8 && p.pattern[j]; ++j) {
          const int N = p.pattern[j];
          const uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, N, p.a_mn, p.b_mn);
          const uint32_t a_addr = a0 + view * 128;
          if (++view >= p.a_views) view = 0;
          const uint32_t d0 = tmem_base + ((p.d_mode != 2 && (j & 1) && !p.a_tmem) ? 256 : 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = p.a_mn ? umma_desc_sw128(a_addr + k * 2048, 16384, 1024) : umma_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = p.b_mn ? umma_desc_sw128(b0 + k * 2048, 16384, 1024)
                                       : umma_desc_sw128(b0 + (j & 3) * 4096 + k * 32, 16, 1024);
            const uint32_t d = (p.d_mode == 1) ? tmem_base + ((k & 1) ? 256 : 0) : d0;
            if constexpr (PAIR) tc_mma2_ss(d, da, db, idesc, 1);
            else if (p.a_tmem) tc_mma_ts(d, tmem_base + 448 + k * 8, db, idesc, 1);  // A operand from TMEM (8 columns / K-step)
            else tc_mma_ss(d, da, db, idesc, 1);
          }
          if (p.commit_each) { if constexpr (PAIR) tc_commit2_mc(&dummy_bar); else tc_commit(&dummy_bar); }
        }
        if ((it & 63) == 63) {  // bound the number of uncommitted instructions like a real pipeline does
          if constexpr (PAIR) tc_commit2_mc(&done_bar); else tc_commit(&done_bar);
          mbar_wait(&done_bar, (it >> 6) & 1);
        }
      }
      long long t1 = clock64();
      p.cycles[blockIdx.x] = t1 - t0;
      stop = 1;
      if constexpr (PAIR) {  // tell the peer CTA (its interference warps keep running while the leader issues)
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, 1;" : "=r"(remote) : "r"(smem_u32(const_cast<int*>(&stop))));
        asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(1) : "memory");
      }
    }
  } else if (warp == 2 && lane == 0 && p.tma) {
    // four 16 KB boxes in flight, re-issued as they land (operand refills of a real pipeline: ~40 B/clk)
    int n = 0;
    for (int s = 0; s < 4; ++s, ++n) {
      mbar_expect_tx(&ring_bar[s], 16384);
      tma_load_2d(sX + s * 16384, &tm, &ring_bar[s], 0, (n * 128 + blockIdx.x * 4096) & 0x3ffff);
    }
    int s = 0;
    uint32_t ph = 0;
    while (true) {
      mbar_wait(&ring_bar[s], ph);
      if (stop) break;
      mbar_expect_tx(&ring_bar[s], 16384);
      tma_load_2d(sX + s * 16384, &tm, &ring_bar[s], 0, (n * 128 + blockIdx.x * 4096) & 0x3ffff);
      ++n;
      if (++s == 4) { s = 0; ph ^= 1; }
    }
    // drain the loads still in flight before the CTA may exit
    for (int k = 1; k < 4; ++k) {
      if (++s == 4) { s = 0; ph ^= 1; }
      mbar_wait(&ring_bar[s], ph);
    }
  } else if (warp >= 4 && warp < 4 + p.xf_warps) {
    // GroupNorm + SiLU transform stand-in: 16-byte units, y = silu(a x + b) in place
    const int tid = threadIdx.x - 128;
    int off = tid * 16;
    const int span = 48 * 1024;
    while (!stop) {
      uint4* u = reinterpret_cast<uint4*>(sX + 16384 * (p.tma ? 0 : 0) + (off % span));
      uint4 v = *u;
      float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y), f2 = unpack_bf16x2(v.z), f3 = unpack_bf16x2(v.w);
      f0.x = silu_f(fmaf(f0.x, 1.01f, 0.1f)); f0.y = silu_f(fmaf(f0.y, 1.01f, 0.1f));
      f1.x = silu_f(fmaf(f1.x, 1.01f, 0.1f)); f1.y = silu_f(fmaf(f1.y, 1.01f, 0.1f));
      f2.x = silu_f(fmaf(f2.x, 1.01f, 0.1f)); f2.y = silu_f(fmaf(f2.y, 1.01f, 0.1f));
      f3.x = silu_f(fmaf(f3.x, 1.01f, 0.1f)); f3.y = silu_f(fmaf(f3.y, 1.01f, 0.1f));
      v.x = pack_bf16x2(f0.x, f0.y); v.y = pack_bf16x2(f1.x, f1.y); v.z = pack_bf16x2(f2.x, f2.y); v.w = pack_bf16x2(f3.x, f3.y);
      *u = v;
      off += p.xf_warps * 32 * 16;
    }
  }
  __syncthreads();
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ---- lean issue loop: the whole warp executes it with warp-uniform values (descriptors advance by a 32-bit add in the
// uniform datapath), one elected lane issues. probe_kernel issues from a divergent single-lane branch, where every MMA
// operand goes through an ELECT / R2UR.BROADCAST sequence (~20 SASS instructions per MMA).
template <bool PAIR>
__global__ void __launch_bounds__(256, 1) probe_lean_kernel(const ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  uint32_t rank = 0;
  if constexpr (PAIR) rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i * 2654435761u >> 20 & 0x00ff00ffu);
  if (threadIdx.x == 0) { mbar_init(&done_bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (warp == 1) {
    if constexpr (PAIR) { tmem_alloc2(&tmem_slot, 512); tmem_relinquish2(); }
    else { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_slot, 0);
  if (warp == 0 && (!PAIR || rank == 0)) {
    const bool leader = elect_one();
    const uint64_t desc0 = umma_desc_sw128(0, 16, 1024);
    const uint32_t hi = static_cast<uint32_t>(desc0 >> 32), lo = static_cast<uint32_t>(desc0);
    auto dsc = [&](uint32_t l) { return (static_cast<uint64_t>(hi) << 32) | l; };
    const uint32_t a_lo = lo | (smem_u32(smem) >> 4), b_lo = lo | (smem_u32(smem + 80 * 1024) >> 4);
    const uint32_t id0 = umma_idesc_bf16(PAIR ? 256 : 128, p.pattern[0], 0, 0);
    const uint32_t id1 = umma_idesc_bf16(PAIR ? 256 : 128, p.pattern[1] ? p.pattern[1] : p.pattern[0], 0, 0);
    const long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if constexpr (PAIR) tc_mma2_ss(tmem_base, dsc(a_lo + 2 * k), dsc(b_lo + 2 * k), id0, 1);
          else tc_mma_ss(tmem_base, dsc(a_lo + 2 * k), dsc(b_lo + 2 * k), id0, 1);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if constexpr (PAIR) tc_mma2_ss(tmem_base + 256, dsc(a_lo + 8 + 2 * k), dsc(b_lo + 256 + 2 * k), id1, 1);
          else tc_mma_ss(tmem_base + 256, dsc(a_lo + 8 + 2 * k), dsc(b_lo + 256 + 2 * k), id1, 1);
        }
      }
      if ((it & 63) == 63) {
        if (leader) { if constexpr (PAIR) tc_commit2_mc(&done_bar); else tc_commit(&done_bar); }
        mbar_wait(&done_bar, (it >> 6) & 1);
      }
    }
    const long long t1 = clock64();
    if (leader) p.cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

static void run_lean(const char* name, int pair, int n0, int n1, long long* d_cycles) {
  ProbeParams p{};
  p.iters = 4096; p.cycles = d_cycles; p.pattern[0] = n0; p.pattern[1] = n1;
  cudaMemset(d_cycles, 0, 148 * sizeof(long long));
  cudaFuncSetAttribute(probe_lean_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  cudaFuncSetAttribute(probe_lean_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = kSmem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = pair ? 1 : 0;
  cudaError_t err = cudaSuccess;
  if (pair) err = cudaLaunchKernelEx(&cfg, probe_lean_kernel<true>, p);
  else { probe_lean_kernel<false><<<148, 256, kSmem>>>(p); err = cudaGetLastError(); }
  cudaError_t err2 = cudaDeviceSynchronize();
  if (err != cudaSuccess || err2 != cudaSuccess) { printf("%-46s FAILED: %s / %s\n", name, cudaGetErrorString(err), cudaGetErrorString(err2)); return; }
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), d_cycles, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  double sum = 0; int cnt = 0;
  for (long long c : h) if (c > 0) { sum += c; ++cnt; }
  const double per_instr = sum / cnt / p.iters / 8.0;
  const double nominal = (n0 + (n1 ? n1 : n0)) / 2.0 / 2.0;  // N / 2 cycles per K = 16 instruction at full rate
  printf("LEAN %-40s : %6.1f cycles / instruction (nominal %5.1f) -> tensor pipe %5.1f %%\n", name, per_instr, nominal, 100.0 * nominal / per_instr);
}

static double run(const char* name, int pair, std::initializer_list<int> pat, int a_views, int xf, int tma,
                  const CUtensorMap& tm, long long* d_cycles, int d_mode = 0, int commit_each = 0, int a_tmem = 0,
                  int a_mn = 0, int b_mn = 0) {
  ProbeParams p{};
  p.a_mn = a_mn; p.b_mn = b_mn;
  p.a_tmem = a_tmem;
  p.d_mode = d_mode;
  p.commit_each = commit_each;
  p.pair = pair; p.a_views = a_views; p.iters = 4096; p.xf_warps = xf; p.tma = tma; p.cycles = d_cycles;
  int j = 0;
  double nominal = 0;  // tensor-pipe cycles of one pattern at full rate: N / 2 per K = 16 instruction, 4 per block
  for (int n : pat) { p.pattern[j++] = n; nominal += 4 * n / 2.0; }
  cudaMemset(d_cycles, 0, 148 * sizeof(long long));
  cudaFuncSetAttribute(probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  cudaFuncSetAttribute(probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = kSmem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = pair ? 1 : 0;
  cudaEventRecord(e0);
  cudaError_t err = cudaSuccess;
  if (pair) err = cudaLaunchKernelEx(&cfg, probe_kernel<true>, tm, p);
  else { probe_kernel<false><<<148, 256, kSmem>>>(tm, p); err = cudaGetLastError(); }
  cudaEventRecord(e1);
  cudaError_t err2 = cudaDeviceSynchronize();
  if (err != cudaSuccess || err2 != cudaSuccess) {
    printf("%-46s FAILED: %s / %s\n", name, cudaGetErrorString(err), cudaGetErrorString(err2));
    return 0;
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), d_cycles, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  double sum = 0; int cnt = 0;
  for (long long c : h) if (c > 0) { sum += c; ++cnt; }
  const double cyc = sum / cnt / p.iters;
  printf("%-46s a_views=%d xf=%d tma=%d d=%d : %7.1f cycles / pattern (nominal %5.0f) -> tensor pipe %5.1f %%   [%.2f ms]\n", name,
         a_views, xf, tma, d_mode, cyc, nominal, 100.0 * nominal / cyc, ms);
  return cyc;
}

int main() {
  long long* d_cycles;
  cudaMalloc(&d_cycles, 148 * sizeof(long long));
  void* gbuf;
  cudaMalloc(&gbuf, 64ull << 20);
  cudaMemset(gbuf, 0, 64ull << 20);
  CUtensorMap tm;
  uint64_t dims[2] = {64, (64ull << 20) / 128};
  uint64_t str[1] = {128};
  uint32_t box[2] = {64, 128};
  if (make_tmap_bf16(&tm, gbuf, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("tensor map failed\n"); return 1; }
  if (getenv("PROBE_LEAN")) {
    for (int rep = 0; rep < 2; ++rep) {
      for (int n : {256, 192, 160, 128, 96, 80, 64, 32, 16}) {
        char nm[64];
        snprintf(nm, sizeof nm, "single N=%d", n); run_lean(nm, 0, n, 0, d_cycles);
        snprintf(nm, sizeof nm, "pair   N=%d", n); run_lean(nm, 1, n, 0, d_cycles);
      }
      run_lean("pair   N=256 / N=128 alternating blocks", 1, 256, 128, d_cycles);
      run_lean("single N=256 / N=128 alternating blocks", 0, 256, 128, d_cycles);
      run("(old issue loop) pair   N=128", 1, {128}, 1, 0, 0, tm, d_cycles);
      run("(old issue loop) single N=128", 0, {128}, 1, 0, 0, tm, d_cycles);
      run("(old issue loop) single N=64", 0, {64}, 1, 0, 0, tm, d_cycles);
    }
    return 0;
  }
  for (int rep = 0; rep < 2; ++rep) {  // first round warms the clocks
    run("single N=256", 0, {256}, 1, 0, 0, tm, d_cycles);
    run("single N=128", 0, {128}, 1, 0, 0, tm, d_cycles);
    run("single N=64", 0, {64}, 1, 0, 0, tm, d_cycles);
    run("single N=256,256 two accumulators", 0, {256, 256}, 2, 0, 0, tm, d_cycles);
    run("single N=128,128 two accumulators", 0, {128, 128}, 2, 0, 0, tm, d_cycles);
    run("single N=64,64 two accumulators", 0, {64, 64}, 2, 0, 0, tm, d_cycles);
    run("single N=256 A from TMEM", 0, {256, 256}, 2, 0, 0, tm, d_cycles, 0, 0, 1);
    run("single N=128 A from TMEM", 0, {128, 128}, 2, 0, 0, tm, d_cycles, 0, 0, 1);
    run("single N=64  A from TMEM", 0, {64, 64}, 2, 0, 0, tm, d_cycles, 0, 0, 1);
    run("single N=128 + commit each", 0, {128, 128}, 2, 0, 0, tm, d_cycles, 0, 1);
    run("single N=256 A MN-major (V^T as A)", 0, {256, 256}, 1, 0, 0, tm, d_cycles, 0, 0, 0, 1, 0);
    run("single N=128 A MN-major", 0, {128, 128}, 1, 0, 0, tm, d_cycles, 0, 0, 0, 1, 0);
    run("single N=128 B MN-major (V as B)", 0, {128, 128}, 1, 0, 0, tm, d_cycles, 0, 0, 0, 0, 1);
    run("single N=256 B MN-major", 0, {256, 256}, 1, 0, 0, tm, d_cycles, 0, 0, 0, 0, 1);
    run("pair   N=256", 1, {256}, 1, 0, 0, tm, d_cycles);
    run("pair   N=128", 1, {128}, 1, 0, 0, tm, d_cycles);
    run("pair   N=64", 1, {64}, 1, 0, 0, tm, d_cycles);
    run("pair   N=128 shifted views", 1, {128}, 6, 0, 0, tm, d_cycles);
    run("pair   N=256 shifted views", 1, {256}, 6, 0, 0, tm, d_cycles);
    run("pair   mix 128,256,256,128", 1, {128, 256, 256, 128}, 4, 0, 0, tm, d_cycles);
    run("single mix 128,256,256,128", 0, {128, 256, 256, 128}, 4, 0, 0, tm, d_cycles);
    run("pair   N=256 alternating accumulators", 1, {256}, 1, 0, 0, tm, d_cycles, 1);
    run("pair   N=128 alternating accumulators", 1, {128}, 1, 0, 0, tm, d_cycles, 1);
    run("pair   N=64  alternating accumulators", 1, {64}, 1, 0, 0, tm, d_cycles, 1);
    run("pair   N=128,128 two accumulators per 4", 1, {128, 128}, 2, 0, 0, tm, d_cycles, 0);
    run("pair   N=256,256 two accumulators per 4", 1, {256, 256}, 2, 0, 0, tm, d_cycles, 0);
    run("pair   mix, alternating accumulators", 1, {128, 256, 256, 128}, 4, 0, 0, tm, d_cycles, 1);
    run("pair   mix, one accumulator", 1, {128, 256, 256, 128}, 4, 0, 0, tm, d_cycles, 2);
    run("pair   mix, commit after every 4 MMAs", 1, {128, 256, 256, 128}, 4, 0, 0, tm, d_cycles, 0, 1);
    run("pair   N=128,128 commit after every 4 MMAs", 1, {128, 128}, 2, 0, 0, tm, d_cycles, 0, 1);
    run("pair   N=256,256 commit after every 4 MMAs", 1, {256, 256}, 2, 0, 0, tm, d_cycles, 0, 1);
    run("pair   N=128 + TMA", 1, {128}, 6, 0, 1, tm, d_cycles);
    run("pair   N=128 + transform", 1, {128}, 6, 4, 0, tm, d_cycles);
    run("pair   N=128 + TMA + transform", 1, {128}, 6, 4, 1, tm, d_cycles);
    run("pair   mix + TMA", 1, {128, 256, 256, 128}, 4, 0, 1, tm, d_cycles);
    run("pair   mix + TMA + transform", 1, {128, 256, 256, 128}, 4, 4, 1, tm, d_cycles);
    run("pair   N=256 + TMA + transform", 1, {256}, 6, 4, 1, tm, d_cycles);
    run("single N=256 + TMA", 0, {256}, 1, 0, 1, tm, d_cycles);
  }
  return 0;
}
