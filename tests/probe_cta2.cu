// Raw tcgen05.mma issue-rate probe: one CTA (cta_group::1, M = 128) against a CTA pair
// (cta_group::2, M = 256, the streamed/B operand split across the two SMs of the pair).
//
// Why: every re-tiling of the fused CE passes lands at the single-CTA operand-fetch floor
// (profiles/r1_pair_kernel_experiments.md: 128x128x16 SS instructions cost 88 clk against 64 nominal,
// 128x64x16 75 clk, TS 128x128x16 93 clk).  A CTA pair halves the B-operand shared-memory reads per SM
// and instruction; this probe measures what that buys BEFORE the sweeps are rewritten for it.
//
// Timing only, no TMA: operands are a constant bf16 pattern (2^-7) written by the CTA itself, so
// D = (#MMAs x 16) x 2^-14 exactly, which checks that both CTAs of a pair really accumulated.
// One cluster per SM pair over the whole GPU (all 148 SMs busy, realistic clocks); the leader thread
// issues `reps` x 4 MMAs back to back (four K = 16 slices of a 64-wide K tile, like the sweeps), commits
// once and waits; cycles per instruction = clock64 difference / count.
//
// Build + run (sm_100a):  bash tools/build_probe.sh && timeout 60 tests/_probe/probe_cta2
// Not part of the product library; not a pytest module.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../recboard_b200/csrc/ptx.cuh"

using namespace rb;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int CTAS>
__device__ __forceinline__ void p_tmem_alloc(uint32_t* slot, uint32_t ncols) {
  if constexpr (CTAS == 1)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
}
template <int CTAS>
__device__ __forceinline__ void p_tmem_relinquish() {
  if constexpr (CTAS == 1)
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  else
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CTAS>
__device__ __forceinline__ void p_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (CTAS == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// completion of all MMAs issued so far by this thread -> mbarrier at the same offset in every CTA of `mask`
template <int CTAS>
__device__ __forceinline__ void p_commit(uint64_t* bar, uint16_t mask) {
  if constexpr (CTAS == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
template <int CTAS>
__device__ __forceinline__ void p_mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (CTAS == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CTAS>
__device__ __forceinline__ void p_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  if constexpr (CTAS == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

constexpr int A_BYTES = 16384;        // 128 rows x 64 bf16, K-major, one 128-byte swizzle row per row
constexpr int B_BYTES = 65536;        // K-major: up to 256 rows x 128 B; MN-major: up to 4 groups of 16 KB
constexpr int SMEM_BYTES = A_BYTES + B_BYTES + 1024;
constexpr uint32_t FILL = 0x3C003C00u;   // two bf16 2^-7

// CTAS: CTAs per MMA (1 | 2).  N: MMA N (columns of D; each CTA of a pair holds N / 2 rows of B).
// TS: A operand from tensor memory (the softmax tile of the CE passes) instead of shared memory.
// BMN: B operand MN-major (the transposed read MMA2 of the CE passes makes), else K-major.
template <int CTAS, int N, bool TS, bool BMN>
__global__ void __launch_bounds__(128, 1) probe_kernel(long long* __restrict__ cycles, float* __restrict__ d00, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + A_BYTES;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();

  for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = FILL;
  fence_proxy_async_smem();   // generic-proxy writes -> visible to the MMA's async-proxy reads
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {            // one warp of EACH CTA (the pair allocates the same columns in both SMs)
    p_tmem_alloc<CTAS>(&tmem_slot, 512);
    p_tmem_relinquish<CTAS>();
  }
  tc_fence_before();
  cluster_sync_all();         // also the CTA-wide barrier (every thread of every CTA takes part)
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;

  if (TS) {                   // A tile in tensor memory: 128 lanes x 64 bf16 = 32 columns at column 256
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = FILL;
    tmem_st32p(tmem + lane_base + 256, v);
    tmem_st_wait();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
  }

  if (warp == 0) {
    if (rank == 0) {          // the leader CTA issues for the pair
      constexpr uint32_t idesc = make_idesc(FMT_BF16, 128 * CTAS, N, 0, BMN ? 1 : 0);
      constexpr uint32_t dhi = smem_desc_hi(1024);
      const uint32_t a_lo = smem_desc_lo(smem_u32(a_smem), 16);
      const uint32_t b_lo = smem_desc_lo(smem_u32(b_smem), BMN ? 16384 : 16);
      constexpr uint32_t kstep = BMN ? 2048 : 32;   // bytes between K = 16 slices
      if (elect_one()) {
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t bd = smem_desc(dhi, b_lo + ((kk * kstep) >> 4));
            if (TS) p_mma_ts<CTAS>(tmem, tmem + 256 + kk * 8, bd, idesc, (r | kk) != 0);
            else    p_mma_ss<CTAS>(tmem, smem_desc(dhi, a_lo + ((kk * 32) >> 4)), bd, idesc, (r | kk) != 0);
          }
        }
        p_commit<CTAS>(&bar, static_cast<uint16_t>((1u << CTAS) - 1u));
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        cycles[blockIdx.x / CTAS] = t1 - t0;
      }
      __syncwarp();
    } else {
      mbar_wait(&bar, 0);     // the multicast arrival of the leader's commit
    }
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  {                           // D[row 0 of this CTA][column 0]: proves the MMAs accumulated in BOTH CTAs
    uint32_t v[32];
    tmem_ld32(tmem + lane_base, v);
    tmem_ld_wait();
    if (threadIdx.x == 0) d00[blockIdx.x] = __uint_as_float(v[0]);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) p_tmem_dealloc<CTAS>(tmem, 512);
}

// The MMA mix of one step of the fused CE passes (pair.cuh), issued back to back with no hand-shakes: per step
// 16 SS instructions (S_g[h] = X_g . Y_half^T: 2 stationary tiles x 2 K-chunks x 4 slices, N = 64) and 8 TS
// instructions (A_g += P_g[h] . Y_half: 2 tiles x 4 slices, N = 128, A = the S columns reinterpreted as bf16, B =
// the MN-major view of the same half tile).  Single CTA: the "raw" floor of the experiments file (1405 clk per
// step).  CTA pair: the same instruction count covers four stationary tiles (two per SM), the streamed half tile
// split by rows for MMA1 and by feature columns for MMA2.  Timing only (accumulators are not checked).
constexpr int MIX_X_BYTES = 65536;   // X0 | X1, two 16 KB K-chunks each
constexpr int MIX_Y_BYTES = 32768;   // one streamed tile: two 16 KB K-chunks of 128 rows
constexpr int MIX_SMEM_BYTES = MIX_X_BYTES + MIX_Y_BYTES + 1024;
template <int CTAS>
__global__ void __launch_bounds__(128, 1) mix_kernel(long long* __restrict__ cycles, float* __restrict__ d00, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* x_smem = smem;
  uint8_t* y_smem = smem + MIX_X_BYTES;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < (MIX_X_BYTES + MIX_Y_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = FILL;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    p_tmem_alloc<CTAS>(&tmem_slot, 512);
    p_tmem_relinquish<CTAS>();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    if (rank == 0) {
      constexpr uint32_t idesc1 = make_idesc(FMT_BF16, 128 * CTAS, 64, 0, 0);
      constexpr uint32_t idesc2 = make_idesc(FMT_BF16, 128 * CTAS, 128, 0, 1);
      constexpr uint32_t dhi = smem_desc_hi(1024);
      const uint32_t x_lo = smem_desc_lo(smem_u32(x_smem), 16);
      const uint32_t y_lo1 = smem_desc_lo(smem_u32(y_smem), 16);
      const uint32_t y_lo2 = smem_desc_lo(smem_u32(y_smem), 16384);
      constexpr uint32_t half1 = 8192 / CTAS;   // MMA1: rows [64h, 64h+64) of the tile, split by rows across a pair
      if (elect_one()) {
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
          const uint32_t h = r & 1;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const uint32_t s_tmem = tmem + g * 128 + h * 64;
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                p_mma_ss<CTAS>(s_tmem, smem_desc(dhi, x_lo + ((g * 32768 + c * 16384 + kk * 32) >> 4)),
                               smem_desc(dhi, y_lo1 + ((c * 16384 + h * half1 + kk * 32) >> 4)), idesc1, (c | kk) != 0);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)   // a pair holds 64 of the 128 feature columns per SM: one 16 KB group, no LBO step
              p_mma_ts<CTAS>(tmem + 256 + g * 128, s_tmem + kk * 8, smem_desc(dhi, y_lo2 + ((h * 8192 + kk * 2048) >> 4)),
                             idesc2, (r | kk) != 0);
          }
        }
        p_commit<CTAS>(&bar, static_cast<uint16_t>((1u << CTAS) - 1u));
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        cycles[blockIdx.x / CTAS] = t1 - t0;
      }
      __syncwarp();
    } else {
      mbar_wait(&bar, 0);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (threadIdx.x == 0) d00[blockIdx.x] = 0.f;
  if (warp == 0) p_tmem_dealloc<CTAS>(tmem, 512);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int CTAS, int N, bool TS, bool BMN>
static void run(int sms, int reps) {
  auto kern = probe_kernel<CTAS, N, TS, BMN>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const int grid = (sms / CTAS) * CTAS, clusters = grid / CTAS;
  long long* cyc; float* d00;
  CK(cudaMalloc(&cyc, clusters * sizeof(long long)));
  CK(cudaMalloc(&d00, grid * sizeof(float)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = SMEM_BYTES;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms = 0.f;
  for (int it = 0; it < 3; ++it) {   // the last launch is the one reported
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, kern, cyc, d00, reps));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  std::vector<long long> hc(clusters); std::vector<float> hd(grid);
  CK(cudaMemcpy(hc.data(), cyc, clusters * sizeof(long long), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hd.data(), d00, grid * sizeof(float), cudaMemcpyDeviceToHost));
  const double n_mma = 4.0 * reps;
  double sum = 0, mn = 1e30, mx = 0;
  for (long long c : hc) { const double v = c / n_mma; sum += v; mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
  const float expect = static_cast<float>(n_mma * 16.0 / 16384.0);
  int bad = 0;
  for (float v : hd) bad += (v != expect);
  const double flop = clusters * n_mma * 2.0 * (128.0 * CTAS) * N * 16.0;
  printf("{\"ctas\": %d, \"M\": %d, \"N\": %d, \"a\": \"%s\", \"b\": \"%s\", \"clk_per_mma\": %.1f, \"clk_min\": %.1f, "
         "\"clk_max\": %.1f, \"clk_per_mma_per_128x128x16\": %.1f, \"kernel_ms\": %.3f, \"tflops\": %.0f, "
         "\"d00_expected\": %g, \"ctas_with_wrong_d00\": %d}\n",
         CTAS, 128 * CTAS, N, TS ? "tmem" : "smem", BMN ? "mn-major" : "k-major", sum / clusters, mn, mx,
         (sum / clusters) * 128.0 / N, ms, flop / (ms * 1e-3) / 1e12, expect, bad);
  fflush(stdout);
  CK(cudaFree(cyc)); CK(cudaFree(d00));
}

template <int CTAS>
static void run_mix(int sms, int reps) {
  auto kern = mix_kernel<CTAS>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MIX_SMEM_BYTES));
  const int grid = (sms / CTAS) * CTAS, clusters = grid / CTAS;
  long long* cyc; float* d00;
  CK(cudaMalloc(&cyc, clusters * sizeof(long long)));
  CK(cudaMalloc(&d00, grid * sizeof(float)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = MIX_SMEM_BYTES;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms = 0.f;
  for (int it = 0; it < 3; ++it) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, kern, cyc, d00, reps));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  std::vector<long long> hc(clusters);
  CK(cudaMemcpy(hc.data(), cyc, clusters * sizeof(long long), cudaMemcpyDeviceToHost));
  double sum = 0;
  for (long long c : hc) sum += static_cast<double>(c) / reps;
  // executed flops per step and SM: 2 tiles x (128 x 64 x 128 scores + 128 x 128 x 64 accumulate) x 2 = 8.4 MFLOP
  const double flop = static_cast<double>(grid) * reps * 2.0 * (2.0 * 128 * 64 * 128 + 2.0 * 128 * 128 * 64);
  printf("{\"mix\": \"pair-kernel step: 16 SS N=64 + 8 TS N=128\", \"ctas\": %d, \"clk_per_step\": %.0f, "
         "\"clk_per_step_per_sm_work\": %.0f, \"kernel_ms\": %.3f, \"executed_tflops\": %.0f}\n",
         CTAS, sum / clusters, sum / clusters / CTAS, ms, flop / (ms * 1e-3) / 1e12);
  fflush(stdout);
  CK(cudaFree(cyc)); CK(cudaFree(d00));
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 2048;
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  fprintf(stderr, "%s, %d SMs, reps %d (x4 MMAs)\n", p.name, p.multiProcessorCount, reps);
  const int sms = p.multiProcessorCount;
  // single CTA, the instruction shapes the sweeps use today
  run<1, 64, false, false>(sms, reps);
  run<1, 128, false, false>(sms, reps);
  run<1, 256, false, false>(sms, reps);
  run<1, 128, true, true>(sms, reps);
  // CTA pair: M = 256, each SM reads its own 128 A rows and half of B
  run<2, 64, false, false>(sms, reps);
  run<2, 128, false, false>(sms, reps);
  run<2, 256, false, false>(sms, reps);
  run<2, 128, true, true>(sms, reps);
  run<2, 256, true, true>(sms, reps);
  // the fused CE passes' instruction mix, single CTA (known floor: ~1405 clk per step) against a pair
  run_mix<1>(sms, reps);
  run_mix<2>(sms, reps);
  return 0;
}
