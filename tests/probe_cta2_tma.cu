// Protocol probe for the CTA-pair sweeps: the smallest complete 2-CTA GEMM
//     D[256 x 128] = A[256 x K] . B[128 x K]^T      (bf16 in, fp32 out, K = 64 * KT streamed through a 2-stage ring)
// with exactly the hand-shakes the rewritten sweeps need, checked against a host reference:
//   * both CTAs of a cluster load their own half of every operand tile (A rows [128r, +128), B rows [64r, +64))
//     with the cta_group::2 TMA form whose completion goes to the LEADER's `full[s]` mbarrier (mapa address);
//     the leader alone posts expect_tx for both halves;
//   * the leader's MMA warp issues tcgen05.mma.cta_group::2 (M = 256, N = 128) and frees a stage for BOTH
//     producers with one multicast commit on `empty[s]`; the accumulator-ready signal is a multicast commit too;
//   * every epilogue thread of both CTAs reads its own 128 x 128 half of D from its own tensor memory and then
//     arrives REMOTELY on the leader's `drained` mbarrier (count 256) -- the s_empty hand-shake of the sweeps.
// Every mbarrier wait is bounded (ptx.cuh watchdog traps instead of hanging).
//
//   bash tools/build_probe.sh && timeout 30 tests/_probe/probe_cta2_tma
// Prints one JSON line; "ok": true means bit-exact against the host (inputs are small integers).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_bf16.h>

#include "../recboard_b200/csrc/ptx.cuh"

using namespace rb;

constexpr int KT = 8;            // K tiles of 64
constexpr int NS = 2;            // ring stages
constexpr int A_STAGE = 16384;   // 128 rows x 128 B
constexpr int B_STAGE = 8192;    //  64 rows x 128 B
constexpr int SMEM_BYTES = NS * (A_STAGE + B_STAGE) + 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// 2-D tiled load into THIS CTA's shared memory; completion bytes go to `mbar_cluster_addr`, which may be the
// peer CTA's barrier (cta_group::2 form)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t mbar_cluster_addr,
                                                 int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t mbar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mbar_cluster_addr) : "memory");
}
__device__ __forceinline__ void commit_pair(uint64_t* bar) {   // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void mma_pair_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

struct Ctl {
  uint64_t full[NS], empty[NS], done, drained;
  uint32_t tmem_base;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 float* __restrict__ D, int* __restrict__ drained_flag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + NS * A_STAGE;
  __shared__ Ctl ctl;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < NS; ++s) { mbar_init(&ctl.full[s], 1); mbar_init(&ctl.empty[s], 1); }
    mbar_init(&ctl.done, 1);
    mbar_init(&ctl.drained, 256);
    fence_barrier_init();
  }
  if (warp == 1) {   // one warp of each CTA
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl.tmem_base)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of both CTAs are initialised before anything can signal them
  tc_fence_after();
  const uint32_t tmem = ctl.tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------- producer (both CTAs)
    for (int kt = 0; kt < KT; ++kt) {
      const int s = kt % NS, ph = (kt / NS) & 1;
      mbar_wait(&ctl.empty[s], ph ^ 1);   // own copy: the leader's commit is multicast
      if (elect_one()) {
        const uint32_t full_leader = mapa_u32(smem_u32(&ctl.full[s]), 0);
        if (rank == 0) mbar_arrive_expect_tx(&ctl.full[s], 2 * (A_STAGE + B_STAGE));   // both CTAs' halves
        tma_load_2d_pair(a_smem + s * A_STAGE, &tm_a, full_leader, kt * 64, static_cast<int>(rank) * 128);
        tma_load_2d_pair(b_smem + s * B_STAGE, &tm_b, full_leader, kt * 64, static_cast<int>(rank) * 64);
      }
      __syncwarp();
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------- MMA issuer (leader CTA only)
    constexpr uint32_t idesc = make_idesc(FMT_BF16, 256, 128, 0, 0);
    constexpr uint32_t dhi = smem_desc_hi(1024);
    const uint32_t a_lo = smem_desc_lo(smem_u32(a_smem), 16), b_lo = smem_desc_lo(smem_u32(b_smem), 16);
    for (int kt = 0; kt < KT; ++kt) {
      const int s = kt % NS, ph = (kt / NS) & 1;
      mbar_wait(&ctl.full[s], ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          mma_pair_ss(tmem, smem_desc(dhi, a_lo + ((s * A_STAGE + kk * 32) >> 4)),
                      smem_desc(dhi, b_lo + ((s * B_STAGE + kk * 32) >> 4)), idesc, (kt | kk) != 0);
        commit_pair(&ctl.empty[s]);
        if (kt == KT - 1) commit_pair(&ctl.done);
      }
      __syncwarp();
    }
  }
  __syncwarp();

  // ----------------------------------------------------------------- epilogue (all warps of both CTAs)
  mbar_wait(&ctl.done, 0);
  tc_fence_after();
  const int row = static_cast<int>(rank) * 128 + warp * 32 + lane;
  const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll 1
  for (int ch = 0; ch < 4; ++ch) {
    uint32_t v[32];
    tmem_ld32(t_row + ch * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 4)
      *reinterpret_cast<float4*>(D + row * 128 + ch * 32 + i) =
          make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
  }
  tc_fence_before();
  mbar_arrive_remote(mapa_u32(smem_u32(&ctl.drained), 0));   // 256 arrivals from the two CTAs
  if (rank == 0 && threadIdx.x == 0) {
    mbar_wait(&ctl.drained, 0);
    *drained_flag = 1;
  }
  cluster_sync_all();   // nobody leaves (or frees tensor memory) while the peer may still signal or read
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_map(EncodeTiledFn enc, CUtensorMap* m, void* base, int rows, int cols, int box_rows) {
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", static_cast<int>(r)); exit(1); }
}

int main() {
  const int M = 256, N = 128, K = 64 * KT;
  CK(cudaFree(nullptr));
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess) { fprintf(stderr, "no cuTensorMapEncodeTiled\n"); return 1; }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);

  std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K);
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return static_cast<float>(static_cast<int>((s >> 24) % 9) - 4); };   // -4..4
  for (int i = 0; i < M * K; ++i) { fA[i] = rnd(); hA[i] = __float2bfloat16(fA[i]); }
  for (int i = 0; i < N * K; ++i) { fB[i] = rnd(); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD; int* dflag;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dD, M * N * 4)); CK(cudaMalloc(&dflag, 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, M * N * 4)); CK(cudaMemset(dflag, 0, 4));
  CUtensorMap ta, tb;
  make_map(enc, &ta, dA, M, K, 128);
  make_map(enc, &tb, dB, N, K, 64);
  CK(cudaFuncSetAttribute(pair_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  pair_gemm_kernel<<<2, 128, SMEM_BYTES>>>(ta, tb, dD, dflag);   // one cluster (__cluster_dims__)
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hD(M * N);
  int flag = 0;
  CK(cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&flag, dflag, 4, cudaMemcpyDeviceToHost));
  double max_err = 0; int bad = 0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += static_cast<double>(fA[i * K + k]) * fB[j * K + k];
      const double e = fabs(ref - hD[i * N + j]);
      if (!(e == 0)) ++bad;
      if (e > max_err || e != e) max_err = e;
    }
  printf("{\"probe\": \"cta_group::2 TMA + MMA + multicast commit + remote arrive\", \"M\": %d, \"N\": %d, \"K\": %d, "
         "\"max_abs_err\": %g, \"wrong_elements\": %d, \"drained_flag\": %d, \"ok\": %s}\n",
         M, N, K, max_err, bad, flag, (bad == 0 && flag == 1) ? "true" : "false");
  return (bad == 0 && flag == 1) ? 0 : 2;
}
