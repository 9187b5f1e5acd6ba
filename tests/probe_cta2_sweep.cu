// Stand-alone CTA-pair scoring sweep (experiment for DESIGN.md section 6, item 1; NOT the product kernel):
// the EPI_TOPK-style sweep of sweep.cuh -- maximum of every (query row, 128-item tile) of U . W^T, bf16,
// d = 128 -- rebuilt on tcgen05.mma.cta_group::2 so that its duration can be put beside the product's
// single-CTA sweep (0.76 ms for 4096 rows x 1M items on B200, profiles/r1o_ncu_summary.md).
//
// A cluster of two CTAs owns 512 stationary query rows (two 128-row tiles X_0, X_1 per CTA, as XT = 2 in
// sweep.cuh) and streams 128-item tiles; each CTA loads only ITS 64-row half of every streamed tile
// (.cta_group::2 TMA completing on the leader's mbarrier), the leader issues M = 256, N = 128 pair MMAs
// (A = X_x at the same offset in both CTAs, B = the two halves), S is double-buffered per stationary tile
// in tensor memory (4 x 128 columns), and the eight epilogue warps of each CTA reduce their own rows.
// Hand-shakes are the ones proven by probe_cta2_tma.cu: multicast commits for full->epilogue / stage
// release / X release, remote mbarrier arrivals (count 256) for S-buffer release.
//
//   bash tools/build_probe.sh && timeout 60 tests/_probe/probe_cta2_sweep
// prints a correctness line (N = 65536, bit-exact against a naive kernel; inputs are small integers) and a
// timing line (N = 1M).  Every mbarrier wait is bounded (watchdog trap, no hang).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_bf16.h>

#include "../recboard_b200/csrc/ptx.cuh"

using namespace rb;

constexpr int D = 128;
constexpr int NS = 6;                    // ring stages
constexpr int XTILE = 32768;             // one stationary tile: 2 K-chunks x (128 rows x 128 B)
constexpr int X_BYTES = 2 * XTILE;       // X_0 | X_1
constexpr int HALF_CHUNK = 8192;         // 64 rows x 128 B
constexpr int STAGE = 2 * HALF_CHUNK;    // this CTA's half of a streamed tile, both K-chunks
constexpr int SMEM_BYTES = X_BYTES + NS * STAGE + 1024;
constexpr int THREADS = 320;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
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
__device__ __forceinline__ void commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void mma_pair_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

struct Ctl {
  uint64_t full[NS], empty[NS];
  uint64_t x_full, x_empty;
  uint64_t s_full[4], s_empty[4];   // [stationary tile x][buffer]; s_empty is only used in the leader (count 256)
  uint32_t tmem_base;
};

struct Args {
  int n_rows, n_items;      // multiples of 512 / 128
  int n_units;              // n_rows / 512
  int n_tiles;              // n_items / 128
  int n_splits;
  float* tile_max;          // [n_rows][n_tiles]
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
pair_sweep_kernel(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_w, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* x_smem = smem;
  uint8_t* y_smem = smem + X_BYTES;
  __shared__ Ctl ctl;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int total_items = a.n_units * a.n_splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_u);
    tma_prefetch_desc(&tm_w);
    for (int s = 0; s < NS; ++s) { mbar_init(&ctl.full[s], 1); mbar_init(&ctl.empty[s], 1); }
    mbar_init(&ctl.x_full, 1);
    mbar_init(&ctl.x_empty, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&ctl.s_full[i], 1); mbar_init(&ctl.s_empty[i], 256); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl.tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = ctl.tmem_base;

  auto item_range = [&](int item, int& unit, int& t0, int& t1) {
    unit = item % a.n_units;   // split-major: concurrent clusters stream the same tiles (L2 reuse)
    const int split = item / a.n_units;
    t0 = static_cast<int>((static_cast<long long>(split) * a.n_tiles) / a.n_splits);
    t1 = static_cast<int>((static_cast<long long>(split + 1) * a.n_tiles) / a.n_splits);
  };

  if (warp == 0) {
    // ================================================================= producer (both CTAs)
    uint32_t it = 0, k = 0;
    const uint32_t x_full_leader = mapa_u32(smem_u32(&ctl.x_full), 0);
    for (int item = cluster; item < total_items; item += n_clusters, ++k) {
      int unit, t0, t1;
      item_range(item, unit, t0, t1);
      mbar_wait(&ctl.x_empty, (k & 1) ^ 1);
      if (elect_one()) {
        if (rank == 0) mbar_arrive_expect_tx(&ctl.x_full, 2 * X_BYTES);
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
          for (int c = 0; c < 2; ++c)
            tma_load_2d_pair(x_smem + x * XTILE + c * 16384, &tm_u, x_full_leader, c * 64,
                             (unit * 4 + static_cast<int>(rank) * 2 + x) * 128);
      }
      __syncwarp();
      for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t st = it % NS, ph = (it / NS) & 1;
        mbar_wait(&ctl.empty[st], ph ^ 1);
        if (elect_one()) {
          const uint32_t full_leader = mapa_u32(smem_u32(&ctl.full[st]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&ctl.full[st], 2 * STAGE);
#pragma unroll
          for (int c = 0; c < 2; ++c)
            tma_load_2d_pair(y_smem + st * STAGE + c * HALF_CHUNK, &tm_w, full_leader, c * 64,
                             t * 128 + static_cast<int>(rank) * 64);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer (leader CTA only)
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(FMT_BF16, 256, 128, 0, 0);
      constexpr uint32_t dhi = smem_desc_hi(1024);
      const uint32_t x_lo = smem_desc_lo(smem_u32(x_smem), 16), y_lo = smem_desc_lo(smem_u32(y_smem), 16);
      uint32_t it = 0, k = 0;
      for (int item = cluster; item < total_items; item += n_clusters, ++k) {
        int unit, t0, t1;
        item_range(item, unit, t0, t1);
        mbar_wait(&ctl.x_full, k & 1);
        for (int t = t0; t < t1; ++t, ++it) {
          const uint32_t st = it % NS, ph = (it / NS) & 1;
          const uint32_t buf = it & 1, sph = (it >> 1) & 1;
          mbar_wait(&ctl.s_empty[buf], sph ^ 1);       // X_0's buffer: all 256 rows of it (both CTAs) are read
          mbar_wait(&ctl.s_empty[2 + buf], sph ^ 1);   // X_1's
          mbar_wait(&ctl.full[st], ph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              const uint32_t d_tmem = tmem + (x * 2 + buf) * 128;
#pragma unroll
              for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  mma_pair_ss(d_tmem, smem_desc(dhi, x_lo + ((x * XTILE + c * 16384 + kk * 32) >> 4)),
                              smem_desc(dhi, y_lo + ((st * STAGE + c * HALF_CHUNK + kk * 32) >> 4)), idesc, (c | kk) != 0);
              commit_pair(&ctl.s_full[x * 2 + buf]);
            }
            commit_pair(&ctl.empty[st]);
            if (t + 1 == t1) commit_pair(&ctl.x_empty);
          }
          __syncwarp();
        }
      }
      // the last multicast commits (stage / X release) must have landed before either CTA may leave
      if (k > 0) mbar_wait(&ctl.x_empty, (k - 1) & 1);
    }
  } else {
    // ================================================================= epilogue (both CTAs)
    const int g = (warp - 2) >> 2;     // warpgroup == stationary tile X_g of this CTA
    const int q = warp & 3;            // tensor-memory lane quarter of this warp
    const int r = q * 32 + lane;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t s_empty_leader0 = mapa_u32(smem_u32(&ctl.s_empty[g * 2 + 0]), 0);
    const uint32_t s_empty_leader1 = mapa_u32(smem_u32(&ctl.s_empty[g * 2 + 1]), 0);
    uint32_t it = 0;
    for (int item = cluster; item < total_items; item += n_clusters) {
      int unit, t0, t1;
      item_range(item, unit, t0, t1);
      const int srow = (unit * 4 + static_cast<int>(rank) * 2 + g) * 128 + r;
      float* out = a.tile_max + static_cast<long long>(srow) * a.n_tiles;
      for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t buf = it & 1, sph = (it >> 1) & 1;
        mbar_wait(&ctl.s_full[g * 2 + buf], sph);
        tc_fence_after();
        float mx = -INFINITY;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t v[32];
          tmem_ld32(t_lane + (g * 2 + buf) * 128 + ch * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        tc_fence_before();
        mbar_arrive_remote(buf ? s_empty_leader1 : s_empty_leader0);
        if (srow < a.n_rows) out[t] = mx;
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// naive reference: one thread per (row, tile)
__global__ void naive_tile_max(const __nv_bfloat16* __restrict__ U, const __nv_bfloat16* __restrict__ W, int n_rows,
                               int n_tiles, float* __restrict__ out) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(n_rows) * n_tiles) return;
  const int row = static_cast<int>(idx / n_tiles), t = static_cast<int>(idx % n_tiles);
  float u[D];
  for (int k = 0; k < D; ++k) u[k] = __bfloat162float(U[static_cast<long long>(row) * D + k]);
  float mx = -INFINITY;
  for (int j = 0; j < 128; ++j) {
    const __nv_bfloat16* w = W + (static_cast<long long>(t) * 128 + j) * D;
    float s = 0.f;
    for (int k = 0; k < D; ++k) s += u[k] * __bfloat162float(w[k]);
    mx = fmaxf(mx, s);
  }
  out[idx] = mx;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void make_map(EncodeTiledFn enc, CUtensorMap* m, void* base, long long rows, int box_rows) {
  cuuint64_t gdim[2] = {D, static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {D * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", static_cast<int>(r)); exit(1); }
}

__global__ void fill_small_ints(__nv_bfloat16* x, long long n, unsigned seed) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned h = static_cast<unsigned>(i) * 2654435761u + seed;
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  x[i] = __float2bfloat16(static_cast<float>(static_cast<int>(h % 5u) - 2));   // -2..2: dot products are exact in fp32
}

static float run(EncodeTiledFn enc, int sms, int n_rows, int n_items, bool check) {
  __nv_bfloat16 *U, *W; float *tm, *ref = nullptr;
  const int n_tiles = n_items / 128;
  CK(cudaMalloc(&U, static_cast<size_t>(n_rows) * D * 2));
  CK(cudaMalloc(&W, static_cast<size_t>(n_items) * D * 2));
  CK(cudaMalloc(&tm, static_cast<size_t>(n_rows) * n_tiles * 4));
  fill_small_ints<<<(n_rows * D + 255) / 256, 256>>>(U, static_cast<long long>(n_rows) * D, 1u);
  fill_small_ints<<<static_cast<int>((static_cast<long long>(n_items) * D + 255) / 256), 256>>>(W, static_cast<long long>(n_items) * D, 77u);
  CK(cudaMemset(tm, 0xFF, static_cast<size_t>(n_rows) * n_tiles * 4));
  CUtensorMap tu, tw;
  make_map(enc, &tu, U, n_rows, 128);
  make_map(enc, &tw, W, n_items, 64);
  Args a{};
  a.n_rows = n_rows; a.n_items = n_items; a.n_units = n_rows / 512; a.n_tiles = n_tiles; a.tile_max = tm;
  const int clusters = sms / 2;
  a.n_splits = clusters / a.n_units > 0 ? clusters / a.n_units : 1;
  if (a.n_splits > n_tiles) a.n_splits = n_tiles;
  CK(cudaFuncSetAttribute(pair_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms = 0.f, best = 1e30f;
  for (int rep = 0; rep < (check ? 1 : 5); ++rep) {
    CK(cudaEventRecord(e0));
    pair_sweep_kernel<<<clusters * 2, THREADS, SMEM_BYTES>>>(tu, tw, a);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best;
  }
  if (check) {
    const long long n = static_cast<long long>(n_rows) * n_tiles;
    CK(cudaMalloc(&ref, n * 4));
    naive_tile_max<<<static_cast<int>((n + 127) / 128), 128>>>(U, W, n_rows, n_tiles, ref);
    CK(cudaDeviceSynchronize());
    std::vector<float> h1(n), h2(n);
    CK(cudaMemcpy(h1.data(), tm, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h2.data(), ref, n * 4, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (long long i = 0; i < n; ++i) bad += !(h1[i] == h2[i]);
    printf("{\"probe\": \"cta-pair sweep, correctness\", \"rows\": %d, \"items\": %d, \"splits\": %d, \"wrong_tile_maxima\": %lld, "
           "\"of\": %lld, \"ok\": %s}\n", n_rows, n_items, a.n_splits, bad, n, bad == 0 ? "true" : "false");
    CK(cudaFree(ref));
  } else {
    const double flop = 2.0 * n_rows * static_cast<double>(n_items) * D;
    printf("{\"probe\": \"cta-pair sweep, timing\", \"rows\": %d, \"items\": %d, \"splits\": %d, \"clusters\": %d, \"best_ms\": %.3f, "
           "\"last_ms\": %.3f, \"tflops\": %.0f, \"single_cta_product_sweep_ms\": 0.76}\n",
           n_rows, n_items, a.n_splits, clusters, best, ms, flop / (best * 1e-3) / 1e12);
  }
  fflush(stdout);
  CK(cudaFree(U)); CK(cudaFree(W)); CK(cudaFree(tm));
  return best;
}

int main() {
  CK(cudaFree(nullptr));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess) { fprintf(stderr, "no cuTensorMapEncodeTiled\n"); return 1; }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  run(enc, p.multiProcessorCount, 1024, 65536, true);      // 2 units x 37 splits
  run(enc, p.multiProcessorCount, 4096, 65536, true);      // 8 units x 9 splits
  run(enc, p.multiProcessorCount, 4096, 1000064, false);   // the bench shape (N rounded up to a multiple of 128)
  return 0;
}
