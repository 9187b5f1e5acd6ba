// Deterministic scatter-add of rows (dense embedding backward: autograd of `self.Item.embeddings(seqs)`,
// SASRec/main.py:183, run at :249) in ONE cooperative launch:
//
//   dst[idx[i] - base, :] += alpha * w[i] * src[i / row_div, :]      (fixed summation order => bitwise reproducible)
//
//   phase 0   keys (row id; invalid ids and the padding row get the sentinel n_rows and sort last) + first histogram
//   per pass  (11-bit digits: 2 passes up to 4M rows, 3 beyond)   histogram -> grid sync -> exclusive offsets ->
//             grid sync -> stable scatter (ranks from __match_any_sync: equal keys keep their input order) -> grid sync
//   segments  every warp walks 32 sorted positions in order (row loads batched 8 deep, adds in order); a run of
//             equal ids inside the block is added to the table by this warp alone, a run that began earlier goes to
//             lead[b], one that continues goes to trail[b] -> grid sync -> the warp that holds the START of a split
//             run adds trail[b] + lead[b+1] + ... in block order.  Hot rows (Zipf heads, thousands of entries) are
//             summed by many warps in parallel, yet in a fixed order.
//
// Round 1 ran these phases as 13 separate launches (0.23 ms for 204 800 ids, launch-latency-bound); the grid syncs of
// a cooperative launch cost a few microseconds each.  The destination is fp32 or bf16 (a bf16 parameter gets its
// embedding gradient accumulated in place: the fp32 row sum is rounded once and added with a bf16x2 reduction).
#pragma once
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include "ptx.cuh"

namespace rb {

namespace cg = cooperative_groups;

constexpr int RS_CHUNK = 512;         // keys per chunk (one warp ranks a chunk)
constexpr int RS_BITS = 11;
constexpr int RS_BINS = 1 << RS_BITS;
constexpr int SEG_BLOCK = 32;         // sorted positions per warp in the segment phase
constexpr int SC_THREADS = 256;
constexpr int SC_WARPS = SC_THREADS / 32;
constexpr int SC_SMEM_BYTES = SC_WARPS * RS_BINS * 4;   // per-warp digit offsets of the scatter phase (64 KB)

struct ScatterArgs {
  const void* src;          // (rows, d) T
  const int64_t* idx;       // n ids
  long long idx_base;
  void* dst;                // (n_rows, d) TG
  int n;
  long long n_rows;
  int d;
  long long padding_idx;
  float alpha;
  const float* alpha_dev;   // optional device scalar multiplied in
  float* cnt_out;           // optional: cnt_out[row] += cnt_alpha * alpha_dev * (#occurrences of row)
  float cnt_alpha;
  int row_div;              // entry p adds src[p / row_div] (gather-dot backward: row_div = K)
  const float* ew;          // optional per-entry weight
  // workspace
  uint32_t *k0, *v0, *k1, *v1;
  uint32_t* hist;           // [chunks][RS_BINS] (chunk-major: the scatter phase reads a chunk's offsets as one contiguous 8 KB run)
  uint32_t* totals;         // [passes][RS_BINS], zeroed by the host before the launch
  float *lead, *trail;      // [seg_blocks][d]
  int passes;
};

template <typename T>
__device__ __forceinline__ float4 sc_load4(const T* p) {
  if constexpr (sizeof(T) == 4) {
    return *reinterpret_cast<const float4*>(p);
  } else {
    const uint2 raw = *reinterpret_cast<const uint2*>(p);
    return make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xFFFF0000u),
                       __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xFFFF0000u));
  }
}
template <typename T>
__device__ __forceinline__ void sc_store4(T* p, float4 v) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = v;
  } else {
    uint2 pk;
    pk.x = pack_bf16x2(v.x, v.y);
    pk.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = pk;
  }
}
// dst[0..3] += al * acc as fire-and-forget reductions (RED.ADD.F32 / RED.ADD.BF16x2): every (row, column) of the table
// receives exactly ONE such add per call -- a row's entries are summed in registers first, by one warp -- so the result
// does not depend on any ordering, and no warp waits for a table row to come back from DRAM (the read-modify-write
// version spent 50 us of 160 on those round trips).  bf16 tables: the fp32 row sum is rounded to bf16, then added.
template <typename TG>
__device__ __forceinline__ void sc_axpy4(TG* dst, float al, float4 acc) {
  if constexpr (sizeof(TG) == 4) {
    atomicAdd(dst + 0, al * acc.x); atomicAdd(dst + 1, al * acc.y);
    atomicAdd(dst + 2, al * acc.z); atomicAdd(dst + 3, al * acc.w);
  } else {
    __nv_bfloat162* d2 = reinterpret_cast<__nv_bfloat162*>(dst);
    atomicAdd(d2 + 0, __floats2bfloat162_rn(al * acc.x, al * acc.y));
    atomicAdd(d2 + 1, __floats2bfloat162_rn(al * acc.z, al * acc.w));
  }
}

template <typename T, typename TG>
__global__ void __launch_bounds__(SC_THREADS)
scatter_add_coop_kernel(const ScatterArgs a) {
  extern __shared__ uint32_t sc_smem[];
  __shared__ uint32_t scan_part[SC_THREADS];
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  const int n = a.n;
  const int chunks = (n + RS_CHUNK - 1) / RS_CHUNK;
  const int gwarp = blockIdx.x * SC_WARPS + wib, n_gwarps = gridDim.x * SC_WARPS;
  uint32_t *kin = a.k0, *vin = a.v0, *kout = a.k1, *vout = a.v1;
#ifdef RB_X_SCTIME
  unsigned long long tq[16]; int nq = 0;
  auto stamp = [&]() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); if (nq < 16) tq[nq++] = t; };
  stamp();
#define SC_STAMP() stamp()
#else
#define SC_STAMP()
#endif

  for (int pass = 0; pass < a.passes; ++pass) {
    const int shift = pass * RS_BITS;
    uint32_t* totals = a.totals + pass * RS_BINS;
    // ---- histogram of this pass's digit per chunk (pass 0 also makes the keys)
    for (int chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x) {
      uint32_t* h = sc_smem;
      for (int i = tid; i < RS_BINS; i += SC_THREADS) h[i] = 0u;
      __syncthreads();
      const int beg = chunk * RS_CHUNK, end = min(n, beg + RS_CHUNK);
      for (int i = beg + tid; i < end; i += SC_THREADS) {
        uint32_t key;
        if (pass == 0) {
          const long long r = a.idx[i] - a.idx_base;
          key = (r >= 0 && r < a.n_rows && r != a.padding_idx) ? static_cast<uint32_t>(r) : static_cast<uint32_t>(a.n_rows);
          kin[i] = key;
          vin[i] = static_cast<uint32_t>(i);
        } else {
          key = kin[i];
        }
        atomicAdd(&h[(key >> shift) & (RS_BINS - 1)], 1u);
      }
      __syncthreads();
      for (int i = tid; i < RS_BINS; i += SC_THREADS) {
        const uint32_t c = h[i];
        a.hist[static_cast<long long>(chunk) * RS_BINS + i] = c;
        if (c != 0u) atomicAdd(&totals[i], c);   // integer adds: order-independent
      }
      __syncthreads();
    }
    grid.sync(); SC_STAMP();
    // ---- exclusive offsets: hist[digit][chunk] <- (keys with a smaller digit) + (same digit in earlier chunks)
    {
      uint32_t* excl = sc_smem;   // exclusive scan of the digit totals, redone by every block (2048 values)
      constexpr int PER = RS_BINS / SC_THREADS;
      uint32_t loc[PER];
      uint32_t s = 0;
#pragma unroll
      for (int j = 0; j < PER; ++j) { loc[j] = totals[tid * PER + j]; s += loc[j]; }
      scan_part[tid] = s;
      __syncthreads();
      for (int o = 1; o < SC_THREADS; o <<= 1) {
        const uint32_t v = (tid >= o) ? scan_part[tid - o] : 0u;
        __syncthreads();
        scan_part[tid] += v;
        __syncthreads();
      }
      uint32_t run = scan_part[tid] - s;
#pragma unroll
      for (int j = 0; j < PER; ++j) { excl[tid * PER + j] = run; run += loc[j]; }
      __syncthreads();
      for (int dg = gwarp; dg < RS_BINS / 4; dg += n_gwarps) {   // one warp per group of four digits (16-byte accesses)
        uint4* col = reinterpret_cast<uint4*>(a.hist) + dg;       // the group's counters of chunk c sit at col[c * RS_BINS / 4]
        uint4 base = make_uint4(excl[4 * dg], excl[4 * dg + 1], excl[4 * dg + 2], excl[4 * dg + 3]);
        for (int c0 = 0; c0 < chunks; c0 += 32) {
          const int c = c0 + lane;
          const uint4 v = (c < chunks) ? col[static_cast<long long>(c) * (RS_BINS / 4)] : make_uint4(0u, 0u, 0u, 0u);
          uint4 incl = v;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t ux = __shfl_up_sync(0xffffffffu, incl.x, o), uy = __shfl_up_sync(0xffffffffu, incl.y, o);
            const uint32_t uz = __shfl_up_sync(0xffffffffu, incl.z, o), uw = __shfl_up_sync(0xffffffffu, incl.w, o);
            if (lane >= o) { incl.x += ux; incl.y += uy; incl.z += uz; incl.w += uw; }
          }
          if (c < chunks)
            col[static_cast<long long>(c) * (RS_BINS / 4)] =
                make_uint4(base.x + incl.x - v.x, base.y + incl.y - v.y, base.z + incl.z - v.z, base.w + incl.w - v.w);
          base.x += __shfl_sync(0xffffffffu, incl.x, 31); base.y += __shfl_sync(0xffffffffu, incl.y, 31);
          base.z += __shfl_sync(0xffffffffu, incl.z, 31); base.w += __shfl_sync(0xffffffffu, incl.w, 31);
        }
      }
      __syncthreads();
    }
    grid.sync(); SC_STAMP();
    // ---- stable scatter: one warp per chunk, the chunk's digit offsets in shared memory
    {
      uint32_t* base = sc_smem + wib * RS_BINS;
      for (int chunk = gwarp; chunk < chunks; chunk += n_gwarps) {
        {
          const uint4* hsrc = reinterpret_cast<const uint4*>(a.hist + static_cast<long long>(chunk) * RS_BINS);
          uint4* hdst = reinterpret_cast<uint4*>(base);
#pragma unroll 4
          for (int i = lane; i < RS_BINS / 4; i += 32) hdst[i] = hsrc[i];
        }
        __syncwarp();
        const int beg = chunk * RS_CHUNK, end = min(n, beg + RS_CHUNK);
        constexpr int PER = RS_CHUNK / 32;
        uint32_t kreg[PER], vreg[PER];   // the whole chunk in registers: the ranking loop never waits on memory
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const int i = beg + j * 32 + lane;
          kreg[j] = (i < end) ? kin[i] : 0u;
          vreg[j] = (i < end) ? vin[i] : 0u;
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const int i = beg + j * 32 + lane;
          const bool ok = i < end;
          const uint32_t k = kreg[j];
          const uint32_t dgt = ok ? ((k >> shift) & (RS_BINS - 1)) : RS_BINS + lane;  // inactive lanes never match
          const uint32_t peers = __match_any_sync(0xffffffffu, dgt);
          const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
          uint32_t pos = 0;
          if (ok) pos = base[dgt] + rank;
          __syncwarp();
          if (ok && rank == __popc(peers) - 1) base[dgt] += __popc(peers);  // last peer bumps the counter
          __syncwarp();
          if (ok) { kout[pos] = k; vout[pos] = vreg[j]; }
        }
        __syncwarp();
      }
    }
    grid.sync(); SC_STAMP();
    { uint32_t* t = kin; kin = kout; kout = t; t = vin; vin = vout; vout = t; }
  }

  // =================================================================== segment sums over the sorted list (kin, vin)
  const uint32_t* keys = kin;
  const uint32_t* perm = vin;
  const T* src = static_cast<const T*>(a.src);
  TG* dst = static_cast<TG*>(a.dst);
  const int d = a.d;
  const int seg_blocks = (n + SEG_BLOCK - 1) / SEG_BLOCK;
  const float adev = (a.alpha_dev != nullptr) ? __ldg(a.alpha_dev) : 1.f;
  const float al = a.alpha * adev;
  for (int b = gwarp; b < seg_blocks; b += n_gwarps) {
    const int p0 = b * SEG_BLOCK;
    const int cnt = min(SEG_BLOCK, n - p0);
    const uint32_t key = (lane < cnt) ? keys[p0 + lane] : 0xFFFFFFFFu;
    const uint32_t first_key = __shfl_sync(0xffffffffu, key, 0);
    if (static_cast<long long>(first_key) >= a.n_rows) continue;  // sorted: nothing valid in this block
    const uint32_t my_perm = (lane < cnt) ? perm[p0 + lane] : 0u;
    const uint32_t my_src = (a.row_div > 1) ? my_perm / static_cast<uint32_t>(a.row_div) : my_perm;
    const float my_w = (a.ew != nullptr && lane < cnt) ? __ldg(a.ew + my_perm) : 1.f;
    const uint32_t prev_key = (p0 > 0) ? keys[p0 - 1] : 0xFFFFFFFFu;
    const uint32_t next_key = (p0 + SEG_BLOCK < n) ? keys[p0 + SEG_BLOCK] : 0xFFFFFFFFu;
    const uint32_t up = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = (lane == 0) || (key != up);
    const uint32_t heads = __ballot_sync(0xffffffffu, head && lane < cnt);
    const uint32_t valid = __ballot_sync(0xffffffffu, lane < cnt && static_cast<long long>(key) < a.n_rows);
    for (int c0 = 0; c0 < d; c0 += 128) {  // d % 4 == 0
      const int c = c0 + lane * 4;
      const bool col_ok = c < d;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int e0 = 0; e0 < cnt; e0 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t pr = __shfl_sync(0xffffffffu, my_src, (e0 + u) & 31);
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (col_ok && ((valid >> (e0 + u)) & 1u)) v[u] = sc_load4<T>(src + static_cast<long long>(pr) * d + c);
          if (a.ew != nullptr) {  // warp-uniform
            const float w = __shfl_sync(0xffffffffu, my_w, (e0 + u) & 31);
            v[u].x *= w; v[u].y *= w; v[u].z *= w; v[u].w *= w;
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = e0 + u;
          if (e < cnt && ((valid >> e) & 1u)) {
            acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
            const bool run_ends = (e + 1 == cnt) || ((heads >> (e + 1)) & 1u);
            if (run_ends) {  // warp-uniform
              const uint32_t k = __shfl_sync(0xffffffffu, key, e);
              const int j0 = 31 - __clz(heads & ((2u << e) - 1u));            // start of this run inside the block
              const bool began_before = (j0 == 0) && (k == prev_key);
              const bool continues = (e + 1 == SEG_BLOCK) && (k == next_key);
              if (col_ok) {
                if (began_before) {
                  *reinterpret_cast<float4*>(a.lead + static_cast<long long>(b) * d + c) = acc;
                } else if (continues) {
                  *reinterpret_cast<float4*>(a.trail + static_cast<long long>(b) * d + c) = acc;
                } else {
                  sc_axpy4<TG>(dst + static_cast<long long>(k) * d + c, al, acc);
                }
              }
              if (c0 == 0 && a.cnt_out != nullptr && lane == 0 && !began_before && !continues)
                a.cnt_out[k] += a.cnt_alpha * adev * static_cast<float>(e - j0 + 1);
              acc = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
      }
    }
  }
  grid.sync(); SC_STAMP();
  // ---- runs split over several blocks: the warp whose block holds the START of the run sums the pieces in block order
  for (int b = gwarp; b < seg_blocks; b += n_gwarps) {
    const int p0 = b * SEG_BLOCK;
    if (p0 + SEG_BLOCK >= n) continue;  // a run can only continue out of a full block that has a successor
    const uint32_t key = keys[p0 + SEG_BLOCK - 1];
    if (static_cast<long long>(key) >= a.n_rows || keys[p0 + SEG_BLOCK] != key) continue;   // no run leaves this block
    const uint32_t kl = keys[p0 + lane];
    const uint32_t same = __ballot_sync(0xffffffffu, kl == key);
    const int j0 = __ffs(same) - 1;
    if (j0 == 0 && p0 > 0 && keys[p0 - 1] == key) continue;  // the run began earlier: not the owner
    // end of the run (first position whose key differs) by binary search over the sorted keys; hot rows span
    // thousands of blocks
    int lo = p0 + SEG_BLOCK, hi = n;   // keys[lo] == key is known
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (keys[mid] == key) lo = mid + 1; else hi = mid;
    }
    const int run_end = lo;                              // exclusive
    const int nb = (run_end - 1) / SEG_BLOCK;            // last block that holds a piece of the run
    const int run_len = run_end - (p0 + j0);
    for (int c = lane * 4; c < d; c += 128) {
      float4 acc = *reinterpret_cast<const float4*>(a.trail + static_cast<long long>(b) * d + c);
      int x = b + 1;
      for (; x + 8 <= nb + 1; x += 8) {   // eight partials in flight, added in block order
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float4*>(a.lead + static_cast<long long>(x + u) * d + c);
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
      }
      for (; x <= nb; ++x) {
        const float4 v = *reinterpret_cast<const float4*>(a.lead + static_cast<long long>(x) * d + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      sc_axpy4<TG>(dst + static_cast<long long>(key) * d + c, al, acc);
    }
    if (a.cnt_out != nullptr && lane == 0) a.cnt_out[key] += a.cnt_alpha * adev * static_cast<float>(run_len);
  }
#ifdef RB_X_SCTIME
  grid.sync(); stamp();
  if (blockIdx.x == 0 && tid == 0) {
    printf("scatter phases (us):");
    for (int i = 1; i < nq; ++i) printf(" %.1f", (tq[i] - tq[i - 1]) * 1e-3);
    printf("\n");
  }
#endif
}

}  // namespace rb
