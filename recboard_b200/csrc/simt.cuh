// HBM-bound helper kernels around the sweep engine: row gather, deterministic scatter-add
// (stable LSD radix sort + in-order segment sums), operand preparation, partial merges.
#pragma once
#include <cuda_bf16.h>
#include "ptx.cuh"

namespace rb {

constexpr float MASKED_SCORE_F = -1e23f;  // UniSRec/main.py:413

// ------------------------------------------------------------------------------ gather
// out[i,:] = table[idx[i],:] moved as 16-byte vectors; `vpr` = vectors per row.
// (reference: self.Item.embeddings(seqs), SASRec/main.py:183)
__global__ void gather_rows_kernel(const uint4* __restrict__ table, const int64_t* __restrict__ idx,
                                   uint4* __restrict__ out, long long n_idx, long long n_rows, int vpr) {
  const long long total = n_idx * vpr;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; t + 3 * stride < total; t += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long tt = t + u * stride;
      const long long row = tt / vpr;
      const int c = static_cast<int>(tt - row * vpr);
      const long long src = __ldg(idx + row);
      v[u] = (src >= 0 && src < n_rows) ? __ldg(table + src * vpr + c) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) out[t + u * stride] = v[u];
  }
  for (; t < total; t += stride) {
    const long long row = t / vpr;
    const int c = static_cast<int>(t - row * vpr);
    const long long src = __ldg(idx + row);
    out[t] = (src >= 0 && src < n_rows) ? __ldg(table + src * vpr + c) : make_uint4(0, 0, 0, 0);
  }
}

// -------------------------------------------------------------------------- radix sort
// Stable LSD radix sort of (key = row id, val = position) pairs, 8 bits per pass.
// One warp owns a contiguous chunk; ranks inside the chunk come from __match_any_sync so equal
// keys keep their input order (=> the segment sums below run in a fixed order).
constexpr int RS_CHUNK = 2048;

__global__ void rs_hist_kernel(const uint32_t* __restrict__ keys, int n, int shift, uint32_t* __restrict__ hist,
                               int n_chunks) {
  __shared__ uint32_t h[256];
  const int chunk = blockIdx.x;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const int beg = chunk * RS_CHUNK, end = min(n, beg + RS_CHUNK);
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x) atomicAdd(&h[(keys[i] >> shift) & 255], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i * n_chunks + chunk] = h[i];  // digit-major
}

// exclusive scan over hist[256 * n_chunks] (digit-major) by one block
__global__ void rs_scan_kernel(uint32_t* __restrict__ hist, int total) {
  __shared__ uint32_t part[1024];
  const int per = (total + 1023) / 1024;
  const int beg = threadIdx.x * per, end = min(total, beg + per);
  uint32_t s = 0;
  for (int i = beg; i < end; ++i) s += hist[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    uint32_t v = (threadIdx.x >= o) ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = part[threadIdx.x] - s;
  for (int i = beg; i < end; ++i) { const uint32_t v = hist[i]; hist[i] = run; run += v; }
}

__global__ void rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int n, int shift,
                                  const uint32_t* __restrict__ offs, int n_chunks) {
  __shared__ uint32_t base[256];
  const int chunk = blockIdx.x;  // one warp per block
  const int lane = threadIdx.x;
  for (int i = lane; i < 256; i += 32) base[i] = offs[i * n_chunks + chunk];
  __syncwarp();
  const int beg = chunk * RS_CHUNK, end = min(n, beg + RS_CHUNK);
  for (int i0 = beg; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    const bool ok = i < end;
    const uint32_t k = ok ? keys_in[i] : 0u;
    const uint32_t dgt = ok ? ((k >> shift) & 255u) : 256u + lane;  // inactive lanes never match
    const uint32_t peers = __match_any_sync(0xffffffffu, dgt);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t pos = 0;
    if (ok) pos = base[dgt] + rank;
    __syncwarp();
    if (ok && rank == __popc(peers) - 1) base[dgt] += __popc(peers);  // last peer bumps the counter
    __syncwarp();
    if (ok) { keys_out[pos] = k; vals_out[pos] = vals_in[i]; }
  }
}

__global__ void rs_init_kernel(const int64_t* __restrict__ idx, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                               int n, long long n_rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const long long r = idx[i];
    keys[i] = (r >= 0 && r < n_rows) ? static_cast<uint32_t>(r) : static_cast<uint32_t>(n_rows);  // invalid ids sort last, skipped
    vals[i] = static_cast<uint32_t>(i);
  }
}

// One warp per sorted position that starts a run of equal row ids; sums the run in order.
// (reference: embedding_dense_backward, autograd of SASRec/main.py:183 run at :249)
template <typename T>
__global__ void scatter_segments_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ perm,
                                        const T* __restrict__ grad_out, float* __restrict__ grad_table, int n, int d,
                                        long long n_rows, long long padding_idx) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const uint32_t key = keys[w];
  if (static_cast<long long>(key) >= n_rows || static_cast<long long>(key) == padding_idx) return;
  if (w > 0 && keys[w - 1] == key) return;
  for (int c = lane * 4; c < d; c += 128) {  // d % 4 == 0
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = w; e < n && keys[e] == key; ++e) {
      const T* src = grad_out + static_cast<long long>(perm[e]) * d + c;
      if constexpr (sizeof(T) == 4) {
        const float4 v = *reinterpret_cast<const float4*>(src);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      } else {
        const uint2 raw = *reinterpret_cast<const uint2*>(src);
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
        const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
        acc.x += __low2float(a); acc.y += __high2float(a); acc.z += __low2float(b); acc.w += __high2float(b);
      }
    }
    float4* dst = reinterpret_cast<float4*>(grad_table + static_cast<long long>(key) * d + c);
    float4 o = *dst;
    o.x += acc.x; o.y += acc.y; o.z += acc.z; o.w += acc.w;
    *dst = o;
  }
}

// ------------------------------------------------------------------- operand preparation
// labels (int64, global) -> int32 local (label - base) or -1 when outside [0, n_items)
__global__ void labels_local_kernel(const int64_t* __restrict__ labels, long long base, long long n_items,
                                    int* __restrict__ out, int m, int m_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m_pad) {
    int v = -1;
    if (i < m) {
      const long long l = labels[i] - base;
      v = (l >= 0 && l < n_items) ? static_cast<int>(l) : -1;
    }
    out[i] = v;
  }
}

// lse (natural log) -> lse * log2(e), padded with +inf (=> exp2(x - inf) = 0 for padding rows)
__global__ void lse2_kernel(const float* __restrict__ lse, float* __restrict__ out, int m, int m_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m_pad) out[i] = (i < m) ? lse[i] * 1.4426950408889634f : INFINITY;
}

// fp32 (rows,d) -> [hi | lo] (rows, 2*dpad): hi = x with the 13 low mantissa bits cleared (exactly
// representable in TF32), lo = x - hi (exact in fp32; TF32 keeps its top 11 bits).
__global__ void split_hi_lo_kernel(const float* __restrict__ x, float* __restrict__ out, long long rows, int d, int dpad) {
  const long long total = rows * dpad;
  for (long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = t / dpad;
    const int c = static_cast<int>(t - r * dpad);
    float v = (c < d) ? x[r * d + c] : 0.f;
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    out[r * 2 * dpad + c] = hi;
    out[r * 2 * dpad + dpad + c] = v - hi;
  }
}

// seen CSR (int64, global ids) -> int32 CSR with local ids (col - id_base, clamped)
__global__ void csr_local_kernel(const int64_t* __restrict__ crow, const int64_t* __restrict__ col, long long id_base,
                                 int* __restrict__ crow32, int* __restrict__ col32, long long b, long long nnz) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i <= b) crow32[i] = static_cast<int>(crow[i]);
  if (i < nnz) {
    long long v = col[i] - id_base;
    v = v < -1 ? -1 : (v > 0x7ffffffe ? 0x7ffffffe : v);
    col32[i] = static_cast<int>(v);
  }
}

// ------------------------------------------------------------------------ partial merges
__global__ void lse_merge_kernel(const float* __restrict__ pm2, const float* __restrict__ pl, const float* __restrict__ pll,
                                 int n_splits, long long slot_stride, int m, float* __restrict__ row_max,
                                 float* __restrict__ row_sumexp, float* __restrict__ label_logit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  float mx = -INFINITY;
  for (int s = 0; s < n_splits; ++s) mx = fmaxf(mx, pm2[s * slot_stride + i]);
  float l = 0.f, ll = 0.f;
  for (int s = 0; s < n_splits; ++s) {
    l += pl[s * slot_stride + i] * exp2f(pm2[s * slot_stride + i] - mx);
    ll += pll[s * slot_stride + i];
  }
  row_max[i] = mx * 0.6931471805599453f;
  row_sumexp[i] = l;
  label_logit[i] = ll;
}

// out[i] = sum_s part[s][i]   (fixed order => deterministic)
__global__ void partial_sum_kernel(const float* __restrict__ part, int n_splits, long long n, float* __restrict__ out) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < n_splits; ++k) s += part[k * n + i];
    out[i] = s;
  }
}

// ------------------------------------------------------------------------- top-K merge
__device__ __forceinline__ uint32_t f32_orderable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_from_orderable(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}
// key order: score desc, then id asc
__device__ __forceinline__ unsigned long long topk_key(float v, int id) {
  return (static_cast<unsigned long long>(f32_orderable(v)) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(id));
}

template <int E>
__device__ __forceinline__ void warp_bitonic_stage64(unsigned long long (&v)[E], int k, int j) {
  const uint32_t lane = lane_id();
  if (j >= E) {
    const int lj = j / E;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = lane * E + e;
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, v[e], lj);
      const bool up = ((i & k) == 0);
      const bool lower = ((i & j) == 0);
      const bool take_max = (up == lower);
      const unsigned long long mx = v[e] > other ? v[e] : other;
      const unsigned long long mn = v[e] > other ? other : v[e];
      v[e] = take_max ? mx : mn;
    }
  } else {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if ((e & j) == 0) {
        const int i = lane * E + e;
        const bool up = ((i & k) == 0);
        const unsigned long long a = v[e], b = v[e + j];
        const unsigned long long hi = a > b ? a : b, lo = a > b ? b : a;
        v[e] = up ? hi : lo;
        v[e + j] = up ? lo : hi;
      }
    }
  }
}
template <int E>
__device__ __forceinline__ void warp_bitonic_sort64_desc(unsigned long long (&v)[E]) {
#pragma unroll
  for (int k = 2; k <= 32 * E; k <<= 1)
#pragma unroll
    for (int j = k >> 1; j >= 1; j >>= 1) warp_bitonic_stage64<E>(v, k, j);
}
// v is bitonic (as produced by max(best[i], new[n-1-i])): finish with the last merge network, descending
template <int E>
__device__ __forceinline__ void warp_bitonic_merge64_desc(unsigned long long (&v)[E]) {
#pragma unroll
  for (int j = 16 * E; j >= 1; j >>= 1) warp_bitonic_stage64<E>(v, 64 * E /* bit never set => all "up" */, j);
}

// One warp per query row: merge `n_lists` candidate lists of up to `cap` entries each
// (list l of row i at base + (l*list_stride + i)*cap, length cnt[l*list_stride+i] or `cap` when cnt==nullptr)
// into the K best, sorted (score desc, id asc). E*32 >= K.
template <int E>
__global__ void topk_merge_kernel(const float* __restrict__ vals, const int* __restrict__ ids, const int* __restrict__ cnt,
                                  int n_lists, long long list_stride, int cap, long long n_rows, int K, int id_add,
                                  float* __restrict__ out_vals, int* __restrict__ out_ids) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  unsigned long long best[E];
#pragma unroll
  for (int e = 0; e < E; ++e) best[e] = 0ull;
  for (int l = 0; l < n_lists; ++l) {
    const long long slot = l * list_stride + row;
    const int n = cnt ? cnt[slot] : cap;
    for (int base = 0; base < n; base += 32 * E) {
      unsigned long long cur[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int i = base + lane * E + e;
        cur[e] = 0ull;
        if (i < n) {
          const int id = ids[slot * cap + i];
          if (id >= 0) cur[e] = topk_key(vals[slot * cap + i], id);
        }
      }
      warp_bitonic_sort64_desc<E>(cur);
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const unsigned long long rev = __shfl_sync(0xffffffffu, cur[E - 1 - e], 31 - lane);
        best[e] = best[e] > rev ? best[e] : rev;
      }
      warp_bitonic_merge64_desc<E>(best);
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = lane * E + e;
    if (i < K) {
      const unsigned long long key = best[e];
      float v = MASKED_SCORE_F;
      int id = -1;
      if (key != 0ull) {
        v = f32_from_orderable(static_cast<uint32_t>(key >> 32));
        id = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFu)) + id_add;
      }
      out_vals[row * K + i] = v;
      out_ids[row * K + i] = id;
    }
  }
}

}  // namespace rb
