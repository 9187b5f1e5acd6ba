// HBM-bound helper kernels around the sweep engine: row gather, operand preparation, partial merges, the top-K
// finishing kernels (the deterministic scatter-add lives in scatter.cuh).
#pragma once
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include "ptx.cuh"
#include "sweep.cuh"

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

// The same gather over a ROW-SHARDED table whose shards live on the GPUs of one box (input side):
// shard r holds the rows of global ids [starts[r], starts[r+1]) at shards[r] (the local shard, or a peer's mapped
// through CUDA IPC -- the loads then travel over NVLink).  Every rank reads the rows where they live: no collective,
// no replicated copy of the table.  Ids outside [starts[0], starts[n_shards]) (the padding id) give zero rows.
constexpr int PEER_MAX_SHARDS = 16;
struct PeerShards {
  const uint4* base[PEER_MAX_SHARDS];
  long long start[PEER_MAX_SHARDS + 1];
  int n;
};
__global__ void gather_rows_peers_kernel(const PeerShards sh, const int64_t* __restrict__ idx, uint4* __restrict__ out,
                                         long long n_idx, int vpr) {
  const long long total = n_idx * vpr;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  auto fetch = [&](long long tt) -> uint4 {
    const long long row = tt / vpr;
    const int c = static_cast<int>(tt - row * vpr);
    const long long id = __ldg(idx + row);
    if (id < sh.start[0] || id >= sh.start[sh.n]) return make_uint4(0, 0, 0, 0);
    int r = 0;
#pragma unroll
    for (int k = 1; k < PEER_MAX_SHARDS; ++k) r += (k < sh.n && id >= sh.start[k]) ? 1 : 0;
    // plain (coherent) loads: a peer's rows are not read-only for the lifetime of the kernel's module
    return *(sh.base[r] + (id - sh.start[r]) * vpr + c);
  };
  long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; t + 7 * stride < total; t += 8 * stride) {   // 8 loads in flight per thread: NVLink latency is ~2x HBM's
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = fetch(t + u * stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) out[t + u * stride] = v[u];
  }
  for (; t < total; t += stride) out[t] = fetch(t);
}

// --------------------------------------------------------------------------- normalise
// out[i,:] = x[i,:] / max(||x[i,:]||_2, eps)  (reference: F.normalize(weight[1:], dim=-1) and the user
// side, HSTU/main.py:180-184; run once per evaluation sweep instead of once per batch).  One warp per
// row, the row stays in registers between the norm and the scaling (one HBM read, one write);
// d % 8 == 0, d <= 1024.  inv_norm (nullable) receives 1/max(norm, eps).
template <typename TI, typename TO>
__global__ void normalize_rows_kernel(const TI* __restrict__ x, TO* __restrict__ out, float* __restrict__ inv_norm,
                                      long long n_rows, int d, float eps) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const TI* xr = x + row * d;
  float v[4][8];
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int k = c * 256 + lane * 8;
    if (k < d) {
      if constexpr (sizeof(TI) == 2) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(xr + k));
        v[c][0] = __uint_as_float(raw.x << 16); v[c][1] = __uint_as_float(raw.x & 0xFFFF0000u);
        v[c][2] = __uint_as_float(raw.y << 16); v[c][3] = __uint_as_float(raw.y & 0xFFFF0000u);
        v[c][4] = __uint_as_float(raw.z << 16); v[c][5] = __uint_as_float(raw.z & 0xFFFF0000u);
        v[c][6] = __uint_as_float(raw.w << 16); v[c][7] = __uint_as_float(raw.w & 0xFFFF0000u);
      } else {
        const float4 a = __ldg(reinterpret_cast<const float4*>(xr + k));
        const float4 b = __ldg(reinterpret_cast<const float4*>(xr + k + 4));
        v[c][0] = a.x; v[c][1] = a.y; v[c][2] = a.z; v[c][3] = a.w; v[c][4] = b.x; v[c][5] = b.y; v[c][6] = b.z; v[c][7] = b.w;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) ss = fmaf(v[c][e], v[c][e], ss);
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.f / fmaxf(sqrtf(ss), eps);
  if (inv_norm != nullptr && lane == 0) inv_norm[row] = inv;
  TO* orow = out + row * d;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int k = c * 256 + lane * 8;
    if (k < d) {
      if constexpr (sizeof(TO) == 2) {
        uint4 o;
        o.x = pack_bf16x2(v[c][0] * inv, v[c][1] * inv); o.y = pack_bf16x2(v[c][2] * inv, v[c][3] * inv);
        o.z = pack_bf16x2(v[c][4] * inv, v[c][5] * inv); o.w = pack_bf16x2(v[c][6] * inv, v[c][7] * inv);
        *reinterpret_cast<uint4*>(orow + k) = o;
      } else {
        *reinterpret_cast<float4*>(orow + k) = make_float4(v[c][0] * inv, v[c][1] * inv, v[c][2] * inv, v[c][3] * inv);
        *reinterpret_cast<float4*>(orow + k + 4) = make_float4(v[c][4] * inv, v[c][5] * inv, v[c][6] * inv, v[c][7] * inv);
      }
    }
  }
}

// -------------------------------------------------------------------------- gather-dot
// S[m,k] = scale * <U[m,:], table[idx[m,k],:]> without materialising the gathered (M,K,d) tensor
// (reference: itemEmbds[data[IUnseen]] + einsum("BD,BKD->BK"), SASRec/main.py:230-236; sampled softmax
// itemEmbds[cat(pos,negs)] + einsum("MD,MKD->MK"), HSTU/main.py:192-197; BPR/BCE logits, SASRec/main.py:203-206).
// One warp per query row; a group of GL lanes (GL = 16-byte vectors per row, rounded up to a power of
// two, <= 32) covers one table row, so 32/GL gathered rows are in flight per warp and every load is a
// full 16-byte vector.  Ids outside [0, n_rows) score 0.
template <typename T>
__device__ __forceinline__ void load_vec_as_float(const T* p, float (&f)[8]) {
  if constexpr (sizeof(T) == 2) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
    f[0] = __uint_as_float(raw.x << 16); f[1] = __uint_as_float(raw.x & 0xFFFF0000u);
    f[2] = __uint_as_float(raw.y << 16); f[3] = __uint_as_float(raw.y & 0xFFFF0000u);
    f[4] = __uint_as_float(raw.z << 16); f[5] = __uint_as_float(raw.z & 0xFFFF0000u);
    f[6] = __uint_as_float(raw.w << 16); f[7] = __uint_as_float(raw.w & 0xFFFF0000u);
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = f[5] = f[6] = f[7] = 0.f;
  }
}
template <typename T> struct VecElems { static constexpr int value = 16 / sizeof(T); };

template <typename T>
__global__ void gather_dot_kernel(const T* __restrict__ U, const T* __restrict__ table, const int64_t* __restrict__ idx,
                                  float scale, float* __restrict__ S, long long M, int K, long long n_rows, int d, int gl) {
  constexpr int VE = VecElems<T>::value;
  const long long m = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  const int vpr = d / VE;                 // 16-byte vectors per row
  const int sub = lane & (gl - 1), slot = lane / gl, nslot = 32 / gl;
  for (int k0 = 0; k0 < K; k0 += nslot) {
    const int k = k0 + slot;
    long long row = -1;
    if (k < K) row = __ldg(idx + m * K + k);
    const bool ok = row >= 0 && row < n_rows;
    float acc = 0.f;
    for (int c = sub; c < vpr; c += gl) {
      float u[8], w[8];
      load_vec_as_float<T>(U + m * d + c * VE, u);
      if (ok) load_vec_as_float<T>(table + row * d + c * VE, w);
      else { for (int e = 0; e < 8; ++e) w[e] = 0.f; }
#pragma unroll
      for (int e = 0; e < VE; ++e) acc = fmaf(u[e], w[e], acc);
    }
    for (int o = gl >> 1; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (sub == 0 && k < K) S[m * K + k] = acc * scale;
  }
}

// dU[m,:] = scale * sum_k G[m,k] * table[idx[m,k],:]   (k ascending inside a slot, slots combined in a fixed
// butterfly order => deterministic)
template <typename T>
__global__ void gather_dot_du_kernel(const T* __restrict__ table, const int64_t* __restrict__ idx, const float* __restrict__ G,
                                     float scale, float* __restrict__ dU, long long M, int K, long long n_rows, int d, int gl) {
  constexpr int VE = VecElems<T>::value;
  const long long m = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  const int vpr = d / VE;
  const int sub = lane & (gl - 1), slot = lane / gl, nslot = 32 / gl;
  for (int c0 = 0; c0 < vpr; c0 += gl) {   // warp-uniform trip count (the shuffles below need every lane)
    const int c = c0 + sub;
    const bool c_ok = c < vpr;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int k = slot; k < K; k += nslot) {
      const long long row = __ldg(idx + m * K + k);
      if (c_ok && row >= 0 && row < n_rows) {
        const float g = __ldg(G + m * K + k);
        float w[8];
        load_vec_as_float<T>(table + row * d + c * VE, w);
#pragma unroll
        for (int e = 0; e < VE; ++e) acc[e] = fmaf(g, w[e], acc[e]);
      }
    }
    for (int o = gl; o < 32; o <<= 1) {
#pragma unroll
      for (int e = 0; e < VE; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
    }
    if (slot == 0 && c_ok) {
      float* o = dU + m * d + c * VE;
#pragma unroll
      for (int e = 0; e < VE; e += 4)
        *reinterpret_cast<float4*>(o + e) = make_float4(acc[e] * scale, acc[e + 1] * scale, acc[e + 2] * scale, acc[e + 3] * scale);
    }
  }
}

// -------------------------------------------------------------------------------- SpMM
// Y = A X for CSR A (fp32 values) and dense fp32 X (cols, d); optionally acc += beta * Y in the same pass
// (reference: `allEmbds = self.Adj @ allEmbds; avgEmbds += allEmbds / (L+1)`, LightGCN/main.py:83-85).
// A group of GL = d/4 lanes (power of two, <= 32) owns one output row, 32/GL rows per warp; the row's
// non-zeros are walked in order (deterministic), four gathers in flight per lane.
__global__ void spmm_csr_kernel(const int64_t* __restrict__ crow, const int64_t* __restrict__ col,
                                const float* __restrict__ val, const float* __restrict__ X, float* __restrict__ Y,
                                float* __restrict__ acc_out, float beta, long long n_rows, long long n_cols, int d, int gl) {
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (gl - 1), slot = lane / gl, nslot = 32 / gl;
  const long long row = warp * nslot + slot;
  if (row >= n_rows) return;
  const long long e0 = crow[row], e1 = crow[row + 1];
  for (int c = sub * 4; c < d; c += gl * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    long long e = e0;
    for (; e + 4 <= e1; e += 4) {
      float4 x[4];
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long j = __ldg(col + e + u);
        v[u] = __ldg(val + e + u);
        x[u] = (j >= 0 && j < n_cols) ? __ldg(reinterpret_cast<const float4*>(X + j * d + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc.x = fmaf(v[u], x[u].x, acc.x); acc.y = fmaf(v[u], x[u].y, acc.y);
        acc.z = fmaf(v[u], x[u].z, acc.z); acc.w = fmaf(v[u], x[u].w, acc.w);
      }
    }
    for (; e < e1; ++e) {
      const long long j = __ldg(col + e);
      const float v = __ldg(val + e);
      if (j >= 0 && j < n_cols) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(X + j * d + c));
        acc.x = fmaf(v, x.x, acc.x); acc.y = fmaf(v, x.y, acc.y); acc.z = fmaf(v, x.z, acc.z); acc.w = fmaf(v, x.w, acc.w);
      }
    }
    if (Y != nullptr) *reinterpret_cast<float4*>(Y + row * d + c) = acc;
    if (acc_out != nullptr) {
      float4* o = reinterpret_cast<float4*>(acc_out + row * d + c);
      float4 t = *o;
      t.x = fmaf(beta, acc.x, t.x); t.y = fmaf(beta, acc.y, t.y); t.z = fmaf(beta, acc.z, t.z); t.w = fmaf(beta, acc.w, t.w);
      *o = t;
    }
  }
}

__device__ __forceinline__ float4 load4_as_float(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4_as_float(const __nv_bfloat16* p) {
  const uint2 raw = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xFFFF0000u),
                     __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xFFFF0000u));
}

// ------------------------------------------------ exact one-hot correction of dW (small M: no sort)
// dW[label_i] -= g*scale*u_i, dbias[label_i] -= g, summed over the query rows that share a label in index
// order (deterministic).  For the few thousand query rows of a training step a sort is overkill: the first
// query row with a given label owns it (atomicMin), and its warp scans the label vector once for the others.
//   first_of[l] = smallest i with label_i - base == l   (memset to 0x7F7F7F7F before)
__global__ void label_owner_kernel(const int64_t* __restrict__ labels, long long base, long long n_items, int m,
                                   int* __restrict__ first_of) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const long long l = labels[i] - base;
  if (l >= 0 && l < n_items) atomicMin(first_of + l, i);
}
// One warp per query row; only owners work.  TU: storage type of U.  OUT_BF16: the row's fp32 value comes from
// side[i] (written by the dW pass for rows with an owner) and is rounded into dW_bf16; otherwise dW (fp32) is
// updated in place.
// The owner's warp queues the indices of the rows that share its label (ascending) in shared memory and adds
// them in that order, 16 row loads in flight at a time: a Zipf-head label is shared by hundreds of query rows
// and one dependent L2 round trip per row made that single warp the whole kernel's duration (0.107 ms).
constexpr int LABEL_FIX_WARPS = 8;
constexpr int LABEL_FIX_BATCH = 16;
template <typename TU, bool OUT_BF16>
__global__ void __launch_bounds__(LABEL_FIX_WARPS * 32)
label_fix_kernel(const int64_t* __restrict__ labels, long long base, long long n_items, int m, int d,
                 const int* __restrict__ first_of, const TU* __restrict__ U, float gs,
                 float g_bias, const float* __restrict__ g_dev, const float* __restrict__ side,
                 float* __restrict__ dW, __nv_bfloat16* __restrict__ dW_bf16, float* __restrict__ dbias) {
  __shared__ int queue[LABEL_FIX_WARPS][64];
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= m) return;
  const long long mine = labels[i];
  const long long l = mine - base;
  if (l < 0 || l >= n_items || first_of[l] != i) return;
  int* q = queue[threadIdx.x >> 5];
  const float gd = (g_dev != nullptr) ? __ldg(g_dev) : 1.f;
  int count = 0;
  for (int c0 = 0; c0 < d; c0 += 128) {
    const int c = c0 + lane * 4;
    const bool col_ok = c < d;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    count = 0;
    int pend = 0;   // queued indices (warp-uniform), < 32 between iterations
    for (int jb = i & ~31; jb < m; jb += 32) {   // matches can only sit at or after the owner
      const int j = jb + lane;
      const bool is_hit = j >= i && j < m && labels[j] == mine;
      const uint32_t hit = __ballot_sync(0xffffffffu, is_hit);
      if (is_hit) q[pend + __popc(hit & ((1u << lane) - 1u))] = j;
      pend += __popc(hit);
      count += __popc(hit);
      const bool last = jb + 32 >= m;
      while (pend >= 32 || (last && pend > 0)) {   // acc += U[q[0 .. n)] in queue order
        const int n = min(pend, 32);
        __syncwarp();
        for (int k0 = 0; k0 < n; k0 += LABEL_FIX_BATCH) {
          float4 u[LABEL_FIX_BATCH];
#pragma unroll
          for (int e = 0; e < LABEL_FIX_BATCH; ++e)
            if (col_ok && k0 + e < n) u[e] = load4_as_float(U + static_cast<long long>(q[k0 + e]) * d + c);
#pragma unroll
          for (int e = 0; e < LABEL_FIX_BATCH; ++e)
            if (col_ok && k0 + e < n) { acc.x += u[e].x; acc.y += u[e].y; acc.z += u[e].z; acc.w += u[e].w; }
        }
        const int rest = (lane + 32 < pend) ? q[lane + 32] : 0;
        __syncwarp();
        if (lane + 32 < pend) q[lane] = rest;
        pend -= n;
        __syncwarp();
      }
    }
    __syncwarp();   // the queue is rewritten by the next column chunk
    if (col_ok) {
      const float al = -gs * gd;
      if constexpr (OUT_BF16) {
        const float4 v = *reinterpret_cast<const float4*>(side + static_cast<long long>(i) * d + c);
        uint2 pk;
        pk.x = pack_bf16x2(fmaf(al, acc.x, v.x), fmaf(al, acc.y, v.y));
        pk.y = pack_bf16x2(fmaf(al, acc.z, v.z), fmaf(al, acc.w, v.w));
        *reinterpret_cast<uint2*>(dW_bf16 + l * d + c) = pk;
      } else {
        float4* dst = reinterpret_cast<float4*>(dW + l * d + c);
        float4 o = *dst;
        o.x = fmaf(al, acc.x, o.x); o.y = fmaf(al, acc.y, o.y); o.z = fmaf(al, acc.z, o.z); o.w = fmaf(al, acc.w, o.w);
        *dst = o;
      }
    }
  }
  if (dbias != nullptr && lane == 0) dbias[l] += -g_bias * gd * static_cast<float>(count);
}

// fp32 (n) -> bf16
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);   // n % 8 == 0
    uint2 pk;
    pk.x = pack_bf16x2(v.x, v.y);
    pk.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(y + i) = pk;
  }
}

// ------------------------------------------------------------------- row compaction (a4)
// row_index[k] = position of the k-th set byte of mask (k < count), -1 beyond; *count = number of set bytes.
// The device-side replacement of `indices = positives != 0; userEmbds[indices]` (SASRec/main.py:199-200): torch's
// boolean indexing calls nonzero() and waits for the count on the host; here the count stays on the device and the
// fused CE entries take it as m_dev.  One CTA of 1024 threads walks the mask in chunks of 1024 with a running offset
// (n is the B x S positions of a batch: tens of thousands to a few hundred thousand).
__global__ void __launch_bounds__(1024) compact_index_kernel(const unsigned char* __restrict__ mask, long long n,
                                                             int64_t* __restrict__ row_index, int* __restrict__ count) {
  __shared__ int warp_sums[32];
  __shared__ int running;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) running = 0;
  __syncthreads();
  for (long long base = 0; base < n; base += 1024) {
    const long long i = base + tid;
    const int flag = (i < n && mask[i] != 0) ? 1 : 0;
    const unsigned int bal = __ballot_sync(0xffffffffu, flag);
    const int before = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_sums[wid] = __popc(bal);
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wid; ++w) woff += warp_sums[w];
    const int start = running;
    if (flag) row_index[start + woff + before] = i;
    __syncthreads();
    if (tid == 1023) running = start + woff + before + flag;   // the last thread's exclusive offset + its own flag = chunk total
    __syncthreads();
  }
  const int total = running;
  for (long long i = total + tid; i < n; i += 1024) row_index[i] = -1;
  if (tid == 0) *count = total;
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

// lse (natural log) -> -lse * log2(e), padded with -inf (=> exp2(x + aux) = 0 for padding rows)
__global__ void lse2_kernel(const float* __restrict__ lse, float* __restrict__ out, int m, int m_pad, const int* __restrict__ m_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (m_dev != nullptr) m = min(m, max(0, __ldg(m_dev)));   // rows beyond the device-side count do not exist
  if (i < m_pad) out[i] = (i < m) ? -lse[i] * 1.4426950408889634f : -INFINITY;
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

// bias (N, null = zeros) -> bias*log2(e), zero-padded to n_pad
__global__ void bias2_kernel(const float* __restrict__ bias, float* __restrict__ out, long long n, long long n_pad) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n_pad) out[i] = (i < n && bias != nullptr) ? bias[i] * 1.4426950408889634f : 0.f;
}


// ---- finish of the fused CE forward (pair kernel, PASS_FWD): one warp per query row merges the
// per-split partials (m2, l, A) and scores the label column exactly:
//   row_max = m*ln2, row_sumexp = sum_s l_s 2^(m_s-m), dU_unnorm = sum_s A_s 2^(m_s-m)  (= sum_j e^(S_ij-row_max) w_j)
//   label_logit = scale*<u_i, w_label> + bias[label]   (0 when the label is outside this shard)
template <typename T>
__global__ void ce_fwd_finish_kernel(const float* __restrict__ pm2, const float* __restrict__ pl,
                                     const float* __restrict__ pacc, int n_splits, long long slot_stride, int m, int d,
                                     const T* __restrict__ U, const T* __restrict__ W, const float* __restrict__ bias,
                                     const int64_t* __restrict__ labels, long long label_base, long long n_items,
                                     float scale, float* __restrict__ row_max, float* __restrict__ row_sumexp,
                                     float* __restrict__ label_logit, float* __restrict__ du_unnorm,
                                     const int* __restrict__ m_dev) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= m) return;
  if (m_dev != nullptr && row >= __ldg(m_dev)) {   // a row beyond the device-side count: neutral statistics (lse = 0, loss 0)
    if (lane == 0) { row_max[row] = 0.f; row_sumexp[row] = 1.f; label_logit[row] = 0.f; }
    if (du_unnorm != nullptr)
      for (int c = lane * 4; c < d; c += 128) *reinterpret_cast<float4*>(du_unnorm + static_cast<long long>(row) * d + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  float mx = -INFINITY;
  for (int s = 0; s < n_splits; ++s) mx = fmaxf(mx, pm2[s * slot_stride + row]);
  float l = 0.f;
  for (int s = 0; s < n_splits; ++s) l += pl[s * slot_stride + row] * exp2f(pm2[s * slot_stride + row] - mx);
  if (du_unnorm != nullptr) {
    for (int c = lane * 4; c < d; c += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < n_splits; ++s) {
        const float f = exp2f(pm2[s * slot_stride + row] - mx);
        const float4 v = *reinterpret_cast<const float4*>(pacc + (static_cast<long long>(s) * m + row) * d + c);
        acc.x = fmaf(f, v.x, acc.x); acc.y = fmaf(f, v.y, acc.y); acc.z = fmaf(f, v.z, acc.z); acc.w = fmaf(f, v.w, acc.w);
      }
      *reinterpret_cast<float4*>(du_unnorm + static_cast<long long>(row) * d + c) = acc;
    }
  }
  const long long lab = labels[row] - label_base;
  float ll = 0.f;
  if (lab >= 0 && lab < n_items) {
    float dot = 0.f;
    for (int c = lane * 4; c < d; c += 128) {
      const float4 u = load4_as_float(U + static_cast<long long>(row) * d + c);
      const float4 w = load4_as_float(W + lab * d + c);
      dot = fmaf(u.x, w.x, dot); dot = fmaf(u.y, w.y, dot); dot = fmaf(u.z, w.z, dot); dot = fmaf(u.w, w.w, dot);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    ll = dot * scale + (bias != nullptr ? __ldg(bias + lab) : 0.f);
  }
  if (lane == 0) {
    row_max[row] = mx * 0.6931471805599453f;
    row_sumexp[row] = l;
    label_logit[row] = ll;
  }
}

// dU (this shard's piece) = g*scale*( dU_unnorm * e^(row_max_local - lse) - [label in shard] w_label )
// with g = grad_scale * (*grad_scale_dev); summed over shards it is the exact CE gradient wrt U.
template <typename T>
__global__ void ce_du_finish_kernel(const float* __restrict__ du_unnorm, const float* __restrict__ row_max,
                                    const float* __restrict__ lse, const T* __restrict__ W,
                                    const int64_t* __restrict__ labels, long long label_base, long long n_items,
                                    float gs, const float* __restrict__ gs_dev, int m, int d, float* __restrict__ dU,
                                    const int* __restrict__ m_dev) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= m) return;
  if (m_dev != nullptr && row >= __ldg(m_dev)) {   // no such query row: zero gradient
    for (int c = lane * 4; c < d; c += 128) *reinterpret_cast<float4*>(dU + static_cast<long long>(row) * d + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float g = gs * (gs_dev != nullptr ? __ldg(gs_dev) : 1.f);
  const float f = expf(row_max[row] - lse[row]);
  const long long lab = labels[row] - label_base;
  const bool has = lab >= 0 && lab < n_items;
  for (int c = lane * 4; c < d; c += 128) {
    const float4 a = *reinterpret_cast<const float4*>(du_unnorm + static_cast<long long>(row) * d + c);
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has) w = load4_as_float(W + lab * d + c);
    float4 o;
    o.x = g * (a.x * f - w.x); o.y = g * (a.y * f - w.y); o.z = g * (a.z * f - w.z); o.w = g * (a.w * f - w.w);
    *reinterpret_cast<float4*>(dU + static_cast<long long>(row) * d + c) = o;
  }
}

// ------------------------------------------------------------------------ partial merges
__global__ void lse_merge_kernel(const float* __restrict__ pm2, const float* __restrict__ pl, const float* __restrict__ pll,
                                 int n_splits, long long slot_stride, int m, float* __restrict__ row_max,
                                 float* __restrict__ row_sumexp, float* __restrict__ label_logit,
                                 const int* __restrict__ m_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  if (m_dev != nullptr && i >= __ldg(m_dev)) { row_max[i] = 0.f; row_sumexp[i] = 1.f; label_logit[i] = 0.f; return; }
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

// Row-sharded table: the per-rank (row_max, row_sumexp, label_logit) triples, gathered as stats[R][3][M] (natural-log
// units), merged into the global lse and label logit of every query row in one launch (rb_rowstats_merge):
//   lse_i = mx + log(sum_r l_r exp(m_r - mx)),  ll_i = sum_r ll_r   (ranks in order => deterministic)
__global__ void rowstats_merge_kernel(const float* __restrict__ stats, int n_ranks, long long m, float* __restrict__ lse,
                                      float* __restrict__ label_logit) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= m) return;
  float mx = -INFINITY;
  for (int r = 0; r < n_ranks; ++r) mx = fmaxf(mx, stats[(static_cast<long long>(r) * 3 + 0) * m + i]);
  float l = 0.f, ll = 0.f;
  for (int r = 0; r < n_ranks; ++r) {
    const float mr = stats[(static_cast<long long>(r) * 3 + 0) * m + i];
    l += stats[(static_cast<long long>(r) * 3 + 1) * m + i] * expf(mr - mx);
    ll += stats[(static_cast<long long>(r) * 3 + 2) * m + i];
  }
  lse[i] = mx + logf(l);
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

// ------------------------------------------------------------------------- top-K machinery
// key order: score desc, then id asc (topk_key lives in sweep.cuh, shared with the sweep epilogues)

// Warp-wide bitonic network over 32*E values held E per lane (blocked: element index = lane*E + e).
// One compare-exchange stage (k = block size, j = partner distance); "up" blocks sort descending.
template <typename T, int E>
__device__ __forceinline__ void warp_bitonic_stage(T (&v)[E], int k, int j) {
  const uint32_t lane = lane_id();
  if (j >= E) {
    const int lj = j / E;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = lane * E + e;
      const T other = __shfl_xor_sync(0xffffffffu, v[e], lj);
      const bool up = ((i & k) == 0);
      const bool lower = ((i & j) == 0);
      const bool take_max = (up == lower);
      const T mx = v[e] > other ? v[e] : other;
      const T mn = v[e] > other ? other : v[e];
      v[e] = take_max ? mx : mn;
    }
  } else {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if ((e & j) == 0) {
        const int i = lane * E + e;
        const bool up = ((i & k) == 0);
        const T a = v[e], b = v[e + j];
        const T hi = a > b ? a : b, lo = a > b ? b : a;
        v[e] = up ? hi : lo;
        v[e + j] = up ? lo : hi;
      }
    }
  }
}
template <typename T, int E>
__device__ __forceinline__ void warp_bitonic_sort_desc(T (&v)[E]) {
#pragma unroll
  for (int k = 2; k <= 32 * E; k <<= 1)
#pragma unroll
    for (int j = k >> 1; j >= 1; j >>= 1) warp_bitonic_stage<T, E>(v, k, j);
}
// v is bitonic (as produced by max(best[i], new[n-1-i])): finish with the last merge network, descending
template <typename T, int E>
__device__ __forceinline__ void warp_bitonic_merge_desc(T (&v)[E]) {
#pragma unroll
  for (int j = 16 * E; j >= 1; j >>= 1) warp_bitonic_stage<T, E>(v, 64 * E /* bit never set => all "up" */, j);
}
// best <- the 32*E largest of (best U cur); both sorted descending on entry (cur is consumed)
template <typename T, int E>
__device__ __forceinline__ void warp_topk_absorb(T (&best)[E], T (&cur)[E]) {
  const uint32_t lane = lane_id();
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const T rev = __shfl_sync(0xffffffffu, cur[E - 1 - e], 31 - lane);
    best[e] = best[e] > rev ? best[e] : rev;
  }
  warp_bitonic_merge_desc<T, E>(best);
}
// element `idx` (0-based rank) of a blocked warp array
template <typename T, int E>
__device__ __forceinline__ T warp_blocked_get(const T (&v)[E], int idx) {
  T out = v[0];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const T c = __shfl_sync(0xffffffffu, v[e], idx / E);
    if (e == idx % E) out = c;
  }
  return out;
}

// ---- top-K pass 2a: per row, the K-th largest CLEAN tile maximum tau (NaN marks a dirty tile, one that
// holds a seen id: it never supports the threshold).  One warp per row; only values above the current
// threshold are staged (ballot compaction into shared memory) and the staged block is sorted and merged
// into the running top list when it fills up -- about 128 + K ln(n/128) insertions per row instead of a
// sort per 128-value chunk; the next chunk of tile maxima is in flight while the current one is examined.
// The ladder of the candidate sweep comes from the same sorted list: thr[0] = tau0 = K-th largest, checkpoint
// thr[k] = (K >> k)-th largest (k = 1..TOPK_LEVELS while K >> k >= 1, +inf beyond), each lowered by 2^-18 of its
// magnitude so that a top-K member whose tensor-core score differs from a supporter's in the last bits still
// passes; the level counters are zeroed.
template <int E>
__global__ void __launch_bounds__(128)
tilemax_select_kernel(const float* __restrict__ T, int n_tiles, long long n_rows, int K, RowLadder* __restrict__ ladder,
                      int* __restrict__ fb_count) {
  __shared__ float stage_s[4][32 * E];
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  if (row == 0 && lane == 0) *fb_count = 0;   // the fallback's slot counter starts every call at zero
  float* stage = stage_s[threadIdx.x >> 5];
  const float* t = T + row * n_tiles;
  const uint32_t lt = (1u << lane) - 1u;
  float best[E];
#pragma unroll
  for (int e = 0; e < E; ++e) best[e] = -INFINITY;
  float tau = -INFINITY;
  int ns = 0;
  auto flush = [&]() {
    __syncwarp();
    float cur[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = lane * E + e;
      cur[e] = (i < ns) ? stage[i] : -INFINITY;
    }
    warp_bitonic_sort_desc<float, E>(cur);
    warp_topk_absorb<float, E>(best, cur);
    tau = warp_blocked_get<float, E>(best, K - 1);
    ns = 0;
    __syncwarp();
  };
  float nxt[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = e * 32 + lane;
    nxt[e] = (i < n_tiles) ? __ldg(t + i) : -INFINITY;
  }
  for (int base = 0; base < n_tiles; base += 32 * E) {
    float cur[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      cur[e] = nxt[e];
      const int i = base + 32 * E + e * 32 + lane;
      nxt[e] = (i < n_tiles) ? __ldg(t + i) : -INFINITY;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const bool pass = cur[e] > tau;   // false for NaN (dirty tile) and for values that cannot raise the threshold
      const uint32_t m = __ballot_sync(0xffffffffu, pass);
      if (m != 0u) {
        if (pass) stage[ns + __popc(m & lt)] = cur[e];
        ns += __popc(m);
        if (ns > 32 * E - 32) flush();
      }
    }
  }
  if (ns > 0) flush();
  float mine = INFINITY;   // lane l holds thr[l]
  if (lane == 0) mine = tau;   // -inf when fewer than K clean tiles exist
#pragma unroll
  for (int k = 1; k <= TOPK_LEVELS; ++k) {
    const int rank = K >> k;   // warp-uniform
    if (rank >= 1) {
      const float v = warp_blocked_get<float, E>(best, rank - 1);
      if (lane == k) mine = v;
    }
  }
  if (mine > -INFINITY && mine < INFINITY) mine -= fabsf(mine) * 3.8146973e-06f;
  if (lane < 8) {
    ladder[row].thr[lane] = mine;
    ladder[row].cnt[lane] = 0u;
  }
}

// exact fp32 logit of one (query, item) pair: a sequential FMA chain over k = 0..d-1 (the same order in
// every top-K finishing path, so their values agree bit for bit), then scale and bias with explicit roundings
template <typename TW>
__device__ __forceinline__ float exact_logit(const float* __restrict__ u, const TW* __restrict__ w, int d, float scale,
                                             const float* __restrict__ bias, int item) {
  float acc = 0.f;
  constexpr int VE = 16 / sizeof(TW);   // elements per 16-byte vector
  constexpr int NV = 8;                 // vectors fetched ahead of the chain (bf16: 64 elements = one 128-byte line)
  for (int k0 = 0; k0 < d; k0 += NV * VE) {  // d % 8 == 0
    uint4 raw[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (k0 + v * VE < d) raw[v] = __ldg(reinterpret_cast<const uint4*>(w + k0 + v * VE));
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int k = k0 + v * VE;
      if (k < d) {
        if constexpr (sizeof(TW) == 2) {
          const float4 ua = *reinterpret_cast<const float4*>(u + k);
          const float4 ub = *reinterpret_cast<const float4*>(u + k + 4);
          acc = __fmaf_rn(ua.x, __uint_as_float(raw[v].x << 16), acc); acc = __fmaf_rn(ua.y, __uint_as_float(raw[v].x & 0xFFFF0000u), acc);
          acc = __fmaf_rn(ua.z, __uint_as_float(raw[v].y << 16), acc); acc = __fmaf_rn(ua.w, __uint_as_float(raw[v].y & 0xFFFF0000u), acc);
          acc = __fmaf_rn(ub.x, __uint_as_float(raw[v].z << 16), acc); acc = __fmaf_rn(ub.y, __uint_as_float(raw[v].z & 0xFFFF0000u), acc);
          acc = __fmaf_rn(ub.z, __uint_as_float(raw[v].w << 16), acc); acc = __fmaf_rn(ub.w, __uint_as_float(raw[v].w & 0xFFFF0000u), acc);
        } else {
          const float4 ua = *reinterpret_cast<const float4*>(u + k);
          acc = __fmaf_rn(ua.x, __uint_as_float(raw[v].x), acc); acc = __fmaf_rn(ua.y, __uint_as_float(raw[v].y), acc);
          acc = __fmaf_rn(ua.z, __uint_as_float(raw[v].z), acc); acc = __fmaf_rn(ua.w, __uint_as_float(raw[v].w), acc);
        }
      }
    }
  }
  if (bias != nullptr) return __fmaf_rn(acc, scale, __ldg(bias + item));
  return __fmul_rn(acc, scale);
}

// ---- top-K finish: the row's candidates are (group maximum, group id) pairs of aligned groups of 4 items that
// reached the row's running threshold in the EPI_CAND sweep (a few hundred, in n_sub sub-lists in global memory,
// L2-resident).  Two streaming trips over them, nothing table-sized in shared memory:
//   1. the cut T = K-th largest maximum among the CLEAN groups (each is the score of a distinct unseen item, so the
//      K-th best item scores >= T and every top-K member sits in a group whose maximum reaches T);
//   2. the groups that reach T (about K, plus the dirty ones) are re-scored exactly in fp32, eight groups = 32 items
//      per step as they come, seen items dropped (UniSRec/main.py:413), the K best by (score desc, id asc) kept.
// Only a sub-list that ran over its own capacity flags the row for the fallback below.  One warp per row.
#ifndef TFC_MIN_BLOCKS
#define TFC_MIN_BLOCKS 4
#endif
constexpr int TFC_MAXSUB = 512;    // sub-lists per row the flattened walk handles (prefix array in shared memory)
constexpr int TFC_KEEP = 512;      // candidates per row kept in shared memory between the two trips
constexpr int TFC_SEEN = 64;       // seen ids per row kept in shared memory for the filter
template <typename TW, int E>
__global__ void __launch_bounds__(128, (E <= 4 ? TFC_MIN_BLOCKS : 1))
topk_from_cands_kernel(const TW* __restrict__ U, const TW* __restrict__ W, const float* __restrict__ bias, float scale,
                       int d, long long n_rows, int n_items, const uint2* __restrict__ cand,
                       const int* __restrict__ cand_cnt, int n_sub, int cap, const int* __restrict__ seen_crow,
                       const int* __restrict__ seen_col, int K, int id_add, float* __restrict__ out_vals,
                       int* __restrict__ out_ids, int* __restrict__ overflow) {
  __shared__ __align__(16) float u_s[4][256];
  __shared__ unsigned long long stage_s[4][256];
  __shared__ float fstage_s[4][32 * E];
  __shared__ unsigned int queue_s[4][64];
  __shared__ int pref_s[4][TFC_MAXSUB + 1];
  __shared__ uint2 keep_s[4][TFC_KEEP];   // trip 1 parks the row's candidates here for trip 2 (when they fit)
  __shared__ int seen_s[4][TFC_SEEN];     // the row's seen ids (when they fit): the filter's binary search stays on chip
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 4 + wib;
  if (row >= n_rows) return;
  const int* cnts = cand_cnt + row * n_sub;
  const uint2* lists = cand + row * n_sub * cap;
  const uint32_t lt = (1u << lane) - 1u;
  // ---- a sub-list that overflowed its capacity lost candidates: the row goes to the exact scan
  bool ovf = false;
  for (int sb = lane; sb < n_sub; sb += 32) ovf |= cnts[sb] > cap;
  ovf = __any_sync(0xffffffffu, ovf);
  if (lane == 0) overflow[row] = ovf ? 1 : 0;
  if (ovf) return;
  // ---- the row's candidates as ONE sequence.  A small batch is swept by many splits (256 rows: 148 splits x 2 tile
  // parities = 296 sub-lists of two or three entries each), and walking them one by one costs two dependent L2 round
  // trips per sub-list (386 us for 256 rows).  So: exclusive prefix of the counts in shared memory, and lane l of
  // step s takes entry 32 s + l of the concatenation (binary search for its sub-list).
  int* pref = pref_s[wib];
  const bool flat = n_sub <= TFC_MAXSUB;
  int total = 0;
  if (flat) {
    for (int b0 = 0; b0 < n_sub; b0 += 32) {
      const int c = (b0 + lane < n_sub) ? cnts[b0 + lane] : 0;
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += up;
      }
      if (b0 + lane < n_sub) pref[b0 + lane] = total + inc - c;
      total += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) pref[n_sub] = total;
    __syncwarp();
  }
  uint2* keep = keep_s[wib];
  const bool parked = flat && total <= TFC_KEEP;   // trip 2 reads the candidates back from shared memory
  bool from_keep = false;
  // body(ok, entry) is called warp-uniformly once per 32 candidates, in list order
  auto walk = [&](auto&& body) {
    if (from_keep) {
      for (int e0 = 0; e0 < total; e0 += 32) {
        const int f = e0 + lane;
        body(f < total, (f < total) ? keep[f] : make_uint2(0u, 0u));
      }
    } else if (flat) {
      for (int e0 = 0; e0 < total; e0 += 32) {
        const int f = e0 + lane;
        uint2 x = make_uint2(0u, 0u);
        if (f < total) {
          int lo = 0, hi = n_sub;   // the last sub-list whose first entry is at or before f
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (pref[mid] <= f) lo = mid; else hi = mid;
          }
          x = __ldg(lists + static_cast<long long>(lo) * cap + (f - pref[lo]));
          if (parked) keep[f] = x;
        }
        body(f < total, x);
      }
    } else {
      for (int j = 0; j < n_sub; ++j) {
        const int c = cnts[j];
        const uint2* gl = lists + static_cast<long long>(j) * cap;
        for (int e0 = 0; e0 < c; e0 += 32) {
          const int e = e0 + lane;
          uint2 x = make_uint2(0u, 0u);
          if (e < c) x = __ldg(gl + e);
          body(e < c, x);
        }
      }
    }
  };
  // ---- trip 1: the cut = K-th largest maximum among the clean groups (-inf when there are fewer than K)
  float cut = -INFINITY;
  {
    float* fstage = fstage_s[wib];
    float bestf[E];
#pragma unroll
    for (int e = 0; e < E; ++e) bestf[e] = -INFINITY;
    float kthf = -INFINITY;
    int ns = 0;
    auto flushf = [&]() {
      __syncwarp();
      float cur[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int i = lane * E + e;
        cur[e] = (i < ns) ? fstage[i] : -INFINITY;
      }
      warp_bitonic_sort_desc<float, E>(cur);
      warp_topk_absorb<float, E>(bestf, cur);
      kthf = warp_blocked_get<float, E>(bestf, K - 1);
      ns = 0;
      __syncwarp();
    };
    walk([&](bool ok, uint2 x) {
      float v = -INFINITY;
      if (ok && !(x.y & CAND_DIRTY)) v = __uint_as_float(x.x);
      const bool pass = v > kthf;
      const uint32_t m = __ballot_sync(0xffffffffu, pass);
      if (m != 0u) {
        if (pass) fstage[ns + __popc(m & lt)] = v;
        ns += __popc(m);
        if (ns > 32 * E - 32) flushf();
      }
    });
    if (ns > 0) flushf();
    cut = kthf;
    if (cut > -INFINITY) cut -= fabsf(cut) * 3.8146973e-06f;   // tensor-core vs exact fp32 scores differ in the last bits
  }
  // ---- trip 2: exact scores of the groups that reach the cut
  float* u = u_s[wib];
  unsigned long long* stage = stage_s[wib];
  unsigned int* queue = queue_s[wib];
  for (int k = lane; k < 256; k += 32) u[k] = (k < d) ? static_cast<float>(U[row * d + k]) : 0.f;
  __syncwarp();
  int s_lo = 0, s_hi = 0;
  if (seen_crow != nullptr) { s_lo = seen_crow[row]; s_hi = seen_crow[row + 1]; }
  int* seen_l = seen_s[wib];
  const bool seen_local = (s_hi - s_lo) <= TFC_SEEN;
  if (seen_local)
    for (int k = lane; k < s_hi - s_lo; k += 32) seen_l[k] = __ldg(seen_col + s_lo + k);
  from_keep = parked;
  __syncwarp();
  unsigned long long best[E];
#pragma unroll
  for (int e = 0; e < E; ++e) best[e] = 0ull;
  unsigned long long kth = 0ull;
  int ns = 0;
  auto flush = [&]() {
    __syncwarp();
    for (int base = 0; base < ns; base += 32 * E) {
      unsigned long long cur[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int i = base + lane * E + e;
        cur[e] = (i < ns) ? stage[i] : 0ull;
      }
      warp_bitonic_sort_desc<unsigned long long, E>(cur);
      warp_topk_absorb<unsigned long long, E>(best, cur);
    }
    kth = warp_blocked_get<unsigned long long, E>(best, K - 1);
    ns = 0;
    __syncwarp();
  };
  // eight queued groups = 32 items, one per lane
  auto score_step = [&](int n_groups) {
    const int gi = lane >> 2;
    const int item = (gi < n_groups) ? (static_cast<int>(queue[gi]) << 2) + (lane & 3) : n_items;
    unsigned long long key = 0ull;
    if (item < n_items) key = topk_key(exact_logit<TW>(u, W + static_cast<long long>(item) * d, d, scale, bias, item), item);
    bool pass = key > kth;
    if (pass && s_hi > s_lo) {  // seen items never rank
      if (seen_local) {
        int lo = 0, hi = s_hi - s_lo;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (seen_l[mid] < item) lo = mid + 1; else hi = mid;
        }
        if (lo < s_hi - s_lo && seen_l[lo] == item) pass = false;
      } else {
        int lo = s_lo, hi = s_hi;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (__ldg(seen_col + mid) < item) lo = mid + 1; else hi = mid;
        }
        if (lo < s_hi && __ldg(seen_col + lo) == item) pass = false;
      }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, pass);
    if (pass) stage[ns + __popc(m & lt)] = key;
    ns += __popc(m);
    if (ns > 256 - 32) flush();
  };
  int nq = 0;   // groups waiting in the queue (< 8 between steps, up to 8 + 31 inside one)
  walk([&](bool ok, uint2 x) {
    const bool keep = ok && __uint_as_float(x.x) >= cut;
    const unsigned int gid = x.y & ~CAND_DIRTY;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (keep) queue[nq + __popc(m & lt)] = gid;
    nq += __popc(m);
    __syncwarp();
    while (nq >= 8) {
      score_step(8);
      __syncwarp();
      const unsigned int moved = (lane + 8 < nq) ? queue[lane + 8] : 0u;   // nq <= 39: one lane-wide shift suffices
      __syncwarp();
      if (lane + 8 < nq) queue[lane] = moved;
      nq -= 8;
      __syncwarp();
    }
  });
  if (nq > 0) score_step(nq);
  flush();
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

// ---- top-K fallback: an exact scan of a row's whole catalog (fp32 FMA over the stored operands), skipping seen ids,
// keeping the K best by (score desc, id asc), for the rows whose candidate lists overflowed (massive ties, catalogs too
// small for a threshold; essentially never on a large catalog with spread scores, but then it must not take 0.4 s per
// row either).  ONE cooperative launch that returns at once when no row is flagged:
//   phase 0  the flagged rows are compacted into a list (atomic slot counter)
//   phase 1  work items (flagged row, part of the catalog): one WARP scans FB_PARTS-th of the tiles (lane <-> item) and
//            leaves its sorted top-K keys in the workspace
//   phase 2  one warp per flagged row merges the parts' lists.
// More than FB_ROWS flagged rows go through several rounds of phases 1-2.
constexpr int FB_PARTS = 128;
constexpr int FB_ROWS = 64;
constexpr int FB_THREADS = 256;

template <typename TW, int E>
__global__ void __launch_bounds__(FB_THREADS)
topk_fallback_coop_kernel(const TW* __restrict__ U, const TW* __restrict__ W, const float* __restrict__ bias, float scale, int d,
                          long long n_rows, int n_items, const int* __restrict__ seen_crow, const int* __restrict__ seen_col,
                          int K, int id_add, float* __restrict__ out_vals, int* __restrict__ out_ids,
                          const int* __restrict__ overflow, int* __restrict__ fb_count, int* __restrict__ fb_list,
                          unsigned long long* __restrict__ fb_part) {
  __shared__ __align__(16) float u_s[FB_THREADS / 32][256];
  __shared__ unsigned long long cand_s[FB_THREADS / 32][256];
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gtid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long gthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const int gwarp = static_cast<int>(gtid >> 5), n_gwarps = static_cast<int>(gthreads >> 5);
  for (long long i = gtid; i < n_rows; i += gthreads)
    if (overflow[i] != 0) fb_list[atomicAdd(fb_count, 1)] = static_cast<int>(i);
  grid.sync();
  const int n_flag = *fb_count;
  if (n_flag == 0) return;   // grid-uniform
  float* u = u_s[wib];
  unsigned long long* cand = cand_s[wib];
  const uint32_t lt = (1u << lane) - 1u;
  const int n_tiles = (n_items + 127) / 128;
  const int KP = 32 * E;   // keys kept per part (>= K)
  for (int r0 = 0; r0 < n_flag; r0 += FB_ROWS) {
    const int n_round = min(FB_ROWS, n_flag - r0);
    // ---- phase 1
    for (int w = gwarp; w < n_round * FB_PARTS; w += n_gwarps) {
      const int ri = w / FB_PARTS, part = w % FB_PARTS;
      const long long row = fb_list[r0 + ri];
      const int tb = static_cast<int>(static_cast<long long>(part) * n_tiles / FB_PARTS);
      const int te = static_cast<int>(static_cast<long long>(part + 1) * n_tiles / FB_PARTS);
      __syncwarp();
      for (int k = lane; k < 256; k += 32) u[k] = (k < d) ? static_cast<float>(U[row * d + k]) : 0.f;
      __syncwarp();
      int s_lo = 0, s_hi = 0;
      if (seen_crow != nullptr) { s_lo = seen_crow[row]; s_hi = seen_crow[row + 1]; }
      unsigned long long best[E];
#pragma unroll
      for (int e = 0; e < E; ++e) best[e] = 0ull;
      unsigned long long kth = 0ull;
      int ncand = 0;
      auto flush = [&]() {
        for (int base = 0; base < ncand; base += 32 * E) {
          unsigned long long cur[E];
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int i = base + lane * E + e;
            cur[e] = (i < ncand) ? cand[i] : 0ull;
          }
          warp_bitonic_sort_desc<unsigned long long, E>(cur);
          warp_topk_absorb<unsigned long long, E>(best, cur);
        }
        kth = warp_blocked_get<unsigned long long, E>(best, K - 1);
        ncand = 0;
        __syncwarp();
      };
      for (int tile = tb; tile < te; ++tile) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        int item[4];
        const TW* wrow[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          item[q] = tile * 128 + q * 32 + lane;
          wrow[q] = W + static_cast<long long>(min(item[q], n_items - 1)) * d;
        }
        for (int k = 0; k < d; k += 8) {  // d % 8 == 0
          const float4 ua = *reinterpret_cast<const float4*>(u + k);
          const float4 ub = *reinterpret_cast<const float4*>(u + k + 4);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float w8[8];
            if constexpr (sizeof(TW) == 2) {
              const uint4 raw = __ldg(reinterpret_cast<const uint4*>(wrow[q] + k));
              w8[0] = __uint_as_float(raw.x << 16); w8[1] = __uint_as_float(raw.x & 0xFFFF0000u);
              w8[2] = __uint_as_float(raw.y << 16); w8[3] = __uint_as_float(raw.y & 0xFFFF0000u);
              w8[4] = __uint_as_float(raw.z << 16); w8[5] = __uint_as_float(raw.z & 0xFFFF0000u);
              w8[6] = __uint_as_float(raw.w << 16); w8[7] = __uint_as_float(raw.w & 0xFFFF0000u);
            } else {
              const float4 r0v = __ldg(reinterpret_cast<const float4*>(wrow[q] + k));
              const float4 r1v = __ldg(reinterpret_cast<const float4*>(wrow[q] + k + 4));
              w8[0] = r0v.x; w8[1] = r0v.y; w8[2] = r0v.z; w8[3] = r0v.w; w8[4] = r1v.x; w8[5] = r1v.y; w8[6] = r1v.z; w8[7] = r1v.w;
            }
            acc[q] = __fmaf_rn(ua.x, w8[0], acc[q]); acc[q] = __fmaf_rn(ua.y, w8[1], acc[q]);  // same chain as exact_logit
            acc[q] = __fmaf_rn(ua.z, w8[2], acc[q]); acc[q] = __fmaf_rn(ua.w, w8[3], acc[q]);
            acc[q] = __fmaf_rn(ub.x, w8[4], acc[q]); acc[q] = __fmaf_rn(ub.y, w8[5], acc[q]);
            acc[q] = __fmaf_rn(ub.z, w8[6], acc[q]); acc[q] = __fmaf_rn(ub.w, w8[7], acc[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool valid = item[q] < n_items;
          float sc = __fmul_rn(acc[q], scale);
          if (bias != nullptr && valid) sc = __fmaf_rn(acc[q], scale, __ldg(bias + item[q]));
          const unsigned long long key = valid ? topk_key(sc, item[q]) : 0ull;
          bool pass = key > kth;
          if (pass && s_hi > s_lo) {  // seen items never rank (UniSRec/main.py:413)
            int lo = s_lo, hi = s_hi;
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              if (__ldg(seen_col + mid) < item[q]) lo = mid + 1; else hi = mid;
            }
            if (lo < s_hi && __ldg(seen_col + lo) == item[q]) pass = false;
          }
          const uint32_t m = __ballot_sync(0xffffffffu, pass);
          if (pass) cand[ncand + __popc(m & lt)] = key;
          ncand += __popc(m);
          __syncwarp();
          if (ncand > 256 - 32) flush();
        }
      }
      flush();
      unsigned long long* dstp = fb_part + (static_cast<long long>(ri) * FB_PARTS + part) * KP;
#pragma unroll
      for (int e = 0; e < E; ++e) dstp[lane * E + e] = best[e];
    }
    grid.sync();
    // ---- phase 2
    for (int ri = gwarp; ri < n_round; ri += n_gwarps) {
      const long long row = fb_list[r0 + ri];
      unsigned long long best[E];
#pragma unroll
      for (int e = 0; e < E; ++e) best[e] = 0ull;
      for (int part = 0; part < FB_PARTS; ++part) {
        const unsigned long long* srcp = fb_part + (static_cast<long long>(ri) * FB_PARTS + part) * KP;
        unsigned long long cur[E];
#pragma unroll
        for (int e = 0; e < E; ++e) cur[e] = srcp[lane * E + e];   // already sorted descending, blocked layout
        warp_topk_absorb<unsigned long long, E>(best, cur);
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
    if (r0 + FB_ROWS < n_flag) grid.sync();   // the part lists are reused by the next round
  }
}

// ---- hit matrix of a ranked list: hits[b,k] = 1 if top_ids[b,k] is one of row b's targets (sorted CSR of
// data[IUnseen], UniSRec/main.py:414), else 0; missing entries (id < 0) never hit.  HR / NDCG / RECALL /
// PRECISION / MRR @k are prefix reductions of this (B,K) matrix (metrics.py).
__global__ void topk_hits_kernel(const int* __restrict__ top_ids, const int64_t* __restrict__ tcrow,
                                 const int64_t* __restrict__ tcol, long long n_rows, int K, float* __restrict__ hits) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_rows * K) return;
  const long long row = i / K;
  const long long id = top_ids[i];
  float h = 0.f;
  if (id >= 0) {
    long long lo = tcrow[row], hi = tcrow[row + 1];
    const long long end = hi;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (__ldg(tcol + mid) < id) lo = mid + 1; else hi = mid;
    }
    if (lo < end && __ldg(tcol + lo) == id) h = 1.f;
  }
  hits[i] = h;
}

// ---- every METRIC@k of a batch in one pass over the ranked ids (rb_topk_metrics): no (B,K) hit matrix travels, no
// per-metric launches.  One warp per row (lane <-> rank, 32 ranks at a time); lane m < n_metrics ends up with metric
// m's row value, summed over the warp's rows in double; blocks write partial sums, a second one-block launch adds
// them in block order and divides by the batch size => deterministic.  Metric definitions as metrics.metric_rows
// (HR = any hit, RECALL = hits / max(|targets|, 1), PRECISION = hits / k, NDCG = DCG / IDCG(min(k, |targets|)),
// MRR = 1 / (rank of the first hit + 1)); w = 1 / log2(rank + 2) and its running sum come from the host in float32.
constexpr int TM_MAX_METRICS = 32;
struct MetricSpec { int kind[TM_MAX_METRICS]; int k[TM_MAX_METRICS]; int n; };
__global__ void __launch_bounds__(256)
topk_metrics_kernel(const int* __restrict__ top_ids, const int64_t* __restrict__ tcrow, const int64_t* __restrict__ tcol,
                    long long n_rows, int K, const float* __restrict__ w, const float* __restrict__ w_cum,
                    const MetricSpec spec, double* __restrict__ partial) {
  __shared__ double blk[8][TM_MAX_METRICS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long warp0 = static_cast<long long>(blockIdx.x) * 8 + wib, n_warps = static_cast<long long>(gridDim.x) * 8;
  const int my_kind = lane < spec.n ? spec.kind[lane] : -1, my_k = lane < spec.n ? spec.k[lane] : 0;
  double acc = 0.0;
  for (long long row = warp0; row < n_rows; row += n_warps) {
    const long long t_lo = tcrow[row], t_hi = tcrow[row + 1];
    const int n_t = static_cast<int>(t_hi - t_lo);
    float cnt = 0.f, dcg = 0.f;      // lane m: hits within the first k_m ranks, and their discounted sum
    int first = 0x7fffffff;          // rank of the first hit (same in every lane)
    for (int r0 = 0; r0 < K; r0 += 32) {
      const int r = r0 + lane;
      bool hit = false;
      float wr = 0.f;
      if (r < K) {
        const long long id = top_ids[row * K + r];
        wr = __ldg(w + r);
        if (id >= 0) {
          long long lo = t_lo, hi = t_hi;
          while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (__ldg(tcol + mid) < id) lo = mid + 1; else hi = mid;
          }
          hit = lo < t_hi && __ldg(tcol + lo) == id;
        }
      }
      const uint32_t hm = __ballot_sync(0xffffffffu, hit);
      if (hm != 0u && first == 0x7fffffff) first = r0 + __ffs(hm) - 1;
      // lane m needs the hits of ranks < k_m: ranks r0 .. min(r0+32, k_m)
      const int upto = my_k - r0;   // how many ranks of this chunk count for lane m's metric
      const uint32_t mine = upto >= 32 ? hm : (upto > 0 ? (hm & ((1u << upto) - 1u)) : 0u);
      cnt += static_cast<float>(__popc(mine));
      uint32_t rest = hm;
      while (rest != 0u) {          // warp-uniform loop over the chunk's hits (LOU: at most one)
        const int j = __ffs(rest) - 1;
        rest &= rest - 1u;
        const float wj = __shfl_sync(0xffffffffu, wr, j);
        if ((mine >> j) & 1u) dcg += wj;
      }
    }
    float v = 0.f;
    if (my_kind == 0) v = cnt > 0.f ? 1.f : 0.f;
    else if (my_kind == 1) v = cnt / fmaxf(static_cast<float>(n_t), 1.f);
    else if (my_kind == 2) v = cnt / static_cast<float>(my_k);
    else if (my_kind == 3) { const int n_rel = min(n_t, my_k); v = n_rel > 0 ? dcg / __ldg(w_cum + n_rel - 1) : 0.f; }
    else if (my_kind == 4) v = first < my_k ? 1.f / (static_cast<float>(first) + 1.f) : 0.f;
    acc += static_cast<double>(v);
  }
  blk[wib][lane] = acc;
  __syncthreads();
  if (wib == 0 && lane < spec.n) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += blk[q][lane];
    partial[static_cast<long long>(blockIdx.x) * TM_MAX_METRICS + lane] = s;
  }
}
__global__ void topk_metrics_finish_kernel(const double* __restrict__ partial, int n_blocks, int n_metrics, long long n_rows,
                                           float* __restrict__ out) {
  const int m = threadIdx.x;
  if (m >= n_metrics) return;
  double s = 0.0;
  for (int b = 0; b < n_blocks; ++b) s += partial[static_cast<long long>(b) * TM_MAX_METRICS + m];
  out[m] = static_cast<float>(s / static_cast<double>(n_rows));
}

// ---- merge of R sorted per-shard lists (rb_topk_merge): list l of row i at (l*n_rows + i)*K + e
template <int E>
__global__ void topk_merge_kernel(const float* __restrict__ vals, const int* __restrict__ ids, int n_lists,
                                  long long n_rows, int K, float* __restrict__ out_vals, int* __restrict__ out_ids,
                                  long long list_stride) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  unsigned long long best[E];
#pragma unroll
  for (int e = 0; e < E; ++e) best[e] = 0ull;
  for (int l = 0; l < n_lists; ++l) {
    const long long ebase = static_cast<long long>(l) * list_stride + row * K;   // list_stride = n_rows*K, or the pitch of a packed gather
    for (int base = 0; base < K; base += 32 * E) {
      unsigned long long cur[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int i = base + lane * E + e;
        cur[e] = 0ull;
        if (i < K) {
          const int id = ids[ebase + i];
          if (id >= 0) cur[e] = topk_key(vals[ebase + i], id);
        }
      }
      warp_bitonic_sort_desc<unsigned long long, E>(cur);
      warp_topk_absorb<unsigned long long, E>(best, cur);
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
        id = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFu));
      }
      out_vals[row * K + i] = v;
      out_ids[row * K + i] = id;
    }
  }
}

}  // namespace rb
