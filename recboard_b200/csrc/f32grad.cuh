// fp32-parity CE backward: exact fp32 (FFMA) recompute of the logits and both gradient contractions,
// for the configurations whose reference numbers are fp32 SGEMM (SASRec/main.py:217-219 + autograd at
// :249 on the Beauty / Yelp shapes, d = 64).  One kernel template serves both passes:
//
//     acc[r,:] = sum_c P[r,c] * Y[c,:],   P[r,c] = exp2( c2*<X_r, Y_c> + stat_add[r] + strm_add[c] )
//
//   dU pass: X = U (query rows stationary), Y = W (items streamed), stat_add = -lse2, strm_add = bias2
//   dW pass: X = W (items stationary),      Y = U (rows streamed),  stat_add = bias2, strm_add = -lse2
//            (+ row sums of P -> dbias)
// The label one-hot never enters the tiles (same as the tensor-core passes): the finishing kernels in
// simt.cuh subtract w_label / u_i exactly.  A 64 x 64 tile of P lives in shared memory only.
// Work item = (64-row stationary tile, split of the streamed range); split partials are summed in a
// fixed order afterwards => deterministic.
#pragma once
#include "ptx.cuh"

namespace rb {

constexpr int F32G_TILE = 64;
constexpr int F32G_THREADS = 256;

template <int DP>
struct F32GradCfg {
  static constexpr int LDX = DP + 4;          // padded operand row pitch (floats): conflict-free 128-bit reads
  static constexpr int LDP = F32G_TILE + 4;   // padded P row pitch
  static constexpr int SMEM_BYTES = (2 * F32G_TILE * LDX + F32G_TILE * LDP) * 4;
};

struct F32GradArgs {
  const float* X;         // stationary operand (n_stat, d)
  const float* Y;         // streamed operand   (n_strm, d)
  int n_stat, n_strm, d;
  int n_splits;
  float c2;               // scale * log2(e)
  const float* stat_vec;  // per stationary row, natural units (nullable)
  float stat_mul;         // +1 (bias) or -1 (lse)
  const float* strm_vec;  // per streamed row (nullable)
  float strm_mul;
  float out_scale;            // acc is stored as acc * out_scale * (*scale_dev)
  float rowsum_scale;         // row sums are stored as sum * rowsum_scale * (*scale_dev)
  const float* scale_dev;     // optional device scalar
  float* acc_out;             // [n_splits][n_stat][d]
  float* rowsum_out;          // [n_splits][n_stat], nullable
  // device-side count of QUERY rows (nullable) and which side they are on: 0 = stationary (dU pass), 1 = streamed (dW pass)
  const int* m_dev;
  int m_side;
};

template <int DP>
__global__ void __launch_bounds__(F32G_THREADS, 2) ce_grad_f32_kernel(const F32GradArgs a) {
  using C = F32GradCfg<DP>;
  extern __shared__ __align__(16) float f32g_smem[];
  float* Xs = f32g_smem;
  float* Ys = Xs + F32G_TILE * C::LDX;
  float* Ps = Ys + F32G_TILE * C::LDX;

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n_tiles = (a.n_stat + F32G_TILE - 1) / F32G_TILE;   // the grid and the output pitch follow the host's capacity
  int n_strm_eff = a.n_strm, n_stat_eff = a.n_stat;
  if (a.m_dev != nullptr) {
    const int m = max(0, *a.m_dev);
    if (a.m_side == 0) n_stat_eff = min(n_stat_eff, m); else n_strm_eff = min(n_strm_eff, m);
  }
  const int n_strm_tiles = (n_strm_eff + F32G_TILE - 1) / F32G_TILE;
  const int tile = blockIdx.x % n_tiles, split = blockIdx.x / n_tiles;
  const int t0 = static_cast<int>((static_cast<long long>(split) * n_strm_tiles) / a.n_splits);
  const int row0 = tile * F32G_TILE;
  // a stationary tile beyond the device-side row count streams nothing and stores zeros
  const int t1 = (row0 >= n_stat_eff) ? t0 : static_cast<int>((static_cast<long long>(split + 1) * n_strm_tiles) / a.n_splits);
  constexpr int VPR = DP / 4;   // float4 per padded row
  constexpr int NG = DP / 64;   // float4 column groups per thread in the second contraction

  // stationary tile -> shared memory (rows beyond n_stat and columns beyond d are zero)
  for (int v = tid; v < F32G_TILE * VPR; v += F32G_THREADS) {
    const int r = v / VPR, c = (v - r * VPR) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < a.n_stat && c < a.d) x = __ldg(reinterpret_cast<const float4*>(a.X + static_cast<long long>(row0 + r) * a.d + c));
    *reinterpret_cast<float4*>(Xs + r * C::LDX + c) = x;
  }
  float sa[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    sa[i] = (a.stat_vec != nullptr && r < a.n_stat) ? a.stat_mul * 1.4426950408889634f * __ldg(a.stat_vec + r) : 0.f;
  }

  // Packed fp32 arithmetic (fma.rn.f32x2): a three-register FFMA issues every second cycle per scheduler on this
  // architecture, the packed form does two FMAs in the same slot.  Accumulators are register PAIRS: output columns
  // (c, c+1) here, the even-k / odd-k partial sums of a score below.
  uint64_t acc2[4][NG][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int e = 0; e < 2; ++e) acc2[i][g][e] = 0ull;
  float rs[4] = {0.f, 0.f, 0.f, 0.f};

  for (int t = t0; t < t1; ++t) {
    const int col0 = t * F32G_TILE;
    __syncthreads();  // previous tile's Ys / Ps reads are done (and Xs is visible on the first trip)
    for (int v = tid; v < F32G_TILE * VPR; v += F32G_THREADS) {
      const int r = v / VPR, c = (v - r * VPR) * 4;
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col0 + r < n_strm_eff && c < a.d) y = __ldg(reinterpret_cast<const float4*>(a.Y + static_cast<long long>(col0 + r) * a.d + c));
      *reinterpret_cast<float4*>(Ys + r * C::LDX + c) = y;
    }
    float sb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + tx + 16 * j;
      sb[j] = (c < n_strm_eff) ? ((a.strm_vec != nullptr) ? a.strm_mul * 1.4426950408889634f * __ldg(a.strm_vec + c) : 0.f) : -INFINITY;
    }
    __syncthreads();

    // ---- S = X Y^T : rows ty*4+i, columns tx+16j
    uint64_t s2[4][4];   // (sum over even k, sum over odd k)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s2[i][j] = 0ull;
#pragma unroll 4
    for (int k = 0; k < DP; k += 4) {
      ulonglong2 xv[4], yv[4];   // four consecutive k as two packed pairs: operands come paired straight from the loads
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const ulonglong2*>(Xs + (ty * 4 + i) * C::LDX + k);
#pragma unroll
      for (int j = 0; j < 4; ++j) yv[j] = *reinterpret_cast<const ulonglong2*>(Ys + (tx + 16 * j) * C::LDX + k);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s2[i][j] = ffma2(xv[i].x, yv[j].x, s2[i][j]);
          s2[i][j] = ffma2(xv[i].y, yv[j].y, s2[i][j]);
        }
    }
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float e, o;
        unpack2(s2[i][j], e, o);
        s[i][j] = e + o;
      }
    // ---- P tile
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = ex2_approx(fmaf(s[i][j], a.c2, sa[i] + sb[j]));
        rs[i] += p;
        Ps[(ty * 4 + i) * C::LDP + tx + 16 * j] = p;
      }
    __syncthreads();

    // ---- acc += P Y : rows ty*4+i, columns tx*4 + 64g
#pragma unroll 2
    for (int jj = 0; jj < F32G_TILE; jj += 4) {
      float4 pv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = *reinterpret_cast<const float4*>(Ps + (ty * 4 + i) * C::LDP + jj);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint64_t pp[4];   // (p, p) per row
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float p = (u == 0) ? pv[i].x : (u == 1) ? pv[i].y : (u == 2) ? pv[i].z : pv[i].w;
          pp[i] = pack2(p, p);
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const ulonglong2 yv = *reinterpret_cast<const ulonglong2*>(Ys + (jj + u) * C::LDX + tx * 4 + 64 * g);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc2[i][g][0] = ffma2(pp[i], yv.x, acc2[i][g][0]);
            acc2[i][g][1] = ffma2(pp[i], yv.y, acc2[i][g][1]);
          }
        }
      }
    }
  }

  const float sdev = (a.scale_dev != nullptr) ? __ldg(a.scale_dev) : 1.f;
  const float osc = a.out_scale * sdev;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    if (r < a.n_stat) {
      float* o = a.acc_out + (static_cast<long long>(split) * a.n_stat + r) * a.d;
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const int c = tx * 4 + 64 * g;
        float v0, v1, v2, v3;
        unpack2(acc2[i][g][0], v0, v1);
        unpack2(acc2[i][g][1], v2, v3);
        if (c < a.d) *reinterpret_cast<float4*>(o + c) = make_float4(v0 * osc, v1 * osc, v2 * osc, v3 * osc);
      }
    }
  }
  if (a.rowsum_out != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v = rs[i];
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);  // the 16 tx lanes of a row
      const int r = row0 + ty * 4 + i;
      if (tx == 0 && r < a.n_stat) a.rowsum_out[static_cast<long long>(split) * a.n_stat + r] = v * a.rowsum_scale * sdev;
    }
  }
}

}  // namespace rb
