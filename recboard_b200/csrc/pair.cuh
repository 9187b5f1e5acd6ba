// The training sweep engine ("pair kernel"): the two passes of the fused full-catalog cross-entropy
// (reference: einsum("MD,ND->MN") + F.cross_entropy and its autograd, SASRec/main.py:217-219,249).
//
// One CTA keeps TWO stationary 128-row tiles X0, X1 in shared memory (one per epilogue warpgroup) and
// streams 128-row tiles Y_j of the other operand through a TMA ring.  Per streamed tile and
// warpgroup g the tensor pipe runs
//     MMA1   S_g  = X_g . Y_j^T            (SMEM x SMEM -> TMEM, 128x128 fp32)
//     MMA2   A_g += P_g . Y_j              (TMEM x SMEM -> TMEM, 128 x d fp32)
// where P_g = 2^(c*S_g - ref) is produced by warpgroup g straight from TMEM registers, rounded to
// bf16 and written back over the first 64 columns of S_g (the A operand of MMA2 is read from TMEM:
// the softmax tile never touches shared or global memory).  Issue order
//     m2_0(j) m1_0(j+1) m2_1(j) m1_1(j+1)
// so that the exponentials of one warpgroup overlap the MMAs of the other.
//
//   PASS_FWD  rows stationary, items streamed: ref = lazily updated running row maximum;
//             outputs per (row, split): (m, l = sum P, A = sum_j P_ij w_j)  -> lse and dU
//   PASS_DW   items stationary, rows streamed: ref = lse of the streamed row (global softmax);
//             outputs A = sum_i P_ij u_i -> dW (and row sums -> dbias)
// The label one-hot never enters the tiles: dU subtracts w_label and dW subtracts u_i exactly, in
// fp32, in the finishing kernels (simt.cuh).
#pragma once
#include "ptx.cuh"

namespace rb {

enum : int { PASS_FWD = 0, PASS_DW = 1 };
constexpr int PAIR_THREADS = 320;
constexpr float PAIR_RESCALE_TH = 16.f;  // log2 units: P stays below 2^16 before the row reference moves

struct PairArgs {
  int n_stat;        // valid rows of the stationary operand
  int n_strm;        // valid rows of the streamed operand
  int n_pair_tiles;  // ceil(n_stat / 256)
  int n_strm_tiles;  // ceil(n_strm / 128)
  int n_splits;      // streamed range of every pair tile is cut into n_splits work items
  int d;             // true feature width
  int stat_pad;      // n_pair_tiles * 256 (row pitch of the per-row partial arrays)
  float scale;       // logits = scale * <u,w> + bias
  const float* bias2;  // per ITEM bias * log2(e) (nullable); FWD: streamed columns, DW: stationary rows
  // PASS_FWD outputs
  float* part_m2;    // [n_splits][stat_pad]  reference (log2 domain)
  float* part_l;     // [n_splits][stat_pad]  sum_j 2^(x_ij - m2)
  // PASS_DW inputs
  const float* lse2;         // per streamed query row: lse*log2(e), padded with +inf to a tile multiple
  float gscale;              // host part of the dW scale (g * scale)
  float rscale;              // host part of the dbias scale (g)
  const float* gscale_dev;   // optional device scalar multiplied in
  float* rowsum_out;         // [n_splits][n_stat] sum over streamed rows of P (dbias), nullable
  // both
  float* acc_out;    // [n_splits][n_stat][d]
};

template <int PASS_, int KC_, int NS_, bool BIAS_>
struct PairCfg {
  static constexpr int PASS = PASS_, KC = KC_, NS = NS_;
  static constexpr bool BIAS = BIAS_;
  static constexpr int DPAD = KC_ * 64;
  static constexpr int TILE_BYTES = KC_ * 128 * 128;  // one 128-row operand tile
  static constexpr int CTRL_BYTES = 1024;
  static constexpr int SMEM_BYTES = (2 + NS_) * TILE_BYTES + CTRL_BYTES + 1024 /*align*/;
  static constexpr int TMEM_COLS = 512;                // S0 | S1 | A0 | A1
  static constexpr int ACC0 = 256, ACC1 = 256 + DPAD;
  static_assert(KC_ == 1 || KC_ == 2, "d <= 128");
  static_assert(SMEM_BYTES <= 227 * 1024, "SMEM budget");
};

struct PairControl {
  uint64_t full[8], empty[8];
  uint64_t x_full, x_empty;
  uint64_t s_full[2], p_full[2];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(PairControl) <= 1024, "control block");

// acc[lane][col0 .. col0+ncols) *= f   (TMEM round trip; rare: only when a row reference moves)
__device__ __noinline__ void pair_rescale_acc(uint32_t t_acc, int ncols, float f) {
  for (int c = 0; c < ncols; c += 32) {
    uint32_t v[32];
    tmem_ld32(t_acc + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
    tmem_st32p(t_acc + c, v);
  }
  tmem_st_wait();
}

template <class C>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
pair_kernel(const __grid_constant__ CUtensorMap tm_stat, const __grid_constant__ CUtensorMap tm_strm,
            const PairArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* x_smem = smem;                         // X0 | X1
  uint8_t* y_smem = smem + 2 * C::TILE_BYTES;     // NS stages
  PairControl* bar = reinterpret_cast<PairControl*>(y_smem + C::NS * C::TILE_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_items = a.n_pair_tiles * a.n_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_stat);
    tma_prefetch_desc(&tm_strm);
    for (int i = 0; i < C::NS; ++i) { mbar_init(&bar->full[i], 1); mbar_init(&bar->empty[i], 1); }
    mbar_init(&bar->x_full, 1);
    mbar_init(&bar->x_empty, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&bar->s_full[g], 1);
      mbar_init(&bar->p_full[g], 128);
      mbar_init(&bar->acc_full[g], 1);
      mbar_init(&bar->acc_empty[g], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&bar->tmem_base, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bar->tmem_base;

  auto item_range = [&](int item, int& pair_tile, int& split, int& t0, int& t1) {
    pair_tile = item % a.n_pair_tiles;  // split-major: concurrent CTAs stream the same tiles (L2 reuse)
    split = item / a.n_pair_tiles;
    t0 = static_cast<int>((static_cast<long long>(split) * a.n_strm_tiles) / a.n_splits);
    t1 = static_cast<int>((static_cast<long long>(split + 1) * a.n_strm_tiles) / a.n_splits);
  };

  if (warp == 0) {
    // ======================================================================= TMA producer
    // The whole warp runs the (uniform) control flow; one elected lane issues the copies.
    uint32_t it = 0, k = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++k) {
      int pt, split, t0, t1;
      item_range(item, pt, split, t0, t1);
      mbar_wait(&bar->x_empty, (k & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bar->x_full, 2 * C::TILE_BYTES);
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int c = 0; c < C::KC; ++c)
            tma_load_2d(x_smem + g * C::TILE_BYTES + c * 16384, &tm_stat, &bar->x_full, c * 64, (pt * 2 + g) * 128);
      }
      __syncwarp();
      for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t st = it % C::NS, ph = (it / C::NS) & 1;
        mbar_wait(&bar->empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bar->full[st], C::TILE_BYTES);
#pragma unroll
          for (int c = 0; c < C::KC; ++c)
            tma_load_2d(y_smem + st * C::TILE_BYTES + c * 16384, &tm_strm, &bar->full[st], c * 64, t * 128);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ========================================================================= MMA issuer
    // Warp-converged loop (descriptor arithmetic stays in uniform registers); one elected lane issues
    // the tcgen05.mma / commit instructions.
    constexpr uint32_t idesc1 = make_idesc(FMT_BF16, 128, 128, 0, 0);
    constexpr uint32_t idesc2 = make_idesc(FMT_BF16, 128, C::DPAD, 0, 1);  // A: P from TMEM, B: Y tile MN-major
    constexpr uint32_t dhi = smem_desc_hi(1024);
    const uint32_t x_lo = smem_desc_lo(smem_u32(x_smem), 16);
    const uint32_t y_lo1 = smem_desc_lo(smem_u32(y_smem), 16);      // K-major view of a streamed tile (MMA1)
    const uint32_t y_lo2 = smem_desc_lo(smem_u32(y_smem), 16384);   // MN-major view of the same tile (MMA2)
    uint32_t it = 0, k = 0;

    // S_g = X_g . Y^T
    auto m1 = [&](int g, uint32_t st) {
      const uint32_t d_tmem = tmem_base + g * 128;
      const uint32_t xg = x_lo + ((g * C::TILE_BYTES) >> 4), ys = y_lo1 + ((st * C::TILE_BYTES) >> 4);
#pragma unroll
      for (int c = 0; c < C::KC; ++c) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          mma_f16_ss(d_tmem, smem_desc(dhi, xg + ((c * 16384 + kk * 32) >> 4)),
                     smem_desc(dhi, ys + ((c * 16384 + kk * 32) >> 4)), idesc1, (c | kk) != 0);
      }
      tc_commit(&bar->s_full[g]);
    };
    // A_g (+)= P_g . Y      (P_g: bf16 pairs in columns [0,64) of S_g)
    auto m2 = [&](int g, uint32_t st, bool first) {
      const uint32_t d_tmem = tmem_base + (g == 0 ? C::ACC0 : C::ACC1);
      const uint32_t a_tmem = tmem_base + g * 128;
      const uint32_t ys = y_lo2 + ((st * C::TILE_BYTES) >> 4);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        mma_f16_ts(d_tmem, a_tmem + kk * 8, smem_desc(dhi, ys + ((kk * 2048) >> 4)), idesc2, !(first && kk == 0));
    };

    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++k) {
      int pt, split, t0, t1;
      item_range(item, pt, split, t0, t1);
      const int n = t1 - t0;
      mbar_wait(&bar->x_full, k & 1);
      {
        const uint32_t st = it % C::NS, ph = (it / C::NS) & 1;
        mbar_wait(&bar->full[st], ph);
        tc_fence_after();
        if (elect_one()) {
          m1(0, st);
          m1(1, st);
          if (n == 1) tc_commit(&bar->x_empty);
        }
        __syncwarp();
      }
      for (int j = 0; j < n; ++j, ++it) {
        const uint32_t st = it % C::NS;
        const uint32_t st1 = (it + 1) % C::NS, ph1 = ((it + 1) / C::NS) & 1;
        const uint32_t pph = it & 1;
        if (j + 1 < n) mbar_wait(&bar->full[st1], ph1);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&bar->p_full[g], pph);
          if (j == 0) mbar_wait(&bar->acc_empty[g], (k & 1) ^ 1);
          tc_fence_after();
          if (elect_one()) {
            m2(g, st, j == 0);
            if (g == 1) tc_commit(&bar->empty[st]);          // all four MMAs on Y_j retire before this fires
            if (j + 1 == n) tc_commit(&bar->acc_full[g]);
            if (j + 1 < n) {
              m1(g, st1);
              if (g == 1 && j + 2 == n) tc_commit(&bar->x_empty);  // last use of X0/X1 in this item
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =========================================================================== epilogue
    const int g = (warp - 2) >> 2;     // warpgroup == stationary tile of the pair
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;       // row within the stationary tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t t_s = tmem_base + lane_base + g * 128;
    const uint32_t t_acc = tmem_base + lane_base + (g == 0 ? C::ACC0 : C::ACC1);
    const float c2 = a.scale * 1.4426950408889634f;   // scale > 0 (checked on the host)
    uint32_t it = 0, k = 0;

    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++k) {
      int pt, split, t0, t1;
      item_range(item, pt, split, t0, t1);
      const int srow = (pt * 2 + g) * 128 + r;   // global stationary row
      const bool srow_ok = srow < a.n_stat;

      float m2 = 0.f, l = 0.f;      // FWD: row reference (log2 domain) and sum of P
      float nb = 0.f;               // DW: bias2 of this item row
      float rowsum = 0.f;           // DW: sum_i P (dbias)
      if (C::PASS == PASS_DW && C::BIAS) nb = srow_ok ? __ldg(a.bias2 + srow) : 0.f;

      for (int t = t0; t < t1; ++t, ++it) {
        const int col_base = t * 128;
        mbar_wait(&bar->s_full[g], it & 1);
        tc_fence_after();
        uint32_t raw[128];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) tmem_ld32p(t_s + ch * 32, raw + ch * 32);
        tmem_ld_wait();

        if (C::PASS == PASS_FWD) {
          if (C::BIAS) {  // x = s*c2 + bias2[col] (bias2 is padded to a tile multiple by the host)
            const float4* b4 = reinterpret_cast<const float4*>(a.bias2 + col_base);
#pragma unroll
            for (int c4 = 0; c4 < 32; ++c4) {
              const float4 w = __ldg(b4 + c4);
              raw[c4 * 4 + 0] = __float_as_uint(fmaf(__uint_as_float(raw[c4 * 4 + 0]), c2, w.x));
              raw[c4 * 4 + 1] = __float_as_uint(fmaf(__uint_as_float(raw[c4 * 4 + 1]), c2, w.y));
              raw[c4 * 4 + 2] = __float_as_uint(fmaf(__uint_as_float(raw[c4 * 4 + 2]), c2, w.z));
              raw[c4 * 4 + 3] = __float_as_uint(fmaf(__uint_as_float(raw[c4 * 4 + 3]), c2, w.w));
            }
          }
          const int n_valid = a.n_strm - col_base;
          if (n_valid < 128) {  // last, partial tile: columns beyond the catalog never count
#pragma unroll
            for (int c = 0; c < 128; ++c)
              if (c >= n_valid) raw[c] = 0xff800000u;  // -inf
          }
          float mx[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) mx[i] = fmax3(__uint_as_float(raw[i * 32]), __uint_as_float(raw[i * 32 + 1]), __uint_as_float(raw[i * 32 + 2]));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int c = 3; c < 31; c += 2) mx[i] = fmax3(mx[i], __uint_as_float(raw[i * 32 + c]), __uint_as_float(raw[i * 32 + c + 1]));
            mx[i] = fmaxf(mx[i], __uint_as_float(raw[i * 32 + 31]));
          }
          float cm2 = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          if (!C::BIAS) cm2 *= c2;
          if (t == t0) {
            m2 = cm2;
          } else {
            const bool grow = cm2 > m2 + PAIR_RESCALE_TH;
            if (__any_sync(0xffffffffu, grow)) {
              // S_g(t) is complete => every earlier MMA, including m2_g(t-1), has retired: A_g is quiescent
              const float f = grow ? ex2_approx(m2 - cm2) : 1.f;
              pair_rescale_acc(t_acc, C::DPAD, f);
              l *= f;
              if (grow) m2 = cm2;
            }
          }
          const float nm = -m2;
          float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float x0 = __uint_as_float(raw[ch * 32 + 2 * i]), x1 = __uint_as_float(raw[ch * 32 + 2 * i + 1]);
              const float e0 = ex2_approx(C::BIAS ? x0 + nm : fmaf(x0, c2, nm));
              const float e1 = ex2_approx(C::BIAS ? x1 + nm : fmaf(x1, c2, nm));
              ls[i & 3] += e0 + e1;
              pk[i] = pack_bf16x2(e0, e1);
            }
            tmem_st16(t_s + ch * 16, pk);
          }
          l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
        } else {
          // P^T[item r][query row c] = 2^(s*c2 + bias2_r - lse2_c); lse2 = +inf beyond the last row => 0
          const float4* l4 = reinterpret_cast<const float4*>(a.lse2 + col_base);
          float rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint32_t pk[16];
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              const float4 w = __ldg(l4 + ch * 8 + c4);
              const float e0 = ex2_approx(fmaf(__uint_as_float(raw[ch * 32 + c4 * 4 + 0]), c2, nb - w.x));
              const float e1 = ex2_approx(fmaf(__uint_as_float(raw[ch * 32 + c4 * 4 + 1]), c2, nb - w.y));
              const float e2 = ex2_approx(fmaf(__uint_as_float(raw[ch * 32 + c4 * 4 + 2]), c2, nb - w.z));
              const float e3 = ex2_approx(fmaf(__uint_as_float(raw[ch * 32 + c4 * 4 + 3]), c2, nb - w.w));
              if (a.rowsum_out != nullptr) { rs[0] += e0; rs[1] += e1; rs[2] += e2; rs[3] += e3; }
              pk[c4 * 2 + 0] = pack_bf16x2(e0, e1);
              pk[c4 * 2 + 1] = pack_bf16x2(e2, e3);
            }
            tmem_st16(t_s + ch * 16, pk);
          }
          rowsum += (rs[0] + rs[1]) + (rs[2] + rs[3]);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bar->p_full[g]);
      }  // tiles

      // ---- per-item outputs: the accumulator row of this thread
      mbar_wait(&bar->acc_full[g], k & 1);
      tc_fence_after();
      float osc = 1.f;
      if (C::PASS == PASS_DW) osc = a.gscale * (a.gscale_dev != nullptr ? __ldg(a.gscale_dev) : 1.f);
      float* o = a.acc_out + (static_cast<long long>(split) * a.n_stat + srow) * a.d;
#pragma unroll 1
      for (int ch = 0; ch < C::DPAD / 32; ++ch) {
        uint32_t v[32];
        tmem_ld32(t_acc + ch * 32, v);
        tmem_ld_wait();
        if (srow_ok) {
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const int col = ch * 32 + c4 * 4;
            if (col < a.d) {  // d % 8 == 0 (checked on the host)
              float4 w;
              w.x = __uint_as_float(v[c4 * 4 + 0]) * osc;
              w.y = __uint_as_float(v[c4 * 4 + 1]) * osc;
              w.z = __uint_as_float(v[c4 * 4 + 2]) * osc;
              w.w = __uint_as_float(v[c4 * 4 + 3]) * osc;
              *reinterpret_cast<float4*>(o + col) = w;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&bar->acc_empty[g]);
      if (C::PASS == PASS_FWD) {
        const long long pslot = static_cast<long long>(split) * a.stat_pad + (pt * 2 + g) * 128 + r;
        a.part_m2[pslot] = m2;
        a.part_l[pslot] = l;
      } else if (a.rowsum_out != nullptr && srow_ok) {
        a.rowsum_out[static_cast<long long>(split) * a.n_stat + srow] =
            rowsum * a.rscale * (a.gscale_dev != nullptr ? __ldg(a.gscale_dev) : 1.f);
      }
    }  // items
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

}  // namespace rb
