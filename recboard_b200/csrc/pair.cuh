// The training sweep engine ("pair kernel"): the two passes of the fused full-catalog cross-entropy
// (reference: einsum("MD,ND->MN") + F.cross_entropy and its autograd, SASRec/main.py:217-219,249).
//
// One CTA keeps TWO stationary 128-row tiles X0, X1 in shared memory and streams 128-row tiles Y_t of the other
// operand through a TMA ring.  Per streamed tile and stationary tile g the tensor pipe runs
//     MMA1   S_g  = X_g . Y_t^T          (SMEM x SMEM -> TMEM, 128 x 128 fp32: eight N = 128 instructions at d = 128)
//     MMA2   A_g += P_g . Y_t            (TMEM x SMEM -> TMEM, 128 x d fp32, K = the 128 streamed rows)
// where P_g = 2^(c*S - ref) is produced straight from TMEM registers, rounded to bf16 and written back over S_g
// (the A operand of MMA2 is read from TMEM: the softmax tile never touches shared or global memory).
//
// TMEM map (512 columns): A_0 | A_1 (128 each, d of them used) | S_0 | S_1 (128 each): ONE score buffer per
// stationary tile, so the chain of a tile is serial -- MMA1(t) -> softmax(t) -> MMA2(t) -> MMA1(t+1) -- and the
// tensor pipe alternates between the two chains: while the epilogue turns S_0 into P_0 it runs MMA2_1 and the next
// MMA1_1 (1024 clk at d = 128).  That only pays if the softmax of a 128 x 128 tile takes well under those 1024 clk,
// which two warps per scheduler cannot do (round 1 measured such a design at 1.94 ms against 1.60 for 64-column
// MMA1s with double-buffered S, although the N = 64 instructions waste 25 % of the tensor pipe: 48 clk for half the
// work of a 64-clk N = 128 instruction, tests/probe_cta2.cu).  So the epilogue is FOUR warpgroups, two per
// stationary tile, each taking 64 of the tile's 128 score columns (thread <-> TMEM lane <-> stationary row; the two
// threads of a row sit in different warpgroups): four warps per scheduler hide each other's latencies and every
// thread has half the work.  576 threads: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..17 =
// epilogue warpgroups (g = stationary tile, ch = column half).
//
//   PASS_FWD  rows stationary, items streamed: ref = lazily updated running row maximum, agreed between the two
//             threads of a row through shared memory once per tile; outputs per (row, split):
//             (m, l = sum P, A = sum_j P_ij w_j)  -> lse and dU
//   PASS_DW   items stationary, rows streamed: ref = lse of the streamed row (global softmax);
//             outputs A = sum_i P_ij u_i -> dW (and row sums -> dbias)
// The label one-hot never enters the tiles: dU subtracts w_label and dW subtracts u_i exactly, in
// fp32, in the finishing kernels (simt.cuh).
// A per-streamed-row fp32 vector ("aux": bias*log2e of the streamed items in PASS_FWD, -lse*log2e of the
// streamed query rows in PASS_DW) rides along with every streamed tile as a 512-byte bulk copy.
#pragma once
#include "ptx.cuh"

namespace rb {

enum : int { PASS_FWD = 0, PASS_DW = 1 };
constexpr int PAIR_THREADS = 576;
constexpr float PAIR_RESCALE_TH = 16.f;  // log2 units: P stays below 2^16 before the row reference moves
// Every PAIR_POLY_EVERY-th element pair takes its exponentials from the FMA pipe (ex2_poly2) instead of
// the MUFU unit; 0 = all on MUFU.  With four epilogue warpgroups the passes are bound by the latency of the serial
// MMA1 -> softmax -> MMA2 chain, not by MUFU throughput: measured fwd+dU / dW 1.56 / 1.58 ms all-MUFU, 1.59 / 1.69 ms
// with every third pair on the FMA pipe, 1.63 / 1.72 ms with every second -- fewer instructions win.  (A build whose
// dW epilogue does no exponentials at all runs 1.60 ms: the softmax arithmetic is hidden; at 2.15 TFLOP executed per
// pass, 1.55 ms is the rate the measured sustained cuBLAS bf16 figure, 1.41 PFLOP/s under the 1 kW cap, allows.)
#ifndef PAIR_POLY_EVERY
#define PAIR_POLY_EVERY 0
#endif
__device__ __forceinline__ constexpr bool pair_use_poly(int i) { return PAIR_POLY_EVERY > 0 && (i % (PAIR_POLY_EVERY > 0 ? PAIR_POLY_EVERY : 1)) == PAIR_POLY_EVERY - 1; }

struct PairArgs {
  int n_stat;        // valid rows of the stationary operand
  int n_strm;        // valid rows of the streamed operand
  int n_pair_tiles;  // ceil(n_stat / 256)  (d-split variant: ceil(n_stat / 128))
  int n_strm_tiles;  // ceil(n_strm / 128)
  int n_splits;      // streamed range of every pair tile is cut into n_splits work items
  int d;             // true feature width
  int stat_pad;      // n_pair_tiles * 256 (d-split: * 128): row pitch of the per-row partial arrays
  float scale;       // logits = scale * <u,w> + bias
  // per STREAMED row, padded to n_strm_tiles*128: FWD = bias*log2(e) (BIAS only, 0 padding),
  //                                                DW  = -lse*log2(e) (-inf padding => P = 0)
  const float* aux;
  const float* bias2_stat;  // DW with BIAS: bias*log2(e) of the stationary item rows
  // PASS_FWD outputs
  float* part_m2;    // [n_splits][stat_pad]  reference (log2 domain)
  float* part_l;     // [n_splits][stat_pad]  sum_j 2^(x_ij - m2)
  // PASS_DW
  float gscale;              // host part of the dW scale (g * scale)
  float rscale;              // host part of the dbias scale (g)
  const float* gscale_dev;   // optional device scalar multiplied in
  float* rowsum_out;         // [n_splits][n_stat] sum over streamed rows of P (dbias), nullable
  // both
  float* acc_out;    // [n_splits][n_stat][d]
  // PASS_DW with a bf16 gradient (n_splits == 1): rows are stored as bf16 here instead of fp32 in acc_out;
  // rows that still get an exact fp32 one-hot correction (slot_of_row[row] < n_strm: the first query row with
  // that label) are ALSO stored in fp32 at side[slot][d], where the correction is applied before rounding.
  void* out_bf16;              // [n_stat][d] bf16, nullable
  const int* slot_of_row;      // [n_stat]
  float* side;                 // [n_slots][d]
  // Device-side row count (nullable): only the first *m_dev QUERY rows exist (stationary rows in PASS_FWD, streamed rows
  // in PASS_DW); tiles beyond them are skipped.  The host sizes buffers, tensor maps and the grid for the capacity.
  const int* m_dev;
  int accumulate;              // bf16 output only: rows are ADDED to what out_bf16 already holds (the parameter's existing gradient)
};

// DS_ ("d split", 128 < d <= 256): the accumulator of a 128-row stationary tile is 256 columns wide, so the CTA keeps
// ONE stationary tile X (128 rows x 4 K-chunks, the same 64 KB as two d = 128 tiles) and the two chains g = 0, 1
// work on the SAME rows: both compute S = X . Y_t^T over all of d (16 K-steps) and chain g accumulates output
// columns [128 g, 128 g + 128) from the g-th column half of the streamed tile.  S and its exponentials are computed
// twice (executed flop per pair: 2*(2d) + 2d against 2d + 2d), which keeps every hand-shake, the TMEM map and the
// epilogue of the d <= 128 kernel; both chains see bit-identical scores, so they take the same rescale decisions.
template <int PASS_, int KC_, int NS_, bool BIAS_, bool DS_ = false>
struct PairCfg {
  static constexpr int PASS = PASS_, KC = KC_, NS = NS_;
  static constexpr bool BIAS = BIAS_;
  static constexpr bool DS = DS_;
  static constexpr bool AUX = (PASS_ == PASS_DW) || BIAS_;   // a per-streamed-row vector rides with the tiles
  static constexpr int DPAD = KC_ * 64;                // accumulator columns of one chain
  static constexpr int XK = DS_ ? 4 : KC_;             // 64-element K chunks of an operand tile (MMA1 runs 4 XK K-steps)
  static constexpr int STAT_ROWS = DS_ ? 128 : 256;    // stationary rows per work item
  static constexpr int TILE_BYTES = XK * 128 * 128;    // one 128-row operand tile
  static constexpr int X_BYTES = DS_ ? TILE_BYTES : 2 * TILE_BYTES;
  static constexpr int AUX_BYTES = 512;
  static constexpr int CTRL_BYTES = 1024 + 4096;      // barriers + the row exchange between the column halves
  static constexpr int SMEM_BYTES = X_BYTES + NS_ * TILE_BYTES + NS_ * AUX_BYTES + CTRL_BYTES + 1024 /*align*/;
  static constexpr int TMEM_COLS = 512;
  static constexpr int ACC0 = 0, ACC1 = 128, S0 = 256;   // S_g at S0 + 128 g
  static_assert(KC_ == 1 || KC_ == 2, "128 accumulator columns per chain");
  static_assert(!DS_ || KC_ == 2, "the d-split variant accumulates two 128-column halves");
  static_assert(SMEM_BYTES <= 227 * 1024, "SMEM budget");
};

struct PairControl {
  uint64_t full[8], empty[8];
  uint64_t x_full, x_empty;
  uint64_t s_full[2];                    // per stationary tile
  uint64_t p_full[2][4];                 // per stationary tile and 32-column piece (column half * 2 + piece of the half)
  uint64_t unused_[2];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  uint32_t pad_[31];
  float xchg[2][2][2][128];              // [tile parity][g][column half][row]: hand-over between the two threads of a row
};
static_assert(sizeof(PairControl) <= 1024 + 4096, "control block");

// 1-D bulk copy global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// acc[lane][col0 .. col0+ncols) *= f   (TMEM round trip; rare: only when a row reference moves)
static __device__ __noinline__ void pair_rescale_acc(uint32_t t_acc, int ncols, float f) {
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

// 16-byte shared-memory load by shared-window address (the aux vectors: a pointer derived from the dynamic shared
// array through casts compiles to generic loads otherwise)
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

// named barrier over the 256 epilogue threads of stationary tile g (ids 1, 2; id 0 is __syncthreads)
__device__ __forceinline__ void pair_bar_sync(int g) { asm volatile("bar.sync %0, 256;" ::"r"(g + 1) : "memory"); }

// 576 threads leave 96 registers per thread (the register file is handed out as if the CTA had 20 warps); the
// epilogues below are written to fit without spilling.
template <class C>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
pair_kernel(const __grid_constant__ CUtensorMap tm_stat, const __grid_constant__ CUtensorMap tm_strm,
            const PairArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* x_smem = smem;                         // X0 | X1   (d-split: one tile of four K chunks)
  uint8_t* y_smem = smem + C::X_BYTES;            // NS stages
  float* aux_smem = reinterpret_cast<float*>(y_smem + C::NS * C::TILE_BYTES);   // NS x 128 floats
  PairControl* bar = reinterpret_cast<PairControl*>(y_smem + C::NS * C::TILE_BYTES + C::NS * C::AUX_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // effective extents: the host's, or what the device-side count of query rows leaves of them
  int n_stat = a.n_stat, n_pair_tiles = a.n_pair_tiles, n_strm_tiles = a.n_strm_tiles;
  if (a.m_dev != nullptr) {
    const int m = max(0, *a.m_dev);
    if (C::PASS == PASS_FWD) { n_stat = min(n_stat, m); n_pair_tiles = min(n_pair_tiles, (n_stat + C::STAT_ROWS - 1) / C::STAT_ROWS); }
    else n_strm_tiles = min(n_strm_tiles, (min(a.n_strm, m) + 127) / 128);
  }
  const int total_items = n_pair_tiles * a.n_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_stat);
    tma_prefetch_desc(&tm_strm);
    for (int i = 0; i < C::NS; ++i) { mbar_init(&bar->full[i], 1); mbar_init(&bar->empty[i], 1); }
    mbar_init(&bar->x_full, 1);
    mbar_init(&bar->x_empty, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&bar->s_full[g], 1);
      for (int pz = 0; pz < 4; ++pz) mbar_init(&bar->p_full[g][pz], 128);
      mbar_init(&bar->acc_full[g], 1);
      mbar_init(&bar->acc_empty[g], 256);
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
    pair_tile = item % n_pair_tiles;  // split-major: concurrent CTAs stream the same tiles (L2 reuse)
    split = item / n_pair_tiles;
    t0 = static_cast<int>((static_cast<long long>(split) * n_strm_tiles) / a.n_splits);
    t1 = static_cast<int>((static_cast<long long>(split + 1) * n_strm_tiles) / a.n_splits);
  };

  if (C::PASS == PASS_DW && n_strm_tiles == 0) {
    // no query rows at all (device-side count 0): the gradient of this pass is zero; nothing to stream
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int pt = item % n_pair_tiles, split = item / n_pair_tiles;
      for (int i = threadIdx.x; i < C::STAT_ROWS * (a.d / 4); i += blockDim.x) {
        const int row = pt * C::STAT_ROWS + i / (a.d / 4), c = (i % (a.d / 4)) * 4;
        if (row < a.n_stat) {
          if (a.out_bf16 != nullptr) {
            if (!a.accumulate) *reinterpret_cast<uint2*>(static_cast<uint16_t*>(a.out_bf16) + static_cast<long long>(row) * a.d + c) = make_uint2(0u, 0u);
          } else {
            *reinterpret_cast<float4*>(a.acc_out + (static_cast<long long>(split) * a.n_stat + row) * a.d + c) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (c == 0 && a.rowsum_out != nullptr) a.rowsum_out[static_cast<long long>(split) * a.n_stat + row] = 0.f;
        }
      }
    }
  } else if (warp == 0) {
    // ======================================================================= TMA producer
    // The whole warp runs the (uniform) control flow; one elected lane issues the copies.
    uint32_t it = 0, k = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++k) {
      int pt, split, t0, t1;
      item_range(item, pt, split, t0, t1);
      mbar_wait(&bar->x_empty, (k & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bar->x_full, C::X_BYTES);
        if (C::DS) {
#pragma unroll
          for (int c = 0; c < C::XK; ++c)
            tma_load_2d(x_smem + c * 16384, &tm_stat, &bar->x_full, c * 64, pt * 128);
        } else {
#pragma unroll
          for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int c = 0; c < C::KC; ++c)
              tma_load_2d(x_smem + g * C::TILE_BYTES + c * 16384, &tm_stat, &bar->x_full, c * 64, (pt * 2 + g) * 128);
        }
      }
      __syncwarp();
      for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t st = it % C::NS, ph = (it / C::NS) & 1;
        mbar_wait(&bar->empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bar->full[st], C::TILE_BYTES + (C::AUX ? C::AUX_BYTES : 0));
#pragma unroll
          for (int c = 0; c < C::XK; ++c)
            tma_load_2d(y_smem + st * C::TILE_BYTES + c * 16384, &tm_strm, &bar->full[st], c * 64, t * 128);
          if (C::AUX)
            bulk_load_1d(aux_smem + st * 128, a.aux + static_cast<long long>(t) * 128, C::AUX_BYTES, &bar->full[st]);
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

    // S_g = X_g . Y^T   (the streamed tile in stage st)
    auto m1 = [&](int g, uint32_t st) {
      const uint32_t d_tmem = tmem_base + C::S0 + g * 128;
      const uint32_t xg = x_lo + (C::DS ? 0u : ((g * C::TILE_BYTES) >> 4)), ys = y_lo1 + ((st * C::TILE_BYTES) >> 4);
#pragma unroll
      for (int c = 0; c < C::XK; ++c) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          mma_f16_ss(d_tmem, smem_desc(dhi, xg + ((c * 16384 + kk * 32) >> 4)),
                     smem_desc(dhi, ys + ((c * 16384 + kk * 32) >> 4)), idesc1, (c | kk) != 0);
      }
      tc_commit(&bar->s_full[g]);
    };
    // A_g (+)= P_g . Y for the 32 streamed rows of one piece (pz = column half * 2 + piece): two K = 16 steps.
    // P_g: bf16 pairs; streamed rows [0,64) sit in columns [0,32) of S_g, rows [64,128) in [64,96).
    auto m2_piece = [&](int g, int pz, uint32_t st, bool first) {
      const uint32_t d_tmem = tmem_base + (g == 0 ? C::ACC0 : C::ACC1);
      const uint32_t a_tmem = tmem_base + C::S0 + g * 128 + (pz >> 1) * 64 + (pz & 1) * 16;
      // d-split: chain g multiplies by the g-th column half (K chunks 2g, 2g+1) of the streamed tile
      const uint32_t ys = y_lo2 + ((st * C::TILE_BYTES + (C::DS ? g * 2 * 16384 : 0) + pz * 4096) >> 4);
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
        mma_f16_ts(d_tmem, a_tmem + kk * 8, smem_desc(dhi, ys + ((kk * 2048) >> 4)), idesc2, !(first && kk == 0));
    };

    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++k) {
      int pt, split, t0, t1;
      item_range(item, pt, split, t0, t1);
      const int n = t1 - t0;
      mbar_wait(&bar->x_full, k & 1);
      {  // the first streamed tile of the item: both score buffers are free (the MMA2s of the previous item were issued
         // before this point and the tensor pipe executes in order)
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
        const bool more = j + 1 < n;
        if (more) mbar_wait(&bar->full[st1], ph1);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (j == 0) mbar_wait(&bar->acc_empty[g], (k & 1) ^ 1);
          // the softmax pieces arrive in the order (half 0, piece 0), (half 1, piece 0), (half 0, piece 1), (half 1,
          // piece 1) -- the two column halves work side by side -- and MMA2 starts on the first while the last are
          // still being exponentiated
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const int pz = (o & 1) * 2 + (o >> 1);
            mbar_wait(&bar->p_full[g][pz], it & 1);
            tc_fence_after();
            if (elect_one()) m2_piece(g, pz, st, j == 0 && o == 0);
            __syncwarp();
          }
          if (elect_one()) {
            if (g == 1) tc_commit(&bar->empty[st]);   // every MMA on Y_j retires before this fires
            if (!more) tc_commit(&bar->acc_full[g]);
            if (more) {
              m1(g, st1);                              // in order behind MMA2_g: S_g / P_g is free again
              if (g == 1 && j + 2 == n) tc_commit(&bar->x_empty);  // last use of X0/X1 in this item
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =========================================================================== epilogue
    const int wgi = (warp - 2) >> 2;   // epilogue warpgroup 0..3
    const int g = wgi & 1;             // stationary tile of the pair
    const int ch = wgi >> 1;           // column half of the score tile this warpgroup handles
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;       // row within the stationary tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t t_s = tmem_base + lane_base + C::S0 + g * 128 + ch * 64;   // this thread's 64 score columns; P goes to the first 32
    constexpr int HALF = C::DPAD / 2;  // accumulator columns this thread rescales / writes out
    const uint32_t t_acc = tmem_base + lane_base + (g == 0 ? C::ACC0 : C::ACC1) + ch * HALF;
    const float c2 = a.scale * 1.4426950408889634f;   // scale > 0 (checked on the host)
    uint32_t it = 0, k = 0;

    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++k) {
      int pt, split, t0, t1;
      item_range(item, pt, split, t0, t1);
      const int srow = (C::DS ? pt : pt * 2 + g) * 128 + r;   // global stationary row
      const bool stats_owner = ch == 0 && (!C::DS || g == 0);   // the thread that reports the row's scalars
      const bool srow_ok = srow < n_stat;

      float m2 = 0.f, l = 0.f;      // FWD: row reference (log2 domain) and this thread's share of sum P
      float nb = 0.f;               // DW: bias2 of this item row
      float rowsum = 0.f;           // DW: this thread's share of sum_i P (dbias)
      if (C::PASS == PASS_DW && C::BIAS) nb = srow_ok ? __ldg(a.bias2_stat + srow) : 0.f;

      for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t st = it % C::NS;
        const uint32_t aux_s = smem_u32(aux_smem + st * 128 + ch * 64);   // this thread's 64 aux values
        // the aux vector arrived with the tile (same mbarrier); the stage cannot be refilled before this
        // warpgroup's p_full arrival of the tile, so the phase is stable while we look at it
        if (C::AUX) mbar_wait(&bar->full[st], (it / C::NS) & 1);
        const int col_base = t * 128 + ch * 64;
        mbar_wait(&bar->s_full[g], it & 1);
        tc_fence_after();
        if (C::PASS == PASS_FWD) {
          // Two trips through the thread's 64 score columns, 32 at a time: first the row maximum (the reference has to
          // be settled, and agreed with the row's other thread, before the first exponential), then the exponentials.
          // Holding all 64 values instead needs more than the 96 registers a 576-thread CTA leaves per thread, and a
          // spill here goes to L2 (200 KB of the SM's 256 KB are shared memory): measured 6.5 ms against 1.6.
          const int n_valid = a.n_strm - col_base;   // < 64 only in the last, partial tile: columns beyond the catalog never count
          auto load_piece = [&](int hf, uint32_t (&raw)[32]) {
            tmem_ld32p(t_s + hf * 32, raw);
            tmem_ld_wait();
            if (C::BIAS) {  // x = s*c2 + bias2[col]
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4) {
                const float4 w = lds128(aux_s + (hf * 8 + c4) * 16);
                raw[c4 * 4 + 0] = __float_as_uint(fmaf(__uint_as_float(raw[c4 * 4 + 0]), c2, w.x));
                raw[c4 * 4 + 1] = __float_as_uint(fmaf(__uint_as_float(raw[c4 * 4 + 1]), c2, w.y));
                raw[c4 * 4 + 2] = __float_as_uint(fmaf(__uint_as_float(raw[c4 * 4 + 2]), c2, w.z));
                raw[c4 * 4 + 3] = __float_as_uint(fmaf(__uint_as_float(raw[c4 * 4 + 3]), c2, w.w));
              }
            }
            if (n_valid - hf * 32 < 32) {
#pragma unroll
              for (int c = 0; c < 32; ++c)
                if (hf * 32 + c >= n_valid) raw[c] = 0xff800000u;  // -inf
            }
          };
          float cm2 = -INFINITY;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t raw[32];
            load_piece(hf, raw);
            float mx = fmax3(__uint_as_float(raw[0]), __uint_as_float(raw[1]), __uint_as_float(raw[2]));
#pragma unroll
            for (int c = 3; c < 31; c += 2) mx = fmax3(mx, __uint_as_float(raw[c]), __uint_as_float(raw[c + 1]));
            cm2 = fmax3(cm2, mx, __uint_as_float(raw[31]));
          }
          if (!C::BIAS) cm2 *= c2;
          // the two threads of a row (column halves, different warpgroups) agree on the tile maximum
          bar->xchg[it & 1][g][ch][r] = cm2;
          pair_bar_sync(g);
          cm2 = fmaxf(cm2, bar->xchg[it & 1][g][ch ^ 1][r]);
          if (t == t0) {
            m2 = cm2;
          } else {
            const bool grow = cm2 > m2 + PAIR_RESCALE_TH;
            if (__any_sync(0xffffffffu, grow)) {
              // A_g is quiescent here: MMA1 of this tile was issued behind MMA2 of the previous one and tcgen05.commit
              // completes in issue order, so s_full[g] (waited above) implies that MMA2 has retired; MMA2 of this tile
              // cannot start before our p_full arrivals.
              const float f = grow ? ex2_approx(m2 - cm2) : 1.f;
              pair_rescale_acc(t_acc, HALF, f);   // this thread's half of the row's accumulator
              l *= f;
              if (grow) m2 = cm2;
            }
          }
          const uint64_t nm2 = pack2(-m2, -m2), c22 = pack2(c2, c2);
          uint64_t ls2[2] = {0ull, 0ull};   // two packed running sums (4 fp32 lanes)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t raw[32];
            load_piece(hf, raw);
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
              uint32_t pk[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint64_t xr = pack2(__uint_as_float(raw[pc * 16 + 2 * i]), __uint_as_float(raw[pc * 16 + 2 * i + 1]));
                const uint64_t x2 = C::BIAS ? fadd2(xr, nm2) : ffma2(xr, c22, nm2);   // log2-domain argument
                float e0, e1;
                if (pair_use_poly(pc * 8 + i)) {
                  ex2_poly2(x2, e0, e1);
                } else {
                  float x0, x1;
                  unpack2(x2, x0, x1);
                  e0 = ex2_approx(x0);
                  e1 = ex2_approx(x1);
                }
                ls2[i & 1] = fadd2(ls2[i & 1], pack2(e0, e1));
                pk[i] = pack_bf16x2(e0, e1);
              }
              tmem_st8(t_s + hf * 16 + pc * 8, pk);
            }
            tmem_st_wait();   // a 32-column piece of P is complete: MMA2 may start on it
            tc_fence_before();
            mbar_arrive(&bar->p_full[g][ch * 2 + hf]);
          }
          {
            float s0, s1, s2, s3;
            unpack2(ls2[0], s0, s1);
            unpack2(ls2[1], s2, s3);
            l += (s0 + s1) + (s2 + s3);
          }
        } else {
          // P^T[item r][query row c] = 2^(s*c2 + bias2_r + aux_c), aux_c = -lse2_c (-inf beyond the last row => 0).
          // Element-wise, so the 64 columns go through the registers 32 at a time.
          const uint64_t c22 = pack2(c2, c2), nb2 = pack2(nb, nb);
          uint64_t rs2[2] = {0ull, 0ull};
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t raw[32];   // (both halves fetched up front: no gain at d = 128, 4 % slower at d = 64 -- measured)
            tmem_ld32p(t_s + hf * 32, raw);
            tmem_ld_wait();
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
              uint32_t pk[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 w = lds128(aux_s + (hf * 8 + pc * 4 + (i >> 1)) * 16);
                uint64_t off = (i & 1) ? pack2(w.z, w.w) : pack2(w.x, w.y);
                if (C::BIAS) off = fadd2(off, nb2);
                const uint64_t xr = pack2(__uint_as_float(raw[pc * 16 + 2 * i]), __uint_as_float(raw[pc * 16 + 2 * i + 1]));
                const uint64_t x2 = ffma2(xr, c22, off);
                float e0, e1;
                if (pair_use_poly(pc * 8 + i)) {
                  ex2_poly2(x2, e0, e1);
                } else {
                  float x0, x1;
                  unpack2(x2, x0, x1);
                  e0 = ex2_approx(x0);
                  e1 = ex2_approx(x1);
                }
                if (C::BIAS) rs2[i & 1] = fadd2(rs2[i & 1], pack2(e0, e1));   // row sums feed dbias only
                pk[i] = pack_bf16x2(e0, e1);
              }
              tmem_st8(t_s + hf * 16 + pc * 8, pk);
            }
            tmem_st_wait();   // a 32-column piece of P is complete: MMA2 may start on it
            tc_fence_before();
            mbar_arrive(&bar->p_full[g][ch * 2 + hf]);
          }
          if (C::BIAS) {
            float s0, s1, s2, s3;
            unpack2(rs2[0], s0, s1);
            unpack2(rs2[1], s2, s3);
            rowsum += (s0 + s1) + (s2 + s3);
          }
        }
      }  // tiles

      // ---- per-item outputs: the two threads of a row first combine their partial sums (the column-half-1 thread
      //      hands its share to the column-half-0 thread), then each writes its half of the accumulator row
      if (C::PASS == PASS_FWD || C::BIAS) {
        const uint32_t xb = (it & 1);   // the slot of the NEXT tile's parity: idle between two items
        if (ch == 1) bar->xchg[xb][g][1][r] = (C::PASS == PASS_FWD) ? l : rowsum;
        pair_bar_sync(g);
        if (ch == 0) {
          const float other = bar->xchg[xb][g][1][r];
          if (C::PASS == PASS_FWD) l += other; else rowsum += other;
        }
        pair_bar_sync(g);   // the slot is free again before the next item's first tile uses it
      }
      mbar_wait(&bar->acc_full[g], k & 1);
      tc_fence_after();
      float osc = 1.f;
      if (C::PASS == PASS_DW) osc = a.gscale * (a.gscale_dev != nullptr ? __ldg(a.gscale_dev) : 1.f);
      float* o = a.acc_out + (static_cast<long long>(split) * a.n_stat + srow) * a.d;
      const bool bf16_out = (C::PASS == PASS_DW) && a.out_bf16 != nullptr;   // kernel-uniform
      uint32_t* ob = nullptr;   // bf16 row as packed pairs
      if (bf16_out) {
        ob = reinterpret_cast<uint32_t*>(a.out_bf16) + (static_cast<long long>(srow) * a.d >> 1);
        const int slot = srow_ok ? __ldg(a.slot_of_row + srow) : 0x7F7F7F7F;   // owner query row of this label, if any
        o = (slot < a.n_strm) ? a.side + static_cast<long long>(slot) * a.d : nullptr;
      }
#pragma unroll 1
      for (int cc = 0; cc < HALF / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(t_acc + cc * 32, v);
        tmem_ld_wait();
        const int col0 = (C::DS ? g * 128 : 0) + ch * HALF + cc * 32;   // first output column of this chunk
        if (bf16_out) {
          if (srow_ok) {
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              const int col = col0 + c8 * 8;
              if (col < a.d) {  // d % 8 == 0 (checked on the host)
                float w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) w[e] = __uint_as_float(v[c8 * 8 + e]) * osc;
                if (a.accumulate) {   // kernel-uniform: the gradient buffer already holds a value (e.g. the gather's rows)
                  const uint4 old = *reinterpret_cast<const uint4*>(ob + (col >> 1));
                  w[0] += __uint_as_float(old.x << 16); w[1] += __uint_as_float(old.x & 0xFFFF0000u);
                  w[2] += __uint_as_float(old.y << 16); w[3] += __uint_as_float(old.y & 0xFFFF0000u);
                  w[4] += __uint_as_float(old.z << 16); w[5] += __uint_as_float(old.z & 0xFFFF0000u);
                  w[6] += __uint_as_float(old.w << 16); w[7] += __uint_as_float(old.w & 0xFFFF0000u);
                }
                uint4 pk;
                pk.x = pack_bf16x2(w[0], w[1]); pk.y = pack_bf16x2(w[2], w[3]);
                pk.z = pack_bf16x2(w[4], w[5]); pk.w = pack_bf16x2(w[6], w[7]);
                *reinterpret_cast<uint4*>(ob + (col >> 1)) = pk;
                if (o != nullptr) {
                  *reinterpret_cast<float4*>(o + col) = make_float4(w[0], w[1], w[2], w[3]);
                  *reinterpret_cast<float4*>(o + col + 4) = make_float4(w[4], w[5], w[6], w[7]);
                }
              }
            }
          }
        } else if (srow_ok) {
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const int col = col0 + c4 * 4;
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
      if (stats_owner) {
        if (C::PASS == PASS_FWD) {
          const long long pslot = static_cast<long long>(split) * a.stat_pad + (C::DS ? pt : pt * 2 + g) * 128 + r;
          a.part_m2[pslot] = m2;
          a.part_l[pslot] = l;
        } else if (a.rowsum_out != nullptr && srow_ok) {
          a.rowsum_out[static_cast<long long>(split) * a.n_stat + srow] =
              rowsum * a.rscale * (a.gscale_dev != nullptr ? __ldg(a.gscale_dev) : 1.f);
        }
      }
    }  // items
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

}  // namespace rb
