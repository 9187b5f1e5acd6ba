// The scoring sweep engine: one warp-specialised tcgen05 kernel template that computes tiles of
//     S = X_stat (128 rows) x Y_strm^T (BN rows)      (both operands K-major, fed by TMA)
// for a stationary 128-row tile against a range of streamed tiles, with four fused epilogues:
//
//   EPI_DENSE  store S (compat path for recommend_from_full, reference SASRec/main.py:228)
//   EPI_LSE    online (max, sum-exp) + label-logit pick   (F.cross_entropy fwd, SASRec/main.py:217-219)
//   EPI_GRAD   P = exp2(S*c - lse2) -> bf16 tile G in SMEM -> second MMA  Acc += G x Y_strm
//              (autograd of SASRec/main.py:217-219: dU = P.W with rows stationary,
//               dW = P^T.U with items stationary) -- the (M,N) matrix is never written
//   EPI_TOPK   masked maximum of every (row, 128-item tile): pass 1 of the exact top-K
//              (UniSRec/main.py:408-435 without dense (B,N); simt.cuh finishes the selection)
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2..9 = two epilogue warpgroups (thread <-> TMEM lane <-> stationary row); warpgroup g owns
// the S/G buffer g, i.e. the tiles of parity g, so two tiles are always in flight per SM.
#pragma once
#include "ptx.cuh"

namespace rb {

enum : int { EPI_DENSE = 0, EPI_LSE = 1, EPI_GRAD = 2, EPI_TOPK = 3 };
enum : int { DT_BF16 = 0, DT_TF32X3 = 1 };

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float MASKED_SCORE = -1e23f;  // UniSRec/main.py:413
constexpr int SWEEP_THREADS = 320;

struct SweepArgs {
  int n_stat;        // valid rows of the stationary operand
  int n_strm;        // valid rows of the streamed operand
  int n_stat_tiles;  // ceil(n_stat / 128)
  int n_strm_tiles;  // ceil(n_strm / BN)
  int n_splits;      // the streamed range of every stationary tile is cut into n_splits work items
  int d;             // true feature width (<= KC*64)
  float scale;       // logits = scale * <u,w> + bias
  const float* bias; // per ITEM bias (nullable)
  // EPI_DENSE (rows stationary)
  float* out;
  long long ld_out;
  // EPI_LSE (rows stationary)
  const int* labels;  // per query row: local item index, or -1
  float* part_m2;     // [n_splits][n_stat_tiles*128] running max, log2 domain
  float* part_l;      //   "  sum of 2^(x - m2)
  float* part_ll;     //   "  label logit (natural units) or 0
  // EPI_GRAD
  const float* lse2;  // per query row lse*log2(e), padded with +inf to a tile multiple
  float gscale;       // upstream grad / M
  const float* gscale_dev;  // optional device scalar multiplied into gscale (autograd's grad_output)
  float* acc_out;     // [n_splits][n_stat][d]
  float* rowsum_out;  // [n_splits][n_stat]  sum_j P (items stationary: dbias), nullable
  // EPI_TOPK (rows stationary): pass 1 of the top-K = masked maximum of every (row, 128-item tile)
  const int* seen_crow;  // [n_stat+1] CSR of already-seen LOCAL item ids, sorted per row (nullable)
  const int* seen_col;
  float* tile_max;    // [n_stat][n_strm_tiles]
};

template <int EPI_, int DT_, int KC_, int BN_, int NS_, bool STAT_ROWS_>
struct SweepCfg {
  static constexpr int EPI = EPI_, DT = DT_, KC = KC_, BN = BN_, NS = NS_;
  static constexpr bool STAT_ROWS = STAT_ROWS_;  // true: queries stationary, items streamed
  // storage chunks (128-byte columns groups) per operand row
  static constexpr int KCS = (DT_ == DT_BF16) ? KC_ : 2 * KC_;  // tf32x3: [hi | lo], KC = d/32
  static constexpr int NPAIR = (DT_ == DT_BF16) ? KC_ : 3 * KC_;
  static constexpr int ELEMS_PER_CHUNK = (DT_ == DT_BF16) ? 64 : 32;
  static constexpr int DPAD = KC_ * ELEMS_PER_CHUNK;  // padded feature width
  static constexpr int X_BYTES = KCS * 128 * 128;
  static constexpr int Y_BYTES = KCS * BN_ * 128;
  static constexpr int G_BYTES = (EPI_ == EPI_GRAD) ? (BN_ / 64) * 128 * 128 : 0;
  static constexpr int NG = 2;  // G double buffer (one per epilogue warpgroup)
  static constexpr int CTRL_BYTES = 4096;  // barriers + cross-warpgroup exchange
  static constexpr int SMEM_BYTES = X_BYTES + NS_ * Y_BYTES + NG * G_BYTES + CTRL_BYTES + 1024 /*align*/;
  static constexpr int ACC_COLS = (EPI_ == EPI_GRAD) ? DPAD : 0;
  static constexpr int TMEM_NEED = 2 * BN_ + ACC_COLS;
  static constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  static_assert(TMEM_NEED <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 227 * 1024, "SMEM budget");
  static_assert(BN_ % 64 == 0 && BN_ <= 256, "BN");
  static_assert(EPI_ != EPI_GRAD || DT_ == DT_BF16, "GRAD epilogue is bf16-only for now");
};

struct Control {
  uint64_t full[8], empty[8];
  uint64_t x_full, x_empty;
  uint64_t s_full[2], s_empty[2];
  uint64_t g_full[2], g_empty[2];
  uint64_t acc_full, acc_empty;
  uint32_t tmem_base;
  uint32_t pad_;
  float xchg[3][128];  // warpgroup 1 -> warpgroup 0 hand-over of per-row partials
};
static_assert(sizeof(Control) <= 4096, "control block");

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

__device__ __forceinline__ float max32(const uint32_t (&r)[32]) {
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    m[i] = fmaxf(fmaxf(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])),
                 fmaxf(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
  return fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
}

// named barrier over the 256 epilogue threads (id 1; id 0 is __syncthreads)
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(SWEEP_THREADS, 1)
sweep_kernel(const __grid_constant__ CUtensorMap tm_stat, const __grid_constant__ CUtensorMap tm_strm,
             const SweepArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* x_smem = smem;
  uint8_t* y_smem = x_smem + C::X_BYTES;
  uint8_t* g_smem = y_smem + C::NS * C::Y_BYTES;
  Control* bar = reinterpret_cast<Control*>(g_smem + C::NG * C::G_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_items = a.n_stat_tiles * a.n_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_stat);
    tma_prefetch_desc(&tm_strm);
    for (int i = 0; i < C::NS; ++i) { mbar_init(&bar->full[i], 1); mbar_init(&bar->empty[i], 1); }
    mbar_init(&bar->x_full, 1);
    mbar_init(&bar->x_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar->s_full[i], 1);
      mbar_init(&bar->s_empty[i], 128);
      mbar_init(&bar->g_full[i], 128);
      mbar_init(&bar->g_empty[i], 1);
    }
    mbar_init(&bar->acc_full, 1);
    mbar_init(&bar->acc_empty, 256);
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

  auto item_range = [&](int item, int& stat_tile, int& split, int& t0, int& t1) {
    stat_tile = item % a.n_stat_tiles;  // split-major order: concurrent CTAs share streamed tiles in L2
    split = item / a.n_stat_tiles;
    t0 = static_cast<int>((static_cast<long long>(split) * a.n_strm_tiles) / a.n_splits);
    t1 = static_cast<int>((static_cast<long long>(split + 1) * a.n_strm_tiles) / a.n_splits);
  };

  if (warp == 0) {
    // ======================================================================= TMA producer
    if (lane == 0) {
      uint32_t it = 0, k = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++k) {
        int stat_tile, split, t0, t1;
        item_range(item, stat_tile, split, t0, t1);
        mbar_wait(&bar->x_empty, (k & 1) ^ 1);
        mbar_arrive_expect_tx(&bar->x_full, C::X_BYTES);
#pragma unroll
        for (int c = 0; c < C::KCS; ++c)
          tma_load_2d(x_smem + c * 128 * 128, &tm_stat, &bar->x_full, c * C::ELEMS_PER_CHUNK, stat_tile * 128);
        for (int t = t0; t < t1; ++t, ++it) {
          const uint32_t st = it % C::NS, ph = (it / C::NS) & 1;
          mbar_wait(&bar->empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&bar->full[st], C::Y_BYTES);
#pragma unroll
          for (int c = 0; c < C::KCS; ++c)
            tma_load_2d(y_smem + st * C::Y_BYTES + c * C::BN * 128, &tm_strm, &bar->full[st],
                        c * C::ELEMS_PER_CHUNK, t * C::BN);
        }
      }
    }
  } else if (warp == 1) {
    // ========================================================================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t fmt = (C::DT == DT_BF16) ? FMT_BF16 : FMT_TF32;
      constexpr uint32_t idesc1 = make_idesc(fmt, 128, C::BN, 0, 0);
      // second MMA: A = softmax tile P as bf16 (packed exp2 in the epilogue), B = streamed bf16 tile
      constexpr uint32_t idesc2 = make_idesc(fmt, 128, C::DPAD, 0, 1);
      const uint32_t x_addr = smem_u32(x_smem), y_addr = smem_u32(y_smem), g_addr = smem_u32(g_smem);
      uint32_t it = 0, k = 0;

      auto issue_mma1 = [&](uint32_t tile_it) {
        const uint32_t st = tile_it % C::NS, ph = (tile_it / C::NS) & 1;
        const uint32_t buf = tile_it & 1, sph = (tile_it >> 1) & 1;
        mbar_wait(&bar->full[st], ph);
        mbar_wait(&bar->s_empty[buf], sph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * C::BN;
#pragma unroll
        for (int p = 0; p < C::NPAIR; ++p) {
          int ac, bc;
          if (C::DT == DT_BF16) { ac = p; bc = p; }
          else {
            const int c = p / 3, r = p % 3;  // small terms first: lo*hi, hi*lo, then hi*hi
            ac = (r == 0) ? C::KC + c : c;
            bc = (r == 1) ? C::KC + c : c;
          }
          const uint32_t ab = x_addr + ac * 128 * 128;
          const uint32_t bb = y_addr + st * C::Y_BYTES + bc * C::BN * 128;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ad = make_smem_desc(ab + kk * 32, 16, 1024);
            const uint64_t bd = make_smem_desc(bb + kk * 32, 16, 1024);
            if (C::DT == DT_BF16) mma_f16_ss(d_tmem, ad, bd, idesc1, (p | kk) != 0);
            else mma_tf32_ss(d_tmem, ad, bd, idesc1, (p | kk) != 0);
          }
        }
        tc_commit(&bar->s_full[buf]);
        if (C::EPI != EPI_GRAD) tc_commit(&bar->empty[st]);
      };
      auto issue_mma2 = [&](uint32_t tile_it, bool first_of_item) {
        const uint32_t st = tile_it % C::NS;
        const uint32_t gb = tile_it & 1, gph = (tile_it >> 1) & 1;
        mbar_wait(&bar->g_full[gb], gph);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + 2 * C::BN;
#pragma unroll
        for (int kk = 0; kk < C::BN / 16; ++kk) {
          const uint64_t ad = make_smem_desc(g_addr + gb * C::G_BYTES + (kk / 4) * 128 * 128 + (kk % 4) * 32, 16, 1024);
          const uint64_t bd = make_smem_desc(y_addr + st * C::Y_BYTES + kk * 2048, C::BN * 128, 1024);
          mma_f16_ss(d_tmem, ad, bd, idesc2, !(first_of_item && kk == 0));
        }
        tc_commit(&bar->g_empty[gb]);
        tc_commit(&bar->empty[st]);
      };

      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++k) {
        int stat_tile, split, t0, t1;
        item_range(item, stat_tile, split, t0, t1);
        mbar_wait(&bar->x_full, k & 1);
        if (C::EPI == EPI_GRAD) mbar_wait(&bar->acc_empty, (k & 1) ^ 1);
        const uint32_t it0 = it;
        for (int t = t0; t < t1; ++t, ++it) {
          issue_mma1(it);
          if (C::EPI == EPI_GRAD && it > it0) issue_mma2(it - 1, it - 1 == it0);
        }
        tc_commit(&bar->x_empty);
        if (C::EPI == EPI_GRAD) {
          issue_mma2(it - 1, it - 1 == it0);
          tc_commit(&bar->acc_full);
        }
      }
    }
  } else {
    // =========================================================================== epilogue
    const int wg = (warp - 2) >> 2;    // warpgroup 0/1 == parity of the tiles it owns == S/G buffer
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;       // stationary row within the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + wg * C::BN;
    const float c2 = a.scale * LOG2E;
    const bool plain = (a.bias == nullptr) && (c2 > 0.f);  // fast paths: no bias, positive scale
    constexpr int NCH = C::BN / 32;
    uint32_t it = 0, k = 0;

    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++k) {
      int stat_tile, split, t0, t1;
      item_range(item, stat_tile, split, t0, t1);
      const int srow = stat_tile * 128 + r;  // global stationary row
      const bool srow_ok = srow < a.n_stat;
      const long long pslot = static_cast<long long>(split) * a.n_stat_tiles * 128 + srow;

      // ---- per-item, per-warpgroup state
      float m2 = -INFINITY, l = 0.f, ll = 0.f;     // LSE
      int lab = -1;                                 // LSE / GRAD(rows)
      float my_nb = 0.f;                            // GRAD: -lse2 (rows) or bias*log2e (items)
      float rowsum = 0.f;                           // GRAD(items): sum_j P for dbias
      int seen_cur = 0, seen_end = 0, next_seen = 0x7fffffff;   // TOPK: cursor into the row's seen list

      if (C::EPI == EPI_LSE) lab = srow_ok ? a.labels[srow] : -1;
      if (C::EPI == EPI_GRAD) {
        if (C::STAT_ROWS) {
          lab = srow_ok ? a.labels[srow] : -1;
          my_nb = -a.lse2[srow];  // lse2 is +inf for rows >= n_stat => P = 0 there
        } else {
          my_nb = srow_ok ? ((a.bias != nullptr) ? a.bias[srow] * LOG2E : 0.f) : -INFINITY;
        }
      }
      if (C::EPI == EPI_TOPK) {
        if (a.seen_crow != nullptr && srow_ok) {
          seen_cur = a.seen_crow[srow];
          seen_end = a.seen_crow[srow + 1];
          const int first = t0 * C::BN;  // first streamed row of this split
          int lo = seen_cur, hi = seen_end;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (a.seen_col[mid] < first) lo = mid + 1; else hi = mid;
          }
          seen_cur = lo;
          next_seen = (seen_cur < seen_end) ? a.seen_col[seen_cur] : 0x7fffffff;
        }
      }

      for (int t = t0; t < t1; ++t, ++it) {
        if ((it & 1) != static_cast<uint32_t>(wg)) continue;
        const uint32_t sph = (it >> 1) & 1;
        const int col_base = t * C::BN;                       // first streamed row of the tile
        const int n_valid = min(C::BN, a.n_strm - col_base);  // valid columns in this tile
        const bool full_tile = (n_valid == C::BN);
        mbar_wait(&bar->s_full[wg], sph);
        tc_fence_after();
        if (C::EPI == EPI_GRAD) mbar_wait(&bar->g_empty[wg], sph ^ 1);

        float tmax = -INFINITY;   // TOPK: masked max of this tile for this row
        bool tile_quick = false;
        if (C::EPI == EPI_TOPK) {
          while (next_seen < col_base) {  // skip seen ids that fell into the other warpgroup's tiles
            ++seen_cur;
            next_seen = (seen_cur < seen_end) ? __ldg(a.seen_col + seen_cur) : 0x7fffffff;
          }
          tile_quick = plain && full_tile && (next_seen >= col_base + C::BN);
        }
        uint32_t raw[2][32];
        tmem_ld32(t_lane, raw[0]);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          tmem_ld_wait();
          if (ch + 1 < NCH) tmem_ld32(t_lane + (ch + 1) * 32, raw[(ch + 1) & 1]);
          const uint32_t (&v)[32] = raw[ch & 1];
          const int c0 = ch * 32;                  // first column of the chunk within the tile
          const int nv = n_valid - c0;             // valid columns in this chunk (may be <= 0 or >= 32)

          if (C::EPI == EPI_DENSE) {
            if (srow_ok && nv > 0) {
              float* o = a.out + static_cast<long long>(srow) * a.ld_out + col_base + c0;
              if (a.bias == nullptr && nv >= 32) {
#pragma unroll
                for (int c = 0; c < 32; ++c) o[c] = __uint_as_float(v[c]) * a.scale;
              } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                  if (c < nv) {
                    float s = __uint_as_float(v[c]) * a.scale;
                    if (a.bias != nullptr) s += __ldg(a.bias + col_base + c0 + c);
                    o[c] = s;
                  }
                }
              }
            }
          } else if (C::EPI == EPI_LSE) {
            const int rel = lab - (col_base + c0);
            if (plain && full_tile) {
              const float cm2 = max32(v) * c2;
              if (cm2 > m2) { l *= ex2_approx(m2 - cm2); m2 = cm2; }
              const float nm = -m2;
              float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int c = 0; c < 32; ++c) acc[c & 3] += ex2_approx(fmaf(__uint_as_float(v[c]), c2, nm));
              l += (acc[0] + acc[1]) + (acc[2] + acc[3]);
              if (__any_sync(0xffffffffu, static_cast<uint32_t>(rel) < 32u)) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if (c == rel) ll = __uint_as_float(v[c]) * a.scale;
              }
            } else if (nv > 0) {
              float x[32];
              float cmax = -INFINITY;
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                float b2 = 0.f;
                if (a.bias != nullptr) b2 = (c < nv) ? __ldg(a.bias + col_base + c0 + c) * LOG2E : 0.f;
                x[c] = fmaf(__uint_as_float(v[c]), c2, b2);
                if (c >= nv) x[c] = -INFINITY;
                cmax = fmaxf(cmax, x[c]);
              }
              const float m_new = fmaxf(m2, cmax);
              float acc = 0.f;
#pragma unroll
              for (int c = 0; c < 32; ++c) acc += ex2_approx(x[c] - m_new);
              l = l * ex2_approx(m2 - m_new) + acc;
              m2 = m_new;
              if (__any_sync(0xffffffffu, static_cast<uint32_t>(rel) < 32u)) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if (c == rel) ll = x[c] * LN2;  // natural-log units: scale*s + bias
              }
            }
          } else if (C::EPI == EPI_GRAD) {
            // P = 2^(s*c2 + bias2 - lse2) as packed bf16 pairs {col 2i (low), col 2i+1 (high)}
            uint32_t ph[16];
            if (C::STAT_ROWS) {
              if (plain && full_tile) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  ph[i] = ex2_bf16x2(fmaf(__uint_as_float(v[2 * i]), c2, my_nb), fmaf(__uint_as_float(v[2 * i + 1]), c2, my_nb));
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  float x[2];
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    const int c = 2 * i + h;
                    float nb = my_nb;
                    if (a.bias != nullptr) nb += (c < nv) ? __ldg(a.bias + col_base + c0 + c) * LOG2E : 0.f;
                    x[h] = (c < nv) ? fmaf(__uint_as_float(v[c]), c2, nb) : -INFINITY;
                  }
                  ph[i] = ex2_bf16x2(x[0], x[1]);
                }
              }
              const int rel = lab - (col_base + c0);
              if (__any_sync(0xffffffffu, static_cast<uint32_t>(rel) < 32u)) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if (c == rel) ph[c >> 1] = bsub2_u32(ph[c >> 1], (c & 1) ? 0x3F800000u : 0x00003F80u);  // -= 1.0
              }
            } else {
              // columns are query rows: lse2 is +inf beyond n_strm => P = 0 there (no tail special case)
              const float4* l4 = reinterpret_cast<const float4*>(a.lse2 + col_base + c0);
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4) {
                const float4 w = __ldg(l4 + c4);
                ph[c4 * 2 + 0] = ex2_bf16x2(fmaf(__uint_as_float(v[c4 * 4 + 0]), c2, my_nb - w.x),
                                           fmaf(__uint_as_float(v[c4 * 4 + 1]), c2, my_nb - w.y));
                ph[c4 * 2 + 1] = ex2_bf16x2(fmaf(__uint_as_float(v[c4 * 4 + 2]), c2, my_nb - w.z),
                                           fmaf(__uint_as_float(v[c4 * 4 + 3]), c2, my_nb - w.w));
              }
              // one-hot: does any query row of this chunk have its label inside this item tile?
              int labc = -1;
              if (col_base + c0 + lane < a.n_strm) labc = __ldg(a.labels + col_base + c0 + lane);
              const uint32_t hit = __ballot_sync(0xffffffffu, static_cast<uint32_t>(labc - stat_tile * 128) < 128u);
              if (hit != 0) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                  if (hit & (1u << c)) {
                    const int lc = __shfl_sync(0xffffffffu, labc, c);
                    if (lc == srow) ph[c >> 1] = bsub2_u32(ph[c >> 1], (c & 1) ? 0x3F800000u : 0x00003F80u);
                  }
                }
              }
              if (a.rowsum_out != nullptr) {
                float rs[2] = {0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float2 f = b2_to_f2(ph[i]);
                  rs[0] += f.x;
                  rs[1] += f.y;
                }
                rowsum += rs[0] + rs[1];
              }
            }
            // G tile, K-major, 128-byte swizzle: [k-chunk of 64][row][64 x bf16]
            uint8_t* gdst = g_smem + wg * C::G_BYTES + (c0 / 64) * 128 * 128 + r * 128;
            const int v0 = (c0 % 64) / 8;
#pragma unroll
            for (int vv = 0; vv < 4; ++vv) {
              const uint4 w = make_uint4(ph[vv * 4 + 0], ph[vv * 4 + 1], ph[vv * 4 + 2], ph[vv * 4 + 3]);
              *reinterpret_cast<uint4*>(gdst + (((v0 + vv) ^ (r & 7)) << 4)) = w;
            }
          } else if (C::EPI == EPI_TOPK) {
            if (tile_quick) {
              tmax = fmaxf(tmax, max32(v));  // raw scores; scaled once per tile (scale > 0)
            } else if (nv > 0) {
              float x[32];
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                float sc = __uint_as_float(v[c]) * a.scale;
                if (a.bias != nullptr && c < nv) sc += __ldg(a.bias + col_base + c0 + c);
                x[c] = (c < nv) ? sc : -INFINITY;
              }
              while (next_seen < col_base + c0 + 32) {  // seen ids inside this chunk (ascending)
                const int rel = next_seen - (col_base + c0);
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if (c == rel) x[c] = -INFINITY;
                ++seen_cur;
                next_seen = (seen_cur < seen_end) ? __ldg(a.seen_col + seen_cur) : 0x7fffffff;
              }
              float cm = -INFINITY;
#pragma unroll
              for (int c = 0; c < 32; ++c) cm = fmaxf(cm, x[c]);
              tmax = fmaxf(tmax, cm);
            }
          }
        }  // chunks

        // release the S buffer (all tcgen05.ld of this thread have completed)
        tc_fence_before();
        mbar_arrive(&bar->s_empty[wg]);
        if (C::EPI == EPI_TOPK && srow_ok)
          a.tile_max[static_cast<long long>(srow) * a.n_strm_tiles + t] = tile_quick ? tmax * a.scale : tmax;
        if (C::EPI == EPI_GRAD) {
          fence_proxy_async_smem();
          mbar_arrive(&bar->g_full[wg]);
        }
      }  // tiles

      // ---- per-item outputs
      if (C::EPI == EPI_LSE) {
        // warpgroup 1 hands its partial to warpgroup 0, which merges and writes one slot per row
        if (wg == 1) { bar->xchg[0][r] = m2; bar->xchg[1][r] = l; bar->xchg[2][r] = ll; }
        epi_bar_sync();
        if (wg == 0) {
          const float om = bar->xchg[0][r], ol = bar->xchg[1][r], oll = bar->xchg[2][r];
          const float mm = fmaxf(m2, om);
          float lm = 0.f;
          if (mm > -INFINITY) lm = l * ex2_approx(m2 - mm) + ol * ex2_approx(om - mm);
          a.part_m2[pslot] = mm;
          a.part_l[pslot] = lm;
          a.part_ll[pslot] = ll + oll;
        }
        epi_bar_sync();  // xchg is free again before the next item
      } else if (C::EPI == EPI_GRAD) {
        mbar_wait(&bar->acc_full, k & 1);
        tc_fence_after();
        const float gsc = a.gscale * (a.gscale_dev != nullptr ? __ldg(a.gscale_dev) : 1.f);
        float* o = a.acc_out + (static_cast<long long>(split) * a.n_stat + srow) * a.d;
        const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 2 * C::BN;
#pragma unroll 1
        for (int ch = wg; ch < C::DPAD / 32; ch += 2) {  // the two warpgroups split the columns
          uint32_t rawa[32];
          tmem_ld32(t_acc + ch * 32, rawa);
          tmem_ld_wait();
          if (srow_ok) {
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              const int col = ch * 32 + c4 * 4;
              if (col < a.d) {  // d % 4 == 0 (checked on host)
                float4 w;
                w.x = __uint_as_float(rawa[c4 * 4 + 0]) * gsc;
                w.y = __uint_as_float(rawa[c4 * 4 + 1]) * gsc;
                w.z = __uint_as_float(rawa[c4 * 4 + 2]) * gsc;
                w.w = __uint_as_float(rawa[c4 * 4 + 3]) * gsc;
                *reinterpret_cast<float4*>(o + col) = w;
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&bar->acc_empty);
        if (!C::STAT_ROWS && a.rowsum_out != nullptr) {
          if (wg == 1) bar->xchg[0][r] = rowsum;
          epi_bar_sync();
          if (wg == 0 && srow_ok)
            a.rowsum_out[static_cast<long long>(split) * a.n_stat + srow] = (rowsum + bar->xchg[0][r]) * gsc;
          epi_bar_sync();
        }
      }
    }  // items
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

}  // namespace rb
