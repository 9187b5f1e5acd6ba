// The scoring sweep engine: one warp-specialised tcgen05 kernel template that computes tiles of
//     S = X_stat (128 rows) x Y_strm^T (BN rows)      (both operands K-major, fed by TMA)
// for a stationary 128-row tile against a range of streamed tiles, with three fused epilogues
// (the training passes with their second MMA live in pair.cuh):
//
//   EPI_DENSE  store S (compat path for recommend_from_full, reference SASRec/main.py:228)
//   EPI_LSE    online (max, sum-exp) + label-logit pick   (F.cross_entropy fwd, SASRec/main.py:217-219)
//   EPI_TOPK   maximum of every (row, 128-item tile) of a short PREFIX of the catalog (a few percent): the
//              seed of the exact top-K (UniSRec/main.py:408-435 without dense (B,N)).  The maximum runs over the
//              row's UNSEEN items (the seen ids of a chunk, read off the sorted list as the sweep passes them,
//              are set to -inf; with scale <= 0 a tile holding a seen id is reported as NaN, "dirty", instead),
//              so a user with thousands of seen ids still gets a threshold.  simt.cuh turns the tile maxima of the prefix into the row's
//              starting threshold tau0 = K-th largest (>= K unseen items reach it) and a ladder of
//              checkpoints c_k = (K >> k)-th largest.
//   EPI_CAND   the one sweep over the whole catalog: every (row, aligned group of 4 items) whose maximum
//              reaches the row's RUNNING threshold goes to the (row, split[, warpgroup]) sub-list as
//              (group maximum, group id; groups that hold a seen id are marked dirty and support no count).  The threshold climbs the ladder while the sweep runs: a thread
//              counts its clean candidates that reach the next two checkpoints in registers, publishes the
//              counts every fourth tile (two REDs into the row's shared counters -- all splits of a row share
//              them) and re-reads the counters (the loads in flight behind the tile's arithmetic); once >= K
//              unseen items are known to reach c_k the row's threshold becomes c_k.  About
//              K (2 + log2(N/prefix)) candidates per row instead of the K ln(N/K) of an exact running K-th
//              best -- with a hit path of a dozen instructions per candidate: the epilogue warps are the
//              critical resource of this sweep (measured: 0.68 ms without candidates; every extra 50
//              instructions per candidate cost 0.25 ms; shuffles, local-memory scratch or a fully unrolled
//              per-item path cost 2-12x more).  The finish kernel keeps the groups that can still matter
//              (maximum >= the K-th largest clean group maximum), re-scores them exactly, drops seen items
//              and sorts.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2..9 = two epilogue warpgroups (thread <-> TMEM lane <-> stationary row).
//   XT = 1: one stationary 128-row tile per CTA; warpgroup g owns the S buffer g, i.e. the streamed tiles
//           of parity g, so two tiles are always in flight per SM.
//   XT = 2: two stationary tiles X0, X1 per CTA (256 rows); every streamed tile feeds two MMAs and
//           warpgroup g owns tile X_g with its own double-buffered S.  Each streamed byte fetched from L2
//           now does twice the work: with XT = 1 a 4096-row sweep pulls 32 x the table through the
//           L2->SM fabric (8.2 GB per sweep at N = 1M, d = 128: measured 9.4 TB/s, the binding limit of
//           the plain sweeps), XT = 2 halves that and leaves the tensor pipe as the bound.
#pragma once
#include "ptx.cuh"

namespace rb {

enum : int { EPI_DENSE = 0, EPI_LSE = 1, EPI_TOPK = 3, EPI_CAND = 4 };
enum : int { DT_BF16 = 0, DT_TF32X3 = 1 };

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float MASKED_SCORE = -1e23f;  // UniSRec/main.py:413

// Per-row threshold ladder of the top-K candidate sweep (64 bytes = two sectors):
//   thr[0] = tau0, thr[1..TOPK_LEVELS] = checkpoints (non-decreasing; +inf = unused), thr[7] = +inf
//   cnt[l] = clean candidate groups published so far whose maximum reaches thr[l]  (l >= 1; a lower bound)
constexpr int TOPK_LEVELS = 6;
struct __align__(32) RowLadder {
  float thr[8];
  unsigned int cnt[8];
};
constexpr unsigned int CAND_DIRTY = 0x80000000u;   // the group's tile holds a seen id of the row

struct SweepArgs {
  int n_stat;        // valid rows of the stationary operand
  int n_strm;        // valid rows of the streamed operand
  int n_stat_tiles;  // stationary units: ceil(n_stat / (128 * XT))
  int n_strm_tiles;  // ceil(n_strm / BN)
  int n_splits;      // the streamed range of every stationary tile is cut into n_splits work items
  int d;             // true feature width (<= KC*64)
  float scale;       // logits = scale * <u,w> + bias
  const float* bias; // per ITEM bias (nullable)
  // device-side count of stationary (query) rows, nullable: stationary tiles beyond it are skipped (EPI_LSE)
  const int* m_dev;
  // EPI_DENSE (rows stationary)
  float* out;
  long long ld_out;
  // EPI_LSE (rows stationary)
  const int* labels;  // per query row: local item index, or -1
  float* part_m2;     // [n_splits][n_stat_tiles*128] running max, log2 domain
  float* part_l;      //   "  sum of 2^(x - m2)
  float* part_ll;     //   "  label logit (natural units) or 0
  // EPI_TOPK (rows stationary): pass 1 of the top-K = masked maximum of every (row, 128-item tile)
  const int* seen_crow;  // [n_stat+1] CSR of already-seen LOCAL item ids, sorted per row (nullable)
  const int* seen_col;
  float* tile_max;    // [n_stat][n_strm_tiles]
  // EPI_CAND (rows stationary): the candidate sweep of the top-K
  RowLadder* ladder;                // [n_stat] thresholds (read-only here) + shared level counters (RED + re-read)
  int k_need;                       // K: a checkpoint becomes the threshold once this many unseen items reach it
  uint2* cand;                      // [n_stat][n_sub][cand_cap] (group maximum bits, item id >> 2 | dirty << 31)
  int* cand_cnt;                    // [n_stat][n_sub] groups found per sub-list (> cand_cap: overflow)
  int cand_cap;
  int n_sub;                        // 2*n_splits: one sub-list per (split, tile parity)
};

// key order: score desc, then id asc (shared by every top-K stage)
__device__ __forceinline__ unsigned long long topk_key(float v, int id) {
  return (static_cast<unsigned long long>(f32_orderable(v)) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(id));
}
// logit of one (row, item): the single expression both top-K passes use, so they agree bit for bit
__device__ __forceinline__ float logit_of(uint32_t raw, float scale, const float* bias, int col) {
  if (bias != nullptr) return __fmaf_rn(__uint_as_float(raw), scale, __ldg(bias + col));
  return __fmul_rn(__uint_as_float(raw), scale);  // explicit roundings: no contraction differences between call sites
}

template <int EPI_, int DT_, int KC_, int BN_, int NS_, bool STAT_ROWS_, int XT_ = 1, int NWG_ = 2, bool MAYBE_BIAS_ = true>
struct SweepCfg {
  static constexpr int EPI = EPI_, DT = DT_, KC = KC_, BN = BN_, NS = NS_, XT = XT_;
  // false: the launch has no bias vector -- the bias branches of the epilogues are compiled out, so that the hot
  // bias-free kernels (96 registers per thread with four warpgroups) do not pay registers or spills for them
  static constexpr bool MAYBE_BIAS = MAYBE_BIAS_;
  // Epilogue warpgroups.  2: XT = 1 -> one per tile parity, XT = 2 -> one per stationary tile (both of its S buffers).
  // 4 (XT = 2 only): two per stationary tile, one per tile parity, i.e. one warpgroup per S buffer -- twice the issue
  // slots and latency tolerance for epilogues that are bound by their own instruction stream (the top-K sweeps).
  static constexpr int NWG = NWG_;
  static constexpr int THREADS = 64 + 128 * NWG_;
  static_assert(NWG_ == 2 || (NWG_ == 4 && XT_ == 2), "epilogue warpgroups");
  static constexpr bool STAT_ROWS = STAT_ROWS_;  // true: queries stationary, items streamed
  // storage chunks (128-byte columns groups) per operand row
  static constexpr int KCS = (DT_ == DT_BF16) ? KC_ : 2 * KC_;  // tf32x3: [hi | lo], KC = d/32
  static constexpr int NPAIR = (DT_ == DT_BF16) ? KC_ : 3 * KC_;
  static constexpr int ELEMS_PER_CHUNK = (DT_ == DT_BF16) ? 64 : 32;
  static constexpr int DPAD = KC_ * ELEMS_PER_CHUNK;  // padded feature width
  static constexpr int XTILE_BYTES = KCS * 128 * 128;   // one stationary 128-row tile
  static constexpr int X_BYTES = XT_ * XTILE_BYTES;
  // One pipeline stage = SC 128-byte K-chunks of a streamed tile (BN rows x 128 B each).  Tiles of up to two
  // chunks travel whole (SC = KCS); wider tiles (d = 256 bf16, d = 64 fp32x3: four chunks) travel chunk by
  // chunk (SC = 1): the MMAs of a chunk start as soon as it lands and its slot goes back to the producer as
  // soon as they retire, so the ring stays deep even when X takes 128 KB.
  static constexpr int SC = (KCS <= 2) ? KCS : 1;
  static constexpr int SPT = KCS / SC;              // stages per streamed tile
  static constexpr int STAGE_BYTES = SC * BN_ * 128;
  static constexpr int CTRL_BYTES = 4096;  // barriers + cross-warpgroup exchange
  static constexpr int SMEM_BYTES = X_BYTES + NS_ * STAGE_BYTES + CTRL_BYTES + 1024 /*align*/;
  static constexpr int TMEM_NEED = 2 * XT_ * BN_;
  static constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  static_assert(TMEM_NEED <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 227 * 1024, "SMEM budget");
  static_assert(BN_ % 64 == 0 && BN_ <= 256, "BN");
  static_assert(XT_ == 1 || XT_ == 2, "XT");
  static_assert(NS_ <= 16 && NS_ >= 2, "stage count");
};

struct Control {
  uint64_t full[16], empty[16];
  uint64_t x_full, x_empty;
  uint64_t s_full[4], s_empty[4];
  uint32_t tmem_base;
  uint32_t pad_;
  float xchg[2][3][128];  // [stationary tile][m2, l, ll][row]: hand-over of per-row partials between the two parities' warpgroups
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

// maxima of the four aligned groups of 8 values of a 32-value chunk
__device__ __forceinline__ void group_max8(const uint32_t (&r)[32], float (&m)[4]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float a = fmax3(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1]), __uint_as_float(r[8 * g + 2]));
    const float b = fmax3(__uint_as_float(r[8 * g + 3]), __uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
    m[g] = fmaxf(fmax3(a, b, __uint_as_float(r[8 * g + 6])), __uint_as_float(r[8 * g + 7]));
  }
}

// named barrier over the 256 epilogue threads that share a stationary tile (ids 1, 2; id 0 is __syncthreads)
__device__ __forceinline__ void epi_bar_sync(int xsel = 0) { asm volatile("bar.sync %0, 256;" ::"r"(xsel + 1) : "memory"); }

// ---------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
sweep_kernel(const __grid_constant__ CUtensorMap tm_stat, const __grid_constant__ CUtensorMap tm_strm,
             const SweepArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* x_smem = smem;
  uint8_t* y_smem = x_smem + C::X_BYTES;
  Control* bar = reinterpret_cast<Control*>(y_smem + C::NS * C::STAGE_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int n_stat = a.n_stat, n_stat_tiles = a.n_stat_tiles;   // effective extents under a device-side row count
  if (a.m_dev != nullptr) {
    n_stat = min(n_stat, max(0, *a.m_dev));
    n_stat_tiles = min(n_stat_tiles, (n_stat + 128 * C::XT - 1) / (128 * C::XT));
  }
  const int total_items = n_stat_tiles * a.n_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_stat);
    tma_prefetch_desc(&tm_strm);
    for (int i = 0; i < C::NS; ++i) { mbar_init(&bar->full[i], 1); mbar_init(&bar->empty[i], 1); }
    mbar_init(&bar->x_full, 1);
    mbar_init(&bar->x_empty, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bar->s_full[i], 1);
      mbar_init(&bar->s_empty[i], 128);
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

  auto item_range = [&](int item, int& stat_tile, int& split, int& t0, int& t1) {
    stat_tile = item % n_stat_tiles;  // split-major order: concurrent CTAs share streamed tiles in L2
    split = item / n_stat_tiles;
    t0 = static_cast<int>((static_cast<long long>(split) * a.n_strm_tiles) / a.n_splits);
    t1 = static_cast<int>((static_cast<long long>(split + 1) * a.n_strm_tiles) / a.n_splits);
  };

  if (warp == 0) {
    // ======================================================================= TMA producer
    // The whole warp runs the (uniform) control flow; one elected lane issues the copies.
    uint32_t cs = 0, k = 0;   // cs: stages issued so far
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++k) {
      int stat_tile, split, t0, t1;
      item_range(item, stat_tile, split, t0, t1);
      mbar_wait(&bar->x_empty, (k & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bar->x_full, C::X_BYTES);
#pragma unroll
        for (int x = 0; x < C::XT; ++x)
#pragma unroll
          for (int c = 0; c < C::KCS; ++c)
            tma_load_2d(x_smem + x * C::XTILE_BYTES + c * 128 * 128, &tm_stat, &bar->x_full, c * C::ELEMS_PER_CHUNK,
                        (stat_tile * C::XT + x) * 128);
      }
      __syncwarp();
      for (int t = t0; t < t1; ++t) {
#pragma unroll
        for (int sg = 0; sg < C::SPT; ++sg, ++cs) {
          const uint32_t st = cs % C::NS, ph = (cs / C::NS) & 1;
          mbar_wait(&bar->empty[st], ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&bar->full[st], C::STAGE_BYTES);
#pragma unroll
            for (int c = 0; c < C::SC; ++c)
              tma_load_2d(y_smem + st * C::STAGE_BYTES + c * C::BN * 128, &tm_strm, &bar->full[st],
                          (sg * C::SC + c) * C::ELEMS_PER_CHUNK, t * C::BN);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ========================================================================= MMA issuer
    // Warp-converged loop (descriptor arithmetic stays in uniform registers); one elected lane issues
    // the tcgen05.mma / commit instructions.
    constexpr uint32_t fmt = (C::DT == DT_BF16) ? FMT_BF16 : FMT_TF32;
    constexpr uint32_t idesc1 = make_idesc(fmt, 128, C::BN, 0, 0);
    constexpr uint32_t dhi = smem_desc_hi(1024);
    const uint32_t x_lo = smem_desc_lo(smem_u32(x_smem), 16), y_lo = smem_desc_lo(smem_u32(y_smem), 16);
    uint32_t it = 0, k = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++k) {
      int stat_tile, split, t0, t1;
      item_range(item, stat_tile, split, t0, t1);
      mbar_wait(&bar->x_full, k & 1);
      for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t buf = it & 1, sph = (it >> 1) & 1;
        const uint32_t cs = it * C::SPT;   // first stage of this streamed tile
#pragma unroll
        for (int x = 0; x < C::XT; ++x)    // S buffer: XT=1 per tile parity, XT=2 per (X tile, parity)
          mbar_wait(&bar->s_empty[(C::XT == 1) ? buf : x * 2 + buf], sph ^ 1);
        uint32_t waited = 0;               // stages of this tile already seen full
#pragma unroll
        for (int p = 0; p < C::NPAIR; ++p) {
          int ac, bc, last_use;            // X chunk, Y chunk, last pair that reads Y chunk bc
          if (C::DT == DT_BF16) { ac = p; bc = p; last_use = p; }
          else {
            const int c = p / 3, r = p % 3;  // small terms first: lo*hi, hi*lo, then hi*hi
            ac = (r == 0) ? C::KC + c : c;
            bc = (r == 1) ? C::KC + c : c;
            last_use = (r == 1) ? p : 3 * c + 2;
          }
          const int sg = bc / C::SC;       // stage of the tile that holds Y chunk bc
          if (C::SC > 1) last_use = C::NPAIR - 1;   // whole-tile stages are released after the tile's last MMA
          const uint32_t sidx = cs + sg, st = sidx % C::NS;
          if (!((waited >> sg) & 1u)) {
            mbar_wait(&bar->full[st], (sidx / C::NS) & 1);
            waited |= 1u << sg;
          }
          tc_fence_after();
          if (elect_one()) {
            const uint32_t ys_lo = y_lo + ((st * C::STAGE_BYTES + (bc % C::SC) * C::BN * 128) >> 4);
#pragma unroll
            for (int x = 0; x < C::XT; ++x) {
              const uint32_t bidx = (C::XT == 1) ? buf : x * 2 + buf;
              const uint32_t d_tmem = tmem_base + bidx * C::BN;
              const uint32_t xs_lo = x_lo + ((x * C::XTILE_BYTES + ac * 128 * 128) >> 4);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t ad = smem_desc(dhi, xs_lo + ((kk * 32) >> 4));
                const uint64_t bd = smem_desc(dhi, ys_lo + ((kk * 32) >> 4));
                if (C::DT == DT_BF16) mma_f16_ss(d_tmem, ad, bd, idesc1, (p | kk) != 0);
                else mma_tf32_ss(d_tmem, ad, bd, idesc1, (p | kk) != 0);
              }
              if (p == C::NPAIR - 1) tc_commit(&bar->s_full[bidx]);
            }
            if (p == last_use) tc_commit(&bar->empty[st]);   // every MMA that reads this stage retires before this fires
          }
          __syncwarp();
        }
      }
      if (elect_one()) tc_commit(&bar->x_empty);
      __syncwarp();
    }
  } else {
    // =========================================================================== epilogue
    const int wgi = (warp - 2) >> 2;   // epilogue warpgroup
    // xsel: the stationary tile of the CTA this warpgroup serves; par: the parity of the streamed tiles it handles
    // (-1 = all of them)
    const int xsel = (C::XT == 1) ? 0 : (C::NWG == 4 ? (wgi & 1) : wgi);
    const int par = (C::XT == 1) ? wgi : (C::NWG == 4 ? (wgi >> 1) : -1);
    constexpr int SUBS = (C::XT == 1 || C::NWG == 4) ? 2 : 1;   // candidate sub-lists per (row, split)
    const int subidx = (par < 0) ? 0 : par;
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;       // stationary row within the tile == TMEM lane
    const uint32_t t_lane0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c2 = a.scale * LOG2E;
    const float* const bias = C::MAYBE_BIAS ? a.bias : nullptr;
    const bool plain = (bias == nullptr) && (c2 > 0.f);  // fast paths: no bias, positive scale
    // bias head (BERT4Rec/main.py:83): the 32 biases of a full chunk come as eight 16-byte loads (the address is the same
    // for every row of the warp) when the vector is 16-byte aligned; element by element otherwise
    const bool bias_vec = (bias != nullptr) && ((reinterpret_cast<uintptr_t>(bias) & 15) == 0);
    constexpr int NCH = C::BN / 32;
    uint32_t it = 0, k = 0;

    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++k) {
      int stat_tile, split, t0, t1;
      item_range(item, stat_tile, split, t0, t1);
      const int stile = (C::XT == 1) ? stat_tile : stat_tile * 2 + xsel;   // global 128-row stationary tile
      const int srow = stile * 128 + r;  // global stationary row
      const bool srow_ok = srow < n_stat;
      const long long pslot = static_cast<long long>(split) * a.n_stat_tiles * (128 * C::XT) + srow;

      // ---- per-item, per-warpgroup state
      float m2 = -INFINITY, l = 0.f, ll = 0.f;     // LSE
      int lab = -1;                                 // LSE
      int seen_cur = 0, seen_end = 0, next_seen = 0x7fffffff;   // TOPK: cursor into the row's seen list
      int next2_seen = 0x7fffffff;                              //       (one entry prefetched: no load latency on advance)
      int n_cand = 0;                                           // CAND: entries in this thread's sub-list
      uint2* cand_list = nullptr;
      float tau = INFINITY;                         // CAND: running threshold = thr[lvl]; rows beyond n_stat never hit
      float ck1 = INFINITY, ck2 = INFINITY;         // CAND: the next two checkpoints thr[lvl+1], thr[lvl+2]
      int lvl = 0;                                  // CAND: rung of the ladder this thread stands on
      unsigned int n1 = 0u, n2 = 0u;                // CAND: unpublished clean candidates reaching c1 / c2
      RowLadder* ld = nullptr;                      // CAND: the row's ladder (thresholds + shared counters)
      uint32_t my_tiles = 0;                        // CAND: tiles this thread has handled in this work item

      if (C::EPI == EPI_CAND && srow_ok) {
        ld = a.ladder + srow;
        tau = __ldg(ld->thr); ck1 = __ldg(ld->thr + 1); ck2 = __ldg(ld->thr + 2);
      }
      if (C::EPI == EPI_LSE) lab = srow_ok ? a.labels[srow] : -1;
      if (C::EPI == EPI_TOPK || C::EPI == EPI_CAND) {
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
          next2_seen = (seen_cur + 1 < seen_end) ? a.seen_col[seen_cur + 1] : 0x7fffffff;
        }
      }
      if (C::EPI == EPI_CAND)
        cand_list = a.cand + (static_cast<long long>(srow) * a.n_sub + split * SUBS + subidx) * a.cand_cap;
      // advance the seen cursor by one entry; the entry after next is already in a register
      auto seen_advance = [&]() {
        ++seen_cur;
        next_seen = next2_seen;
        next2_seen = (seen_cur + 1 < seen_end) ? __ldg(a.seen_col + seen_cur + 1) : 0x7fffffff;
      };

      // ---- EPI_CAND: the parked hit chunk of this thread (see the chunk loop) and its service routine
      float pk_m[8], pk_mul = 1.f;
      unsigned int pk_item0 = 0u;    // first item of the parked chunk
      bool pk_pend = false;
      auto serve_parked = [&]() {
        // which of the chunk's 32 columns are seen ids of this row: read off the sorted list from the cursor, which
        // still stands at the first seen id of the TILE (it moves on after the tile) -- on the rare hit path only,
        // the per-tile loop pays nothing for it
        unsigned int pk_seen = 0u;
        if (next_seen < static_cast<int>(pk_item0) + 32) {
          int j = seen_cur, id = next_seen;
          while (id < static_cast<int>(pk_item0) + 32) {
            if (id >= static_cast<int>(pk_item0)) pk_seen |= 1u << (id - static_cast<int>(pk_item0));
            ++j;
            id = (j < seen_end) ? __ldg(a.seen_col + j) : 0x7fffffff;
          }
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float sg = __fmul_rn(pk_m[g], pk_mul);
          if (sg >= tau) {
            // a group of 4 that holds a seen id is "dirty": its maximum may belong to the seen item, so it is listed (the
            // finish kernel re-scores it item by item) but supports no checkpoint.  Dirtiness is per GROUP: a user with
            // thousands of seen ids still has clean groups, counts, and a threshold that climbs.
            const bool dirty = ((pk_seen >> (4 * g)) & 0xFu) != 0u;
            if (n_cand < a.cand_cap)
              cand_list[n_cand] = make_uint2(__float_as_uint(sg), ((pk_item0 >> 2) + g) | (dirty ? CAND_DIRTY : 0u));
            ++n_cand;
            if (!dirty) {  // a clean group's maximum is an unseen item: it supports the checkpoints it reaches
              n1 += (sg >= ck1) ? 1u : 0u;
              n2 += (sg >= ck2) ? 1u : 0u;
            }
          }
        }
        pk_pend = false;
      };

      for (int t = t0; t < t1; ++t, ++it) {
        if (par >= 0 && (it & 1) != static_cast<uint32_t>(par)) continue;
        const uint32_t sph = (it >> 1) & 1;
        const uint32_t bidx = (C::XT == 1) ? par : xsel * 2 + (it & 1);
        const uint32_t t_lane = t_lane0 + bidx * C::BN;
        const int col_base = t * C::BN;                       // first streamed row of the tile
        const int n_valid = min(C::BN, a.n_strm - col_base);  // valid columns in this tile
        const bool full_tile = (n_valid == C::BN);
        // CAND: every fourth tile the thread re-reads its row's level counters (L2: other CTAs bump them);
        // the two loads are in flight behind this tile's arithmetic and are consumed after it
        bool refresh = false;
        uint4 h_lo = make_uint4(0u, 0u, 0u, 0u), h_hi = make_uint4(0u, 0u, 0u, 0u);
        if (C::EPI == EPI_CAND) {
          refresh = (ld != nullptr) && ((my_tiles++ & 3u) == 3u);
          if (refresh) {
            h_lo = __ldcg(reinterpret_cast<const uint4*>(ld->cnt));
            h_hi = __ldcg(reinterpret_cast<const uint4*>(ld->cnt) + 1);
          }
        }
        if (C::EPI == EPI_CAND || C::EPI == EPI_TOPK)
          while (next_seen < col_base) seen_advance();   // seen ids that fell into the other warpgroup's tiles
        // TOPK: the tile maximum is taken over the row's UNSEEN items (seen columns are set to -inf chunk by chunk),
        // which needs scale > 0; otherwise a tile that holds a seen id is reported NaN ("dirty") as a whole
        const bool mask_exact = a.scale > 0.f;
        bool tile_dirty = false;
        mbar_wait(&bar->s_full[bidx], sph);
        tc_fence_after();
        float tmax = -INFINITY;   // TOPK: max of this tile for this row
        const bool tile_quick = plain && full_tile;   // warp-uniform
        // With two epilogue warpgroups the next chunk's TMEM read is in flight behind the current chunk's arithmetic
        // (double-buffered registers); with four, twice as many warps hide each other's reads and the registers go to
        // the epilogue's state instead.  (Measured: pulling all 128 columns into registers first and releasing the S
        // buffer before the arithmetic is SLOWER, 1.12 vs 0.97 ms -- the warp idles through its own TMEM reads.)
        constexpr bool DB = (C::NWG == 2);
        uint32_t raw[DB ? 2 : 1][32];
        if (DB) tmem_ld32(t_lane, raw[0]);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          if (DB) {
            tmem_ld_wait();
            if (ch + 1 < NCH) tmem_ld32(t_lane + (ch + 1) * 32, raw[DB ? ((ch + 1) & 1) : 0]);
          } else {
            tmem_ld32(t_lane + ch * 32, raw[0]);
            tmem_ld_wait();
          }
          uint32_t (&v)[32] = raw[DB ? (ch & 1) : 0];
          const int c0 = ch * 32;                  // first column of the chunk within the tile
          const int nv = n_valid - c0;             // valid columns in this chunk (may be <= 0 or >= 32)
          unsigned int seen32 = 0u;                // TOPK: the row's seen ids among the chunk's 32 columns
          if (C::EPI == EPI_TOPK) {
            while (next_seen < col_base + c0 + 32) {   // sorted list, chunks in order: usually no iteration
              seen32 |= 1u << (next_seen - (col_base + c0));
              seen_advance();
            }
            if (seen32 != 0u) {
              tile_dirty = true;
              if (mask_exact) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if ((seen32 >> c) & 1u) v[c] = 0xff800000u;   // -inf: a seen item never carries a tile maximum
              }
            }
          }

          if (C::EPI == EPI_DENSE) {
            if (srow_ok && nv > 0) {
              float* o = a.out + static_cast<long long>(srow) * a.ld_out + col_base + c0;
              if (bias == nullptr && nv >= 32) {
#pragma unroll
                for (int c = 0; c < 32; ++c) o[c] = __uint_as_float(v[c]) * a.scale;
              } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                  if (c < nv) {
                    float s = __uint_as_float(v[c]) * a.scale;
                    if (bias != nullptr) s += __ldg(bias + col_base + c0 + c);
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
              if (C::DT == DT_BF16) {
                // packed arithmetic (FFMA2 / FADD2); every third pair takes its exponentials from the FMA pipe:
                // this epilogue is MUFU-bound (one ex2 per scored pair, 72 % XU utilisation measured)
                const uint64_t c22 = pack2(c2, c2), nm2 = pack2(nm, nm);
                uint64_t acc2[2] = {0ull, 0ull};
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const uint64_t x2 = ffma2(pack2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), c22, nm2);
                  float e0, e1;
#ifndef LSE_POLY_PICK
#define LSE_POLY_PICK(i) ((i) % 3 == 2)
#endif
                  if (LSE_POLY_PICK(i)) {
                    ex2_poly2(x2, e0, e1);
                  } else {
                    float x0, x1;
                    unpack2(x2, x0, x1);
                    e0 = ex2_approx(x0);
                    e1 = ex2_approx(x1);
                  }
                  acc2[i & 1] = fadd2(acc2[i & 1], pack2(e0, e1));
                }
                float s0, s1, s2, s3;
                unpack2(acc2[0], s0, s1);
                unpack2(acc2[1], s2, s3);
                l += (s0 + s1) + (s2 + s3);
              } else {   // fp32 parity: every exponential at MUFU accuracy
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 32; ++c) acc[c & 3] += ex2_approx(fmaf(__uint_as_float(v[c]), c2, nm));
                l += (acc[0] + acc[1]) + (acc[2] + acc[3]);
              }
              if (__any_sync(0xffffffffu, static_cast<uint32_t>(rel) < 32u)) {
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  if (c == rel) ll = __uint_as_float(v[c]) * a.scale;
              }
            } else if (nv > 0) {
              float x[32];
              float cmax = -INFINITY;
              if (bias_vec && nv >= 32) {
                const float4* bp = reinterpret_cast<const float4*>(bias + col_base + c0);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                  const float4 b = __ldg(bp + g);
                  x[4 * g + 0] = fmaf(__uint_as_float(v[4 * g + 0]), c2, b.x * LOG2E);
                  x[4 * g + 1] = fmaf(__uint_as_float(v[4 * g + 1]), c2, b.y * LOG2E);
                  x[4 * g + 2] = fmaf(__uint_as_float(v[4 * g + 2]), c2, b.z * LOG2E);
                  x[4 * g + 3] = fmaf(__uint_as_float(v[4 * g + 3]), c2, b.w * LOG2E);
                  cmax = fmaxf(cmax, fmaxf(fmaxf(x[4 * g], x[4 * g + 1]), fmaxf(x[4 * g + 2], x[4 * g + 3])));
                }
              } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                  float b2 = 0.f;
                  if (bias != nullptr) b2 = (c < nv) ? __ldg(bias + col_base + c0 + c) * LOG2E : 0.f;
                  x[c] = fmaf(__uint_as_float(v[c]), c2, b2);
                  if (c >= nv) x[c] = -INFINITY;
                  cmax = fmaxf(cmax, x[c]);
                }
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
          } else if (C::EPI == EPI_TOPK) {
            if (tile_quick) {
              tmax = fmaxf(tmax, max32(v));  // raw scores; scaled once per tile (scale > 0)
            } else if (bias_vec && nv >= 32) {  // bias head, full chunk: the same logits as logit_of, biases in vectors
              const float4* bp = reinterpret_cast<const float4*>(bias + col_base + c0);
              float cm = -INFINITY;
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 b = __ldg(bp + g);
                cm = fmaxf(cm, fmaxf(fmaxf(__fmaf_rn(__uint_as_float(v[4 * g]), a.scale, b.x), __fmaf_rn(__uint_as_float(v[4 * g + 1]), a.scale, b.y)),
                                     fmaxf(__fmaf_rn(__uint_as_float(v[4 * g + 2]), a.scale, b.z), __fmaf_rn(__uint_as_float(v[4 * g + 3]), a.scale, b.w))));
              }
              tmax = fmaxf(tmax, cm);
            } else if (nv > 0) {  // bias and/or the last, partial tile
              float cm = -INFINITY;
#pragma unroll
              for (int c = 0; c < 32; ++c)
                if (c < nv) cm = fmaxf(cm, logit_of(v[c], a.scale, bias, col_base + c0 + c));
              tmax = fmaxf(tmax, cm);
            }
          } else if (C::EPI == EPI_CAND) {
            // maxima of the eight aligned groups of 4 items; the hit test is exact (scale > 0 commutes with max)
            float m8[8];
            float mul = 1.f;   // warp-uniform
            if (tile_quick) {
              mul = a.scale;
#pragma unroll
              for (int g = 0; g < 8; ++g)
                m8[g] = fmaxf(fmax3(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2])),
                              __uint_as_float(v[4 * g + 3]));
            } else if (bias_vec && nv >= 32) {  // bias head, full chunk: the same logits as logit_of, biases in vectors
              const float4* bp = reinterpret_cast<const float4*>(bias + col_base + c0);
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 b = __ldg(bp + g);
                m8[g] = fmaxf(fmaxf(__fmaf_rn(__uint_as_float(v[4 * g]), a.scale, b.x), __fmaf_rn(__uint_as_float(v[4 * g + 1]), a.scale, b.y)),
                              fmaxf(__fmaf_rn(__uint_as_float(v[4 * g + 2]), a.scale, b.z), __fmaf_rn(__uint_as_float(v[4 * g + 3]), a.scale, b.w)));
              }
            } else {  // bias and/or the last, partial tile: finished logits
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                m8[g] = -INFINITY;
#pragma unroll
                for (int c = 4 * g; c < 4 * g + 4; ++c)
                  if (c < nv) m8[g] = fmaxf(m8[g], logit_of(v[c], a.scale, bias, col_base + c0 + c));
              }
            }
            const float mm = fmax3(fmax3(m8[0], m8[1], m8[2]), fmax3(m8[3], m8[4], m8[5]), fmaxf(m8[6], m8[7]));
            // A chunk that reaches the threshold (about once per tile and WARP, 0.05 times per tile and thread) is only
            // PARKED here -- its eight group maxima move to spare registers under a predicate, no branch -- and served
            // after the S buffer has gone back to the MMA warp: the release waits for the slowest of the warpgroup's
            // 128 threads, so any divergent work in front of it is paid by the tensor pipe on nearly every tile
            // (measured: 0.68 ms without candidates, 0.97-1.26 ms with the hit path in front of the release).
            const bool hit = __fmul_rn(mm, mul) >= tau;
            if (hit && pk_pend) serve_parked();   // a second hit chunk of the same tile in the same thread: rare
            if (hit) {
#pragma unroll
              for (int g = 0; g < 8; ++g) pk_m[g] = m8[g];
              pk_item0 = static_cast<unsigned int>(col_base + c0);
              pk_mul = mul;
              pk_pend = true;
            }
          }
        }  // chunks

        // release the S buffer (all tcgen05.ld of this thread have completed)
        tc_fence_before();
        mbar_arrive(&bar->s_empty[bidx]);
        if (C::EPI == EPI_CAND) {
          if (pk_pend) serve_parked();   // off the tensor pipe's critical path: the S buffer is already released
          while (next_seen < col_base + C::BN) seen_advance();
          if (refresh) {
            // publish this thread's counts (>= counts of the two rungs above it), then climb as far as the row's
            // shared counters allow
            unsigned int hv[8] = {h_lo.x, h_lo.y, h_lo.z, h_lo.w, h_hi.x, h_hi.y, h_hi.z, h_hi.w};
            if (n1 != 0u) atomicAdd(ld->cnt + lvl + 1, n1);                    // results unused: REDs
            if (n2 != 0u && lvl + 2 <= 7) atomicAdd(ld->cnt + lvl + 2, n2);
            const int lvl0 = lvl;
#pragma unroll
            for (int l = 1; l <= TOPK_LEVELS; ++l) {
              unsigned int c = hv[l];
              if (l == lvl0 + 1) c += n1;      // the loaded value predates this thread's own publication
              if (l == lvl0 + 2) c += n2;
              if (l == lvl + 1 && c >= static_cast<unsigned int>(a.k_need)) lvl = l;
            }
            n1 = 0u; n2 = 0u;
            if (lvl != lvl0) {   // rare: at most TOPK_LEVELS times per work item
              tau = fmaxf(tau, __ldg(ld->thr + lvl));
              ck1 = __ldg(ld->thr + lvl + 1);
              ck2 = (lvl + 2 <= 7) ? __ldg(ld->thr + lvl + 2) : INFINITY;
            }
          }
        }
        if (C::EPI == EPI_TOPK) {
          float out = tile_quick ? __fmul_rn(tmax, a.scale) : tmax;
          if (tile_dirty && !mask_exact) out = __int_as_float(0x7fc00000);  // NaN
          if (srow_ok) a.tile_max[static_cast<long long>(srow) * a.n_strm_tiles + t] = out;
        }
      }  // tiles

      // ---- per-item outputs
      if (C::EPI == EPI_CAND && srow_ok) {
        a.cand_cnt[static_cast<long long>(srow) * a.n_sub + split * SUBS + subidx] = n_cand;
        if (n1 != 0u) atomicAdd(ld->cnt + lvl + 1, n1);   // leftovers still help the row's other splits
        if (n2 != 0u && lvl + 2 <= 7) atomicAdd(ld->cnt + lvl + 2, n2);
      }
      if (C::EPI == EPI_LSE && C::XT == 2 && C::NWG == 2) {  // each warpgroup owns its rows: no hand-over
        a.part_m2[pslot] = m2;
        a.part_l[pslot] = l;
        a.part_ll[pslot] = ll;
      }
      if (C::EPI == EPI_LSE && (C::XT == 1 || C::NWG == 4)) {
        // the odd-parity warpgroup hands its partial to the even-parity one of the same stationary tile, which merges
        // and writes one slot per row
        if (par == 1) { bar->xchg[xsel][0][r] = m2; bar->xchg[xsel][1][r] = l; bar->xchg[xsel][2][r] = ll; }
        epi_bar_sync(xsel);
        if (par == 0) {
          const float om = bar->xchg[xsel][0][r], ol = bar->xchg[xsel][1][r], oll = bar->xchg[xsel][2][r];
          const float mm = fmaxf(m2, om);
          float lm = 0.f;
          if (mm > -INFINITY) lm = l * ex2_approx(m2 - mm) + ol * ex2_approx(om - mm);
          a.part_m2[pslot] = mm;
          a.part_l[pslot] = lm;
          a.part_ll[pslot] = ll + oll;
        }
        epi_bar_sync(xsel);  // xchg is free again before the next item
      }
    }  // items
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

}  // namespace rb
