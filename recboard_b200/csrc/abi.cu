// C ABI of recboard_b200 (see include/recboard_b200.h for the contract and the reference lines
// each entry point replaces).  Host-side planning + kernel launches only; no allocation.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include <nvtx3/nvToolsExt.h>   // header-only: ranges cost nothing unless a profiler is attached

#include "../../include/recboard_b200.h"
#include "sweep.cuh"
#include "pair.cuh"
#include "simt.cuh"
#include "scatter.cuh"
#include "f32grad.cuh"
#include "launch.cuh"

using namespace rb;

// ------------------------------------------------------------------------------- errors
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return static_cast<int>(e);
}
#define RB_CUDA(x)                                   \
  do {                                               \
    cudaError_t e_ = (x);                            \
    if (e_ != cudaSuccess) return cuda_fail(e_, #x); \
  } while (0)
#define RB_LAUNCH_CHECK(name)                                 \
  do {                                                        \
    ++g_launches;                                             \
    cudaError_t e_ = cudaGetLastError();                      \
    if (e_ != cudaSuccess) return cuda_fail(e_, name);        \
  } while (0)

// One NVTX range per C-ABI entry (nsys / ncu timelines show the path's entry points by name).
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define RB_RANGE(name) NvtxRange rb_nvtx_range_(name)

extern "C" const char* rb_last_error(void) { return g_err.c_str(); }
extern "C" const char* rb_version(void) { return "recboard_b200 0.1 (sm_100a)"; }
extern "C" int64_t rb_launch_count(void) { return g_launches.load(); }

// ----------------------------------------------------------------------------- device
struct DevInfo { int ok = 0; int sms = 0; int cc = 0; };
static int get_dev(DevInfo& d) {
  static thread_local DevInfo cache[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(RB_E_NODEVICE, "no CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(e)); }
  if (dev < 0 || dev >= 64) return fail(RB_E_NODEVICE, "device ordinal %d out of range", dev);
  if (!cache[dev].ok) {
    cudaDeviceProp p;
    RB_CUDA(cudaGetDeviceProperties(&p, dev));
    cache[dev].sms = p.multiProcessorCount;
    cache[dev].cc = p.major * 10 + p.minor;
    cache[dev].ok = 1;
  }
  d = cache[dev];
  if (d.cc != 100) return fail(RB_E_NODEVICE, "recboard_b200 kernels are sm_100a only; device is sm_%d", d.cc);
  return 0;
}

// -------------------------------------------------------------------------- TMA maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// Row-major (rows, cols) matrix -> 2-D map with a (box_rows x 128-byte) box, 128-byte swizzle.
static int make_tmap(CUtensorMap* m, const void* base, bool bf16, long long rows, long long cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(RB_E_NODEVICE, "cuTensorMapEncodeTiled not available from the driver");
  const int es = bf16 ? 2 : 4;
  if (reinterpret_cast<uintptr_t>(base) & 15) return fail(RB_E_ALIGN, "operand base %p is not 16-byte aligned", base);
  if ((cols * es) % 16) return fail(RB_E_ALIGN, "row pitch %lld bytes is not a multiple of 16", cols * es);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(cols * es)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / es), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(RB_E_ARG, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld box_rows=%d)", (int)r, rows, cols, box_rows);
  return 0;
}

// --------------------------------------------------------------------------- planning
using rb::BN;   // streamed tile rows (launch.cuh)

struct Plan { int n_stat_tiles, n_strm_tiles, n_splits, grid; };

// Work item = (stationary tile, split of the streamed range).  Pick the smallest split count that
// fills the SMs in whole waves (>= 97 %), keeping at least `min_tiles` streamed tiles per item.
static Plan make_plan(long long n_stat, long long n_strm, int sms, int max_splits_cap, int min_tiles = 8,
                      int stat_rows = 128) {
  Plan p;
  p.n_stat_tiles = static_cast<int>((n_stat + stat_rows - 1) / stat_rows);
  p.n_strm_tiles = static_cast<int>((n_strm + BN - 1) / BN);
  int max_s = std::max(1, std::min(p.n_strm_tiles / min_tiles, max_splits_cap));
  max_s = std::min(max_s, 8 * sms);
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= max_s; ++s) {
    const long long items = 1ll * p.n_stat_tiles * s;
    const long long waves = (items + sms - 1) / sms;
    const double eff = static_cast<double>(items) / static_cast<double>(waves * sms);
    if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    if (best_eff >= 0.97) break;
  }
  p.n_splits = best;
  p.grid = static_cast<int>(std::min<long long>(1ll * p.n_stat_tiles * p.n_splits, sms));
  return p;
}

struct Bump {
  char* base; size_t cap; size_t off = 0;
  Bump(void* b, size_t c) : base(static_cast<char*>(b)), cap(c) {}
  template <typename T> T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
  bool ok() const { return off <= cap; }
};

static int kc_for(int d, int mode) {
  const int per = (mode == RB_MODE_BF16) ? 64 : 32;
  const int kc = (d + per - 1) / per;
  return (mode != RB_MODE_BF16 && kc == 3) ? 4 : kc;   // fp32x3 tiles come in 1, 2 or 4 chunks ([hi|lo] pitch = kc*32)
}
static int check_common(const void* U, const void* W, long long M, long long N, int d, int dtype, int mode) {
  if (!U || !W) return fail(RB_E_ARG, "null operand");
  if (M <= 0 || N <= 0 || d <= 0) return fail(RB_E_ARG, "bad shape M=%lld N=%lld d=%d", M, N, d);
  if (M >= (1ll << 31) - 256 || N >= (1ll << 31) - 256) return fail(RB_E_ARG, "M/N must be < 2^31");
  if (d % 8) return fail(RB_E_ARG, "d=%d must be a multiple of 8", d);
  if (mode == RB_MODE_BF16) {
    if (dtype != RB_DTYPE_BF16) return fail(RB_E_UNSUPPORTED, "bf16 mode needs bf16 operands (cast on the host side)");
    if (d > 256) return fail(RB_E_UNSUPPORTED, "bf16 mode supports d <= 256 (got %d)", d);
  } else if (mode == RB_MODE_FP32X3) {
    if (dtype != RB_DTYPE_F32) return fail(RB_E_UNSUPPORTED, "fp32x3 mode needs fp32 operands");
    if (d > 128) return fail(RB_E_UNSUPPORTED, "fp32x3 mode supports d <= 128 (got %d)", d);
  } else {
    return fail(RB_E_ARG, "unknown mode %d", mode);
  }
  return 0;
}

// ------------------------------------------------------------------- kernel dispatch
// The tcgen05 kernel templates are instantiated in their own translation units (inst_*.cu, one per epilogue /
// pass, compiled in parallel by build.py); this file keeps the host-side planning and the SIMT kernels.
int rb::host_fail(int code, const char* msg) { return fail(code, "%s", msg); }
int rb::host_cuda_fail(cudaError_t e, const char* what) { return cuda_fail(e, what); }
void rb::count_launch() { ++g_launches; }

// Stationary tiles per CTA: two (256 rows) whenever there is a second tile to fill -- halves the L2->SM
// traffic of a sweep (see sweep.cuh).
static int sweep_xt(int mode, int d, long long n_stat) {
  if (mode != RB_MODE_BF16 && d > 64) return 1;   // fp32x3 at d = 128: one [hi|lo] stationary tile is already 128 KB
  return n_stat > 128 ? 2 : 1;
}
static bool pair_ok(int mode, int d, float scale) { return mode == RB_MODE_BF16 && d <= 256 && scale > 0.f; }
// stationary rows per work item of the fused CE passes: two 128-row tiles, or one for 128 < d <= 256 (pair.cuh, DS)
static int pair_rows(int d) { return d > 128 ? 128 : 256; }

// Operand staging: bf16 operands are used in place; fp32x3 operands are split into [hi|lo] in ws.
struct Operand { const void* ptr; long long cols; bool bf16; };
static int stage_operand(const void* X, long long rows, int d, int mode, Bump& ws, Operand& o, cudaStream_t st) {
  if (mode == RB_MODE_BF16) { o = {X, d, true}; return 0; }
  const int dpad = kc_for(d, mode) * 32;
  float* hl = ws.take<float>(static_cast<size_t>(rows) * 2 * dpad);
  if (!ws.ok()) return fail(RB_E_WORKSPACE, "workspace too small (need > %zu bytes)", ws.off);
  const long long total = rows * dpad;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
  split_hi_lo_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(X), hl, rows, d, dpad);
  RB_LAUNCH_CHECK("split_hi_lo_kernel");
  o = {hl, 2ll * dpad, false};
  return 0;
}
static size_t staged_bytes(long long rows, int d, int mode) {
  if (mode == RB_MODE_BF16) return 0;
  return static_cast<size_t>(rows) * 2 * kc_for(d, mode) * 32 * 4 + 256;
}

// ============================================================================ gather
extern "C" int rb_gather_rows(const void* table, const int64_t* idx, void* out, int64_t n_idx, int64_t n_rows,
                              int d, int dtype, rb_stream_t stream) {
  RB_RANGE("rb_gather_rows");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  if (!table || !idx || !out) return fail(RB_E_ARG, "null pointer");
  if (n_idx < 0 || n_rows <= 0 || d <= 0) return fail(RB_E_ARG, "bad shape");
  const int es = dtype == RB_DTYPE_BF16 ? 2 : 4;
  const long long row_bytes = 1ll * d * es;
  if (row_bytes % 16) return fail(RB_E_ALIGN, "row of %lld bytes is not a multiple of 16", row_bytes);
  if ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out)) & 15) return fail(RB_E_ALIGN, "table/out must be 16-byte aligned");
  if (n_idx == 0) return 0;
  const int vpr = static_cast<int>(row_bytes / 16);
  const long long total = n_idx * vpr;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 1ll * dv.sms * 8));
  gather_rows_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(table), idx, static_cast<uint4*>(out), n_idx, n_rows, vpr);
  RB_LAUNCH_CHECK("gather_rows_kernel");
  return 0;
}

// ------------------------------------------------------- gather over peer-mapped shards (one box, NVLink)
// rb_ipc_export / rb_ipc_open: the CUDA IPC plumbing a process needs to map another rank's shard (the owner exports the
// handle of the allocation that holds `ptr` and the offset of `ptr` inside it; a peer opens it with ITS device
// current, which also enables peer access between the two devices).
extern "C" int rb_ipc_export(const void* ptr, void* handle64, int64_t* offset) {
  RB_RANGE("rb_ipc_export");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  if (!ptr || !handle64 || !offset) return fail(RB_E_ARG, "null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CUdeviceptr base = 0; size_t size = 0;
  typedef CUresult (*range_fn_t)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fp)
    return fail(RB_E_NODEVICE, "cuMemGetAddressRange not available from the driver");
  CUresult cr = reinterpret_cast<range_fn_t>(fp)(&base, &size, reinterpret_cast<CUdeviceptr>(ptr));
  if (cr != CUDA_SUCCESS) return fail(RB_E_ARG, "cuMemGetAddressRange failed with CUresult %d", (int)cr);
  cudaIpcMemHandle_t h;
  RB_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  std::memcpy(handle64, &h, 64);
  *offset = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(ptr) - base);
  return 0;
}
extern "C" int rb_ipc_open(const void* handle64, int64_t offset, void** ptr_out, void** base_out) {
  RB_RANGE("rb_ipc_open");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  if (!handle64 || !ptr_out || !base_out) return fail(RB_E_ARG, "null pointer");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  void* base = nullptr;
  RB_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *base_out = base;
  *ptr_out = static_cast<char*>(base) + offset;
  return 0;
}
extern "C" int rb_ipc_close(void* base) {
  if (!base) return 0;
  RB_CUDA(cudaIpcCloseMemHandle(base));
  return 0;
}
extern "C" int rb_gather_rows_peers(const void* const* shard_ptrs, const int64_t* starts, int n_shards, const int64_t* idx,
                                    void* out, int64_t n_idx, int d, int dtype, rb_stream_t stream) {
  RB_RANGE("rb_gather_rows_peers");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  if (!shard_ptrs || !starts || !idx || !out) return fail(RB_E_ARG, "null pointer");
  if (n_shards < 1 || n_shards > PEER_MAX_SHARDS) return fail(RB_E_ARG, "1 <= n_shards <= %d (got %d)", PEER_MAX_SHARDS, n_shards);
  if (n_idx < 0 || d <= 0) return fail(RB_E_ARG, "bad shape");
  const int es = dtype == RB_DTYPE_BF16 ? 2 : 4;
  const long long row_bytes = 1ll * d * es;
  if (row_bytes % 16) return fail(RB_E_ALIGN, "row of %lld bytes is not a multiple of 16", row_bytes);
  PeerShards sh{};
  sh.n = n_shards;
  for (int r = 0; r < n_shards; ++r) {
    if (!shard_ptrs[r] || (reinterpret_cast<uintptr_t>(shard_ptrs[r]) & 15)) return fail(RB_E_ALIGN, "shard %d: null or not 16-byte aligned", r);
    if (starts[r + 1] < starts[r]) return fail(RB_E_ARG, "shard starts must be non-decreasing");
    sh.base[r] = static_cast<const uint4*>(shard_ptrs[r]);
    sh.start[r] = starts[r];
  }
  sh.start[n_shards] = starts[n_shards];
  if (reinterpret_cast<uintptr_t>(out) & 15) return fail(RB_E_ALIGN, "out must be 16-byte aligned");
  if (n_idx == 0) return 0;
  const int vpr = static_cast<int>(row_bytes / 16);
  const long long total = n_idx * vpr;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 1ll * dv.sms * 8));
  gather_rows_peers_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sh, idx, static_cast<uint4*>(out), n_idx, vpr);
  RB_LAUNCH_CHECK("gather_rows_peers_kernel");
  return 0;
}

// ================================================================== row compaction
extern "C" int rb_compact_index(const void* mask, int64_t n, int64_t* row_index, int32_t* count, rb_stream_t stream) {
  RB_RANGE("rb_compact_index");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  if (!mask || !row_index || !count) return fail(RB_E_ARG, "null pointer");
  if (n < 0 || n >= (1ll << 31)) return fail(RB_E_ARG, "bad length %lld", (long long)n);
  compact_index_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(static_cast<const unsigned char*>(mask), n, row_index, count);
  RB_LAUNCH_CHECK("compact_index_kernel");
  return 0;
}

// ========================================================================= normalise
extern "C" int rb_normalize_rows(const void* x, void* out, float* inv_norm, int64_t n_rows, int d, int in_dtype,
                                 int out_dtype, float eps, rb_stream_t stream) {
  RB_RANGE("rb_normalize_rows");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!x || !out) return fail(RB_E_ARG, "null pointer");
  if (n_rows < 0 || d <= 0 || d % 8 || d > 1024) return fail(RB_E_ARG, "bad shape: d=%d must be a multiple of 8, <= 1024", d);
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) return fail(RB_E_ALIGN, "x/out must be 16-byte aligned");
  if (n_rows == 0) return 0;
  const int grid = static_cast<int>((n_rows * 32 + 255) / 256);
  const bool ib = in_dtype == RB_DTYPE_BF16, ob = out_dtype == RB_DTYPE_BF16;
  if ((!ib && in_dtype != RB_DTYPE_F32) || (!ob && out_dtype != RB_DTYPE_F32)) return fail(RB_E_ARG, "unknown dtype");
  if (ib && ob) normalize_rows_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), inv_norm, n_rows, d, eps);
  else if (ib) normalize_rows_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<float*>(out), inv_norm, n_rows, d, eps);
  else if (ob) normalize_rows_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const float*>(x), static_cast<__nv_bfloat16*>(out), inv_norm, n_rows, d, eps);
  else normalize_rows_kernel<float, float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), static_cast<float*>(out), inv_norm, n_rows, d, eps);
  RB_LAUNCH_CHECK("normalize_rows_kernel");
  return 0;
}

// ======================================================================= scatter-add
static int scatter_passes(long long n_rows) {
  int bits = 1;
  while ((1ull << bits) <= static_cast<unsigned long long>(n_rows)) ++bits;  // keys in [0, n_rows]
  return (bits + RS_BITS - 1) / RS_BITS;
}
static size_t scatter_ws_bytes(long long n_idx, int d) {
  const long long chunks = (n_idx + RS_CHUNK - 1) / RS_CHUNK;
  const long long blocks = (n_idx + SEG_BLOCK - 1) / SEG_BLOCK;
  return static_cast<size_t>(n_idx) * 4 * 4 + static_cast<size_t>(chunks) * RS_BINS * 4 + 3 * RS_BINS * 4 +
         2 * static_cast<size_t>(blocks) * d * 4 + 10 * 256;
}
template <typename T, typename TG>
static int scatter_launch(const ScatterArgs& a, int sms, cudaStream_t st) {
  auto kern = scatter_add_coop_kernel<T, TG>;
  RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM_BYTES));
  int per_sm = 0;
  RB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SC_THREADS, SC_SMEM_BYTES));
  if (per_sm < 1) return fail(RB_E_UNSUPPORTED, "scatter_add_coop_kernel does not fit on this device");
  // enough co-resident blocks for the widest phase (one warp per 32 sorted positions), never more than fit at once
  const long long want = (static_cast<long long>(a.n) + SEG_BLOCK * SC_WARPS - 1) / (SEG_BLOCK * SC_WARPS);
  int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(want, 1ll * per_sm * sms)));
  if (const char* e = getenv("RB_SCATTER_GRID")) grid = std::max(1, std::min(atoi(e), per_sm * sms));   // tuning hook
  void* params[] = {const_cast<ScatterArgs*>(&a)};
  RB_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(grid), dim3(SC_THREADS), params, SC_SMEM_BYTES, st));
  RB_LAUNCH_CHECK("scatter_add_coop_kernel");
  return 0;
}
// dst[idx[i]-idx_base] += alpha*(*alpha_dev) * ew[i] * src[i / row_div] in a fixed order; optionally
// cnt_out[row] += cnt_alpha*(*alpha_dev) * (#occurrences of row).  dst is fp32 or bf16 (dst_dtype).
static int scatter_add_impl(const void* grad_out, const int64_t* idx, long long idx_base, void* grad_table,
                            int64_t n_idx, int64_t n_rows, int d, int dtype, int64_t padding_idx, float alpha,
                            const float* alpha_dev, float* cnt_out, float cnt_alpha, void* ws, size_t ws_bytes,
                            cudaStream_t st, int row_div = 1, const float* ew = nullptr, int dst_dtype = RB_DTYPE_F32) {
  if (!grad_out || !idx || !grad_table) return fail(RB_E_ARG, "null pointer");
  if (n_idx < 0 || n_rows <= 0 || d <= 0 || d % 4) return fail(RB_E_ARG, "bad shape (d must be a multiple of 4)");
  if (n_idx >= (1ll << 31) || n_rows >= (1ll << 32) - 1) return fail(RB_E_ARG, "n_idx/n_rows too large");
  if (n_idx == 0) return 0;
  if (!ws || ws_bytes < scatter_ws_bytes(n_idx, d)) return fail(RB_E_WORKSPACE, "workspace too small: need %zu bytes", scatter_ws_bytes(n_idx, d));
  DevInfo dv; if (int r = get_dev(dv)) return r;
  Bump b(ws, ws_bytes);
  const int n = static_cast<int>(n_idx);
  const int chunks = (n + RS_CHUNK - 1) / RS_CHUNK;
  const int seg_blocks = (n + SEG_BLOCK - 1) / SEG_BLOCK;
  ScatterArgs a{};
  a.src = grad_out; a.idx = idx; a.idx_base = idx_base; a.dst = grad_table; a.n = n; a.n_rows = n_rows; a.d = d;
  a.padding_idx = padding_idx; a.alpha = alpha; a.alpha_dev = alpha_dev; a.cnt_out = cnt_out; a.cnt_alpha = cnt_alpha;
  a.row_div = row_div; a.ew = ew; a.passes = scatter_passes(n_rows);
  a.k0 = b.take<uint32_t>(n); a.v0 = b.take<uint32_t>(n); a.k1 = b.take<uint32_t>(n); a.v1 = b.take<uint32_t>(n);
  a.hist = b.take<uint32_t>(static_cast<size_t>(chunks) * RS_BINS);
  a.totals = b.take<uint32_t>(3 * RS_BINS);
  a.lead = b.take<float>(static_cast<size_t>(seg_blocks) * d);
  a.trail = b.take<float>(static_cast<size_t>(seg_blocks) * d);
  if (!b.ok() || a.passes > 3) return fail(RB_E_WORKSPACE, "scatter workspace layout");
  RB_CUDA(cudaMemsetAsync(a.totals, 0, static_cast<size_t>(a.passes) * RS_BINS * 4, st));
  const bool sb = dtype == RB_DTYPE_BF16, db = dst_dtype == RB_DTYPE_BF16;
  if (sb && db) return scatter_launch<__nv_bfloat16, __nv_bfloat16>(a, dv.sms, st);
  if (sb) return scatter_launch<__nv_bfloat16, float>(a, dv.sms, st);
  if (db) return scatter_launch<float, __nv_bfloat16>(a, dv.sms, st);
  return scatter_launch<float, float>(a, dv.sms, st);
}
extern "C" int rb_scatter_add_rows(const void* grad_out, const int64_t* idx, float* grad_table, int64_t n_idx,
                                   int64_t n_rows, int d, int dtype, int64_t padding_idx, void* ws, size_t ws_bytes,
                                   rb_stream_t stream) {
  RB_RANGE("rb_scatter_add_rows");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  return scatter_add_impl(grad_out, idx, 0, grad_table, n_idx, n_rows, d, dtype, padding_idx, 1.f, nullptr, nullptr, 0.f,
                          ws, ws_bytes, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int rb_scatter_add_rows_into(const void* grad_out, const int64_t* idx, void* grad_table, int64_t n_idx,
                                        int64_t n_rows, int d, int dtype, int table_dtype, int64_t padding_idx, void* ws,
                                        size_t ws_bytes, rb_stream_t stream) {
  RB_RANGE("rb_scatter_add_rows_into");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  if (table_dtype != RB_DTYPE_F32 && table_dtype != RB_DTYPE_BF16) return fail(RB_E_ARG, "unknown table dtype %d", table_dtype);
  if (table_dtype == RB_DTYPE_BF16 && (d % 4 || (reinterpret_cast<uintptr_t>(grad_table) & 7))) return fail(RB_E_ALIGN, "bf16 table rows must be 8-byte aligned");
  return scatter_add_impl(grad_out, idx, 0, grad_table, n_idx, n_rows, d, dtype, padding_idx, 1.f, nullptr, nullptr, 0.f,
                          ws, ws_bytes, reinterpret_cast<cudaStream_t>(stream), 1, nullptr, table_dtype);
}

// ========================================================================= gather-dot
static int gather_dot_gl(int d, int dtype) {
  const int vpr = d / (dtype == RB_DTYPE_BF16 ? 8 : 4);
  int gl = 1;
  while (gl < vpr && gl < 32) gl <<= 1;
  return gl;
}
static int gather_dot_check(const void* U, const void* table, const int64_t* idx, int64_t M, int64_t K, int64_t n_rows,
                            int d, int dtype) {
  if (!U || !table || !idx) return fail(RB_E_ARG, "null pointer");
  if (M < 0 || K <= 0 || K >= (1ll << 31) || n_rows <= 0 || d <= 0) return fail(RB_E_ARG, "bad shape M=%lld K=%lld d=%d", (long long)M, (long long)K, d);
  if (dtype != RB_DTYPE_BF16 && dtype != RB_DTYPE_F32) return fail(RB_E_ARG, "unknown dtype %d", dtype);
  if (d % (dtype == RB_DTYPE_BF16 ? 8 : 4)) return fail(RB_E_ALIGN, "rows must be a multiple of 16 bytes");
  if ((reinterpret_cast<uintptr_t>(U) | reinterpret_cast<uintptr_t>(table)) & 15) return fail(RB_E_ALIGN, "U/table must be 16-byte aligned");
  return 0;
}
extern "C" int rb_gather_dot(const void* U, const void* table, const int64_t* idx, float scale, float* S, int64_t M,
                             int64_t K, int64_t n_rows, int d, int dtype, rb_stream_t stream) {
  RB_RANGE("rb_gather_dot");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (int r = gather_dot_check(U, table, idx, M, K, n_rows, d, dtype)) return r;
  if (!S) return fail(RB_E_ARG, "null output");
  if (M == 0) return 0;
  const int grid = static_cast<int>((M * 32 + 255) / 256), gl = gather_dot_gl(d, dtype);
  if (dtype == RB_DTYPE_BF16)
    gather_dot_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(U), static_cast<const __nv_bfloat16*>(table), idx, scale, S, M, (int)K, n_rows, d, gl);
  else
    gather_dot_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(U), static_cast<const float*>(table), idx, scale, S, M, (int)K, n_rows, d, gl);
  RB_LAUNCH_CHECK("gather_dot_kernel");
  return 0;
}
extern "C" int rb_gather_dot_bwd(const void* U, const void* table, const int64_t* idx, const float* G, float scale,
                                 float* dU, float* dTable, int64_t M, int64_t K, int64_t n_rows, int d, int dtype,
                                 int64_t padding_idx, void* ws, size_t ws_bytes, rb_stream_t stream) {
  RB_RANGE("rb_gather_dot_bwd");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (int r = gather_dot_check(U, table, idx, M, K, n_rows, d, dtype)) return r;
  if (!G) return fail(RB_E_ARG, "null upstream gradient");
  if (M == 0) return 0;
  if (dU) {
    const int grid = static_cast<int>((M * 32 + 255) / 256), gl = gather_dot_gl(d, dtype);
    if (dtype == RB_DTYPE_BF16)
      gather_dot_du_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(table), idx, G, scale, dU, M, (int)K, n_rows, d, gl);
    else
      gather_dot_du_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(table), idx, G, scale, dU, M, (int)K, n_rows, d, gl);
    RB_LAUNCH_CHECK("gather_dot_du_kernel");
  }
  if (dTable)  // dTable[idx[m,k]] += scale * G[m,k] * U[m]: the sorted, deterministic scatter with per-entry weights
    return scatter_add_impl(U, idx, 0, dTable, M * K, n_rows, d, dtype, padding_idx, scale, nullptr, nullptr, 0.f, ws, ws_bytes,
                            st, static_cast<int>(K), G);
  return 0;
}

// =============================================================================== SpMM
extern "C" int rb_spmm_csr(const int64_t* crow, const int64_t* col, const float* val, const float* X, float* Y,
                           float* acc, float beta, int64_t n_rows, int64_t n_cols, int d, rb_stream_t stream) {
  RB_RANGE("rb_spmm_csr");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!crow || !col || !val || !X || (!Y && !acc)) return fail(RB_E_ARG, "null pointer");
  if (n_rows < 0 || n_cols <= 0 || d <= 0 || d % 4) return fail(RB_E_ARG, "bad shape (d must be a multiple of 4)");
  if ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(acc)) & 15) return fail(RB_E_ALIGN, "X/Y/acc must be 16-byte aligned");
  if (n_rows == 0) return 0;
  int gl = 1;
  while (gl < d / 4 && gl < 32) gl <<= 1;
  const long long warps = (n_rows + (32 / gl) - 1) / (32 / gl);
  spmm_csr_kernel<<<static_cast<int>((warps * 32 + 255) / 256), 256, 0, st>>>(crow, col, val, X, Y, acc, beta, n_rows, n_cols, d, gl);
  RB_LAUNCH_CHECK("spmm_csr_kernel");
  return 0;
}

// ======================================================================= score dense
extern "C" int rb_score_dense(const void* U, const void* W, const float* bias, float scale, float* S, int64_t M,
                              int64_t N, int d, int dtype, int mode, void* ws, size_t ws_bytes, rb_stream_t stream) {
  RB_RANGE("rb_score_dense");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (int r = check_common(U, W, M, N, d, dtype, mode)) return r;
  if (!S) return fail(RB_E_ARG, "null output");
  Bump b(ws, ws_bytes);
  Operand ou, ow;
  if (int r = stage_operand(U, M, d, mode, b, ou, st)) return r;
  if (int r = stage_operand(W, N, d, mode, b, ow, st)) return r;
  Plan p = make_plan(M, N, dv.sms, 1 << 20);
  CUtensorMap ts, ty;
  if (int r = make_tmap(&ts, ou.ptr, ou.bf16, M, ou.cols, 128)) return r;
  if (int r = make_tmap(&ty, ow.ptr, ow.bf16, N, ow.cols, BN)) return r;
  SweepArgs a{};
  a.n_stat = (int)M; a.n_strm = (int)N; a.n_stat_tiles = p.n_stat_tiles; a.n_strm_tiles = p.n_strm_tiles;
  a.n_splits = p.n_splits; a.d = d; a.scale = scale; a.bias = bias; a.out = S; a.ld_out = N;
  return launch_sweep_dense(mode, kc_for(d, mode), ts, ty, a, p.grid, st);
}

static size_t pair_fwd_ws(long long M, long long N, int d, int sms) {
  Plan p = make_plan(M, N, sms, 1 << 20, 8, pair_rows(d));
  const size_t stat_pad = static_cast<size_t>(p.n_stat_tiles) * pair_rows(d);
  return 2 * stat_pad * p.n_splits * 4 + static_cast<size_t>(p.n_splits) * M * d * 4 +
         static_cast<size_t>(p.n_strm_tiles) * 128 * 4 + 2048;
}

// ============================================================================ CE fwd
// Fused pass (pair kernel, PASS_FWD) when dU_unnorm is requested; statistics-only sweep otherwise.
static int ce_fwd_pair(const DevInfo& dv, const void* U, const void* W, const float* bias, float scale,
                       const int64_t* labels, int64_t label_base, int64_t M, int64_t N, int d, float* row_max,
                       float* row_sumexp, float* label_logit, float* dU_unnorm, Bump& b, cudaStream_t st,
                       const int* m_dev = nullptr) {
  Plan p = make_plan(M, N, dv.sms, 1 << 20, 8, pair_rows(d));
  const long long stat_pad = 1ll * p.n_stat_tiles * pair_rows(d);
  const long long n_pad = 1ll * p.n_strm_tiles * 128;
  float* pm2 = b.take<float>(stat_pad * p.n_splits);
  float* pl = b.take<float>(stat_pad * p.n_splits);
  float* pacc = b.take<float>(static_cast<size_t>(p.n_splits) * M * d);
  float* bias2 = bias ? b.take<float>(n_pad) : nullptr;
  if (!b.ok()) return fail(RB_E_WORKSPACE, "workspace too small: need %zu bytes", b.off);
  if (bias) {
    bias2_kernel<<<(int)((n_pad + 255) / 256), 256, 0, st>>>(bias, bias2, N, n_pad);
    RB_LAUNCH_CHECK("bias2_kernel");
  }
  CUtensorMap ts, ty;
  if (int r = make_tmap(&ts, U, true, M, d, 128)) return r;
  if (int r = make_tmap(&ty, W, true, N, d, 128)) return r;
  PairArgs a{};
  a.n_stat = (int)M; a.n_strm = (int)N; a.n_pair_tiles = p.n_stat_tiles; a.n_strm_tiles = p.n_strm_tiles;
  a.n_splits = p.n_splits; a.d = d; a.stat_pad = (int)stat_pad; a.scale = scale; a.aux = bias2;
  a.part_m2 = pm2; a.part_l = pl; a.acc_out = pacc; a.m_dev = m_dev;
  if (int r = launch_pair_fwd(kc_for(d, RB_MODE_BF16), bias != nullptr, ts, ty, a, p.grid, st)) return r;
  const int grid = (int)((M * 32 + 255) / 256);
  ce_fwd_finish_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
      pm2, pl, pacc, p.n_splits, stat_pad, (int)M, d, static_cast<const __nv_bfloat16*>(U),
      static_cast<const __nv_bfloat16*>(W), bias, labels, label_base, N, scale, row_max, row_sumexp, label_logit, dU_unnorm, m_dev);
  RB_LAUNCH_CHECK("ce_fwd_finish_kernel");
  return 0;
}

extern "C" int rb_ce_fwd(const void* U, const void* W, const float* bias, float scale, const int64_t* labels,
                         int64_t label_base, int64_t M, int64_t N, int d, int dtype, int mode, float* row_max,
                         float* row_sumexp, float* label_logit, float* dU_unnorm, const int32_t* m_dev, void* ws,
                         size_t ws_bytes, rb_stream_t stream) {
  RB_RANGE("rb_ce_fwd");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (int r = check_common(U, W, M, N, d, dtype, mode)) return r;
  if (!labels || !row_max || !row_sumexp || !label_logit) return fail(RB_E_ARG, "null pointer");
  if (!ws) return fail(RB_E_WORKSPACE, "workspace required");
  Bump b(ws, ws_bytes);
  if (dU_unnorm) {
    if (!pair_ok(mode, d, scale))
      return fail(RB_E_UNSUPPORTED, "the fused forward+dU pass needs bf16 mode, d <= 256 and scale > 0");
    return ce_fwd_pair(dv, U, W, bias, scale, labels, label_base, M, N, d, row_max, row_sumexp, label_logit, dU_unnorm, b, st, m_dev);
  }
  Operand ou, ow;
  if (int r = stage_operand(U, M, d, mode, b, ou, st)) return r;
  if (int r = stage_operand(W, N, d, mode, b, ow, st)) return r;
  const int xt = sweep_xt(mode, d, M);
  Plan p = make_plan(M, N, dv.sms, 1 << 20, 8, 128 * xt);
  const long long m_pad = 1ll * p.n_stat_tiles * 128 * xt;
  int* lab32 = b.take<int>(m_pad);
  float* pm2 = b.take<float>(m_pad * p.n_splits);
  float* pl = b.take<float>(m_pad * p.n_splits);
  float* pll = b.take<float>(m_pad * p.n_splits);
  if (!b.ok()) return fail(RB_E_WORKSPACE, "workspace too small: need %zu bytes", b.off);
  labels_local_kernel<<<(int)((m_pad + 255) / 256), 256, 0, st>>>(labels, label_base, N, lab32, (int)M, (int)m_pad);
  RB_LAUNCH_CHECK("labels_local_kernel");
  CUtensorMap ts, ty;
  if (int r = make_tmap(&ts, ou.ptr, ou.bf16, M, ou.cols, 128)) return r;
  if (int r = make_tmap(&ty, ow.ptr, ow.bf16, N, ow.cols, BN)) return r;
  SweepArgs a{};
  a.n_stat = (int)M; a.n_strm = (int)N; a.n_stat_tiles = p.n_stat_tiles; a.n_strm_tiles = p.n_strm_tiles;
  a.n_splits = p.n_splits; a.d = d; a.scale = scale; a.bias = bias; a.labels = lab32;
  a.part_m2 = pm2; a.part_l = pl; a.part_ll = pll; a.m_dev = m_dev;
  if (int r = launch_sweep_lse(mode, kc_for(d, mode), ts, ty, a, p.grid, st, xt)) return r;
  lse_merge_kernel<<<(int)((M + 255) / 256), 256, 0, st>>>(pm2, pl, pll, p.n_splits, m_pad, (int)M, row_max, row_sumexp, label_logit, m_dev);
  RB_LAUNCH_CHECK("lse_merge_kernel");
  return 0;
}

// dU from the forward pass's unnormalised accumulator (see ce_du_finish_kernel)
extern "C" int rb_ce_du_finish(const float* dU_unnorm, const float* row_max, const float* lse, const void* W,
                               const int64_t* labels, int64_t label_base, float scale, float grad_scale,
                               const float* grad_scale_dev, int64_t M, int64_t N, int d, int dtype, float* dU,
                               const int32_t* m_dev, rb_stream_t stream) {
  RB_RANGE("rb_ce_du_finish");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!dU_unnorm || !row_max || !lse || !W || !labels || !dU) return fail(RB_E_ARG, "null pointer");
  if (M <= 0 || N <= 0 || d <= 0 || d % 8) return fail(RB_E_ARG, "bad shape M=%lld N=%lld d=%d", (long long)M, (long long)N, d);
  const int grid = (int)((M * 32 + 255) / 256);
  const float gs = grad_scale * scale;
  if (dtype == RB_DTYPE_BF16)
    ce_du_finish_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(dU_unnorm, row_max, lse, static_cast<const __nv_bfloat16*>(W), labels, label_base, N, gs, grad_scale_dev, (int)M, d, dU, m_dev);
  else if (dtype == RB_DTYPE_F32)
    ce_du_finish_kernel<float><<<grid, 256, 0, st>>>(dU_unnorm, row_max, lse, static_cast<const float*>(W), labels, label_base, N, gs, grad_scale_dev, (int)M, d, dU, m_dev);
  else
    return fail(RB_E_ARG, "unknown dtype %d", dtype);
  RB_LAUNCH_CHECK("ce_du_finish_kernel");
  return 0;
}

// ============================================================================ CE bwd
// dW (+ dbias) through the pair kernel (PASS_DW): softmax tiles without the one-hot, then the exact
// fp32 label correction dW[label_i] -= g*scale*u_i, dbias[label_i] -= g (sorted, deterministic).
static int ce_bwd_dw_pair(const DevInfo& dv, const void* U, const void* W, const float* bias, float scale,
                          const int64_t* labels, int64_t label_base, const float* lse2, float grad_scale,
                          const float* grad_scale_dev, int64_t M, int64_t N, int d, float* dW, float* dbias, Bump& b,
                          cudaStream_t st, __nv_bfloat16* dW_bf16 = nullptr, bool accumulate = false, const int* m_dev = nullptr) {
  Plan p = make_plan(N, M, dv.sms, 64, 8, pair_rows(d));
  const long long n_pad = 1ll * p.n_stat_tiles * pair_rows(d);
  // One-hot correction: up to LABEL_FIX_MAX query rows without a sort (label_owner + label_fix), beyond that
  // through the sorted scatter.  bf16 gradient requested: with one split and the sort-free correction the
  // kernel stores bf16 rows directly (label rows also in fp32 in a side table, corrected there, then rounded);
  // otherwise the fp32 path runs in the workspace and is rounded at the end.
  constexpr long long LABEL_FIX_MAX = 16384;
  const bool small_fix = M <= LABEL_FIX_MAX;
  const bool direct_bf16 = dW_bf16 != nullptr && p.n_splits == 1 && small_fix;
  if (accumulate && !direct_bf16)
    return fail(RB_E_UNSUPPORTED, "in-place accumulation needs the direct bf16 pass (one split of the query range, M <= 16384)");
  if (dW_bf16 != nullptr && !direct_bf16) dW = b.take<float>(static_cast<size_t>(N) * d);
  int* first_of = nullptr; float* side = nullptr;
  if (small_fix) {
    first_of = b.take<int>(N);
    if (direct_bf16) side = b.take<float>(static_cast<size_t>(M) * d);
    if (!b.ok()) return fail(RB_E_WORKSPACE, "workspace too small: need %zu bytes", b.off);
    RB_CUDA(cudaMemsetAsync(first_of, 0x7F, static_cast<size_t>(N) * 4, st));
    label_owner_kernel<<<(int)((M + 255) / 256), 256, 0, st>>>(labels, label_base, N, (int)M, first_of);
    RB_LAUNCH_CHECK("label_owner_kernel");
  }
  float* part = (p.n_splits > 1) ? b.take<float>(static_cast<size_t>(p.n_splits) * N * d) : dW;
  float* rs_part = nullptr;
  if (dbias) rs_part = (p.n_splits > 1) ? b.take<float>(static_cast<size_t>(p.n_splits) * N) : dbias;
  const bool bias_cfg = bias != nullptr || dbias != nullptr;   // the BIAS variant also carries the row sums (dbias)
  float* bias2 = bias_cfg ? b.take<float>(n_pad) : nullptr;
  if (!b.ok()) return fail(RB_E_WORKSPACE, "workspace too small: need %zu bytes", b.off);
  if (bias_cfg) {
    bias2_kernel<<<(int)((n_pad + 255) / 256), 256, 0, st>>>(bias, bias2, N, n_pad);   // null bias -> zeros
    RB_LAUNCH_CHECK("bias2_kernel");
  }
  CUtensorMap ts, ty;
  if (int r = make_tmap(&ts, W, true, N, d, 128)) return r;
  if (int r = make_tmap(&ty, U, true, M, d, 128)) return r;
  PairArgs a{};
  a.n_stat = (int)N; a.n_strm = (int)M; a.n_pair_tiles = p.n_stat_tiles; a.n_strm_tiles = p.n_strm_tiles;
  a.n_splits = p.n_splits; a.d = d; a.stat_pad = (int)n_pad; a.scale = scale; a.bias2_stat = bias2; a.aux = lse2;
  a.gscale = grad_scale * scale; a.rscale = grad_scale; a.gscale_dev = grad_scale_dev; a.acc_out = part; a.rowsum_out = rs_part;
  a.m_dev = m_dev;
  if (direct_bf16) { a.out_bf16 = dW_bf16; a.slot_of_row = first_of; a.side = side; a.accumulate = accumulate ? 1 : 0; }
  if (int r = launch_pair_dw(kc_for(d, RB_MODE_BF16), bias_cfg, ts, ty, a, p.grid, st)) return r;
  if (p.n_splits > 1) {
    const long long n = N * d;
    partial_sum_kernel<<<(int)std::min<long long>((n + 255) / 256, dv.sms * 8), 256, 0, st>>>(part, p.n_splits, n, dW);
    RB_LAUNCH_CHECK("partial_sum_kernel");
    if (dbias) {
      partial_sum_kernel<<<(int)std::min<long long>((N + 255) / 256, dv.sms * 8), 256, 0, st>>>(rs_part, p.n_splits, N, dbias);
      RB_LAUNCH_CHECK("partial_sum_kernel");
    }
  }
  const __nv_bfloat16* Ub = static_cast<const __nv_bfloat16*>(U);
  const int fix_grid = (int)((M * 32 + 255) / 256);
  if (direct_bf16) {
    label_fix_kernel<__nv_bfloat16, true><<<fix_grid, 256, 0, st>>>(labels, label_base, N, (int)M, d, first_of, Ub, grad_scale * scale,
                                                                     grad_scale, grad_scale_dev, side, nullptr, dW_bf16, dbias);
    RB_LAUNCH_CHECK("label_fix_kernel");
    return 0;
  }
  if (small_fix) {
    label_fix_kernel<__nv_bfloat16, false><<<fix_grid, 256, 0, st>>>(labels, label_base, N, (int)M, d, first_of, Ub, grad_scale * scale,
                                                                      grad_scale, grad_scale_dev, nullptr, dW, nullptr, dbias);
    RB_LAUNCH_CHECK("label_fix_kernel");
  } else {
    void* sws = b.take<char>(scatter_ws_bytes(M, d));
    if (!b.ok()) return fail(RB_E_WORKSPACE, "workspace too small: need %zu bytes", b.off);
    if (int r = scatter_add_impl(U, labels, label_base, dW, M, N, d, RB_DTYPE_BF16, -1, -grad_scale * scale, grad_scale_dev,
                                 dbias, -grad_scale, sws, scatter_ws_bytes(M, d), st)) return r;
  }
  if (dW_bf16 != nullptr) {
    const long long n = N * d;
    cast_f32_bf16_kernel<<<(int)std::min<long long>((n / 4 + 255) / 256, dv.sms * 16), 256, 0, st>>>(dW, dW_bf16, n);
    RB_LAUNCH_CHECK("cast_f32_bf16_kernel");
  }
  return 0;
}

// ------------------------------------------------------------- fp32-parity CE backward
// Exact fp32 passes (f32grad.cuh): 64-row stationary tiles, the streamed range cut so that about two
// CTAs per SM are busy; split partials are summed in a fixed order.
static int f32grad_splits(long long n_stat, long long n_strm, int sms) {
  const long long tiles = (n_stat + F32G_TILE - 1) / F32G_TILE, strm_tiles = (n_strm + F32G_TILE - 1) / F32G_TILE;
  long long s = (2ll * sms + tiles - 1) / tiles;
  s = std::min<long long>(s, std::max<long long>(1, strm_tiles / 4));
  return static_cast<int>(std::max<long long>(1, std::min<long long>(s, 64)));
}
static size_t f32grad_ws(long long M, long long N, int d, int sms) {
  const size_t a = static_cast<size_t>(f32grad_splits(M, N, sms) + 1) * M * d * 4 + 1024;
  const int sw = f32grad_splits(N, M, sms);
  const size_t b = (sw > 1 ? static_cast<size_t>(sw) * N * (d + 1) * 4 : 0) + 1024;
  return a + b;
}
static int launch_f32grad(const F32GradArgs& a, cudaStream_t st) {
  const int tiles = (a.n_stat + F32G_TILE - 1) / F32G_TILE;
  const int grid = tiles * a.n_splits;
  if (a.d <= 64) {
    RB_CUDA(cudaFuncSetAttribute(ce_grad_f32_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, F32GradCfg<64>::SMEM_BYTES));
    ce_grad_f32_kernel<64><<<grid, F32G_THREADS, F32GradCfg<64>::SMEM_BYTES, st>>>(a);
  } else {
    RB_CUDA(cudaFuncSetAttribute(ce_grad_f32_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, F32GradCfg<128>::SMEM_BYTES));
    ce_grad_f32_kernel<128><<<grid, F32G_THREADS, F32GradCfg<128>::SMEM_BYTES, st>>>(a);
  }
  RB_LAUNCH_CHECK("ce_grad_f32_kernel");
  return 0;
}
static int ce_bwd_f32(const DevInfo& dv, const float* U, const float* W, const float* bias, float scale,
                      const int64_t* labels, int64_t label_base, const float* lse, float grad_scale,
                      const float* grad_scale_dev, int64_t M, int64_t N, int d, float* dU, float* dW, float* dbias,
                      Bump& b, cudaStream_t st, const int* m_dev = nullptr) {
  if (d > 128) return fail(RB_E_UNSUPPORTED, "fp32 CE backward supports d <= 128");
  const float c2 = scale * 1.4426950408889634f;
  if (dU) {  // rows stationary, items streamed: acc_i = sum_j softmax_ij w_j
    const int ns = f32grad_splits(M, N, dv.sms);
    float* acc = b.take<float>(static_cast<size_t>(M) * d);
    float* part = (ns > 1) ? b.take<float>(static_cast<size_t>(ns) * M * d) : acc;
    if (!b.ok()) return fail(RB_E_WORKSPACE, "workspace too small: need %zu bytes", b.off);
    F32GradArgs a{};
    a.X = U; a.Y = W; a.n_stat = (int)M; a.n_strm = (int)N; a.d = d; a.n_splits = ns; a.c2 = c2;
    a.stat_vec = lse; a.stat_mul = -1.f; a.strm_vec = bias; a.strm_mul = 1.f;
    a.out_scale = 1.f; a.rowsum_scale = 0.f; a.scale_dev = nullptr; a.acc_out = part; a.rowsum_out = nullptr;
    a.m_dev = m_dev; a.m_side = 0;
    if (int r = launch_f32grad(a, st)) return r;
    if (ns > 1) {
      const long long n = M * d;
      partial_sum_kernel<<<(int)std::min<long long>((n + 255) / 256, dv.sms * 8), 256, 0, st>>>(part, ns, n, acc);
      RB_LAUNCH_CHECK("partial_sum_kernel");
    }
    // dU = g*scale*(acc - w_label): the finishing kernel of the fused pass with row_max == lse (factor 1)
    ce_du_finish_kernel<float><<<(int)((M * 32 + 255) / 256), 256, 0, st>>>(acc, lse, lse, W, labels, label_base, N,
                                                                            grad_scale * scale, grad_scale_dev, (int)M, d, dU, m_dev);
    RB_LAUNCH_CHECK("ce_du_finish_kernel");
  }
  if (dW) {  // items stationary, rows streamed: dW_j = g*scale*sum_i softmax_ij u_i, dbias_j = g*sum_i softmax_ij
    const int ns = f32grad_splits(N, M, dv.sms);
    float* part = (ns > 1) ? b.take<float>(static_cast<size_t>(ns) * N * d) : dW;
    float* rs_part = nullptr;
    if (dbias) rs_part = (ns > 1) ? b.take<float>(static_cast<size_t>(ns) * N) : dbias;
    void* sws = b.take<char>(scatter_ws_bytes(M, d));
    if (!b.ok()) return fail(RB_E_WORKSPACE, "workspace too small: need %zu bytes", b.off);
    F32GradArgs a{};
    a.X = W; a.Y = U; a.n_stat = (int)N; a.n_strm = (int)M; a.d = d; a.n_splits = ns; a.c2 = c2;
    a.stat_vec = bias; a.stat_mul = 1.f; a.strm_vec = lse; a.strm_mul = -1.f;
    a.out_scale = grad_scale * scale; a.rowsum_scale = grad_scale; a.scale_dev = grad_scale_dev;
    a.acc_out = part; a.rowsum_out = rs_part; a.m_dev = m_dev; a.m_side = 1;
    if (int r = launch_f32grad(a, st)) return r;
    if (ns > 1) {
      const long long n = N * d;
      partial_sum_kernel<<<(int)std::min<long long>((n + 255) / 256, dv.sms * 8), 256, 0, st>>>(part, ns, n, dW);
      RB_LAUNCH_CHECK("partial_sum_kernel");
      if (dbias) {
        partial_sum_kernel<<<(int)std::min<long long>((N + 255) / 256, dv.sms * 8), 256, 0, st>>>(rs_part, ns, N, dbias);
        RB_LAUNCH_CHECK("partial_sum_kernel");
      }
    }
    // exact one-hot correction: dW[label_i] -= g*scale*u_i, dbias[label_i] -= g   (index order => deterministic)
    if (M <= 16384) {
      int* first_of = b.take<int>(N);
      if (!b.ok()) return fail(RB_E_WORKSPACE, "workspace too small: need %zu bytes", b.off);
      RB_CUDA(cudaMemsetAsync(first_of, 0x7F, static_cast<size_t>(N) * 4, st));
      label_owner_kernel<<<(int)((M + 255) / 256), 256, 0, st>>>(labels, label_base, N, (int)M, first_of);
      RB_LAUNCH_CHECK("label_owner_kernel");
      label_fix_kernel<float, false><<<(int)((M * 32 + 255) / 256), 256, 0, st>>>(labels, label_base, N, (int)M, d, first_of, U,
                                                                                   grad_scale * scale, grad_scale, grad_scale_dev,
                                                                                   nullptr, dW, nullptr, dbias);
      RB_LAUNCH_CHECK("label_fix_kernel");
      return 0;
    }
    return scatter_add_impl(U, labels, label_base, dW, M, N, d, RB_DTYPE_F32, -1, -grad_scale * scale, grad_scale_dev,
                            dbias, -grad_scale, sws, scatter_ws_bytes(M, d), st);
  }
  return 0;
}

static int ce_bwd_impl(const void* U, const void* W, const float* bias, float scale, const int64_t* labels,
                       int64_t label_base, const float* lse, float grad_scale, const float* grad_scale_dev,
                       int64_t M, int64_t N, int d, int dtype, int mode, float* dU, float* dW, float* dbias,
                       void* ws, size_t ws_bytes, rb_stream_t stream, void* dW_bf16, bool accumulate = false,
                       const int* m_dev = nullptr);

extern "C" int rb_ce_bwd(const void* U, const void* W, const float* bias, float scale, const int64_t* labels,
                         int64_t label_base, const float* lse, float grad_scale, const float* grad_scale_dev,
                         int64_t M, int64_t N, int d, int dtype, int mode, float* dU, float* dW, float* dbias,
                         const int32_t* m_dev, void* ws, size_t ws_bytes, rb_stream_t stream) {
  RB_RANGE("rb_ce_bwd");
  return ce_bwd_impl(U, W, bias, scale, labels, label_base, lse, grad_scale, grad_scale_dev, M, N, d, dtype, mode, dU, dW,
                     dbias, ws, ws_bytes, stream, nullptr, false, m_dev);
}

extern "C" int rb_ce_bwd_dw_bf16(const void* U, const void* W, const float* bias, float scale, const int64_t* labels,
                                 int64_t label_base, const float* lse, float grad_scale, const float* grad_scale_dev,
                                 int64_t M, int64_t N, int d, void* dW_bf16, float* dbias, const int32_t* m_dev, void* ws,
                                 size_t ws_bytes, rb_stream_t stream) {
  RB_RANGE("rb_ce_bwd_dw_bf16");
  if (!dW_bf16) return fail(RB_E_ARG, "null output");
  if (reinterpret_cast<uintptr_t>(dW_bf16) & 15) return fail(RB_E_ALIGN, "dW must be 16-byte aligned");
  return ce_bwd_impl(U, W, bias, scale, labels, label_base, lse, grad_scale, grad_scale_dev, M, N, d, RB_DTYPE_BF16,
                     RB_MODE_BF16, nullptr, nullptr, dbias, ws, ws_bytes, stream, dW_bf16, false, m_dev);
}

extern "C" int rb_ce_bwd_dw_bf16_acc(const void* U, const void* W, const float* bias, float scale, const int64_t* labels,
                                     int64_t label_base, const float* lse, float grad_scale, const float* grad_scale_dev,
                                     int64_t M, int64_t N, int d, void* dW_bf16, float* dbias, const int32_t* m_dev, void* ws,
                                     size_t ws_bytes, rb_stream_t stream) {
  RB_RANGE("rb_ce_bwd_dw_bf16_acc");
  if (!dW_bf16) return fail(RB_E_ARG, "null output");
  if (reinterpret_cast<uintptr_t>(dW_bf16) & 15) return fail(RB_E_ALIGN, "dW must be 16-byte aligned");
  return ce_bwd_impl(U, W, bias, scale, labels, label_base, lse, grad_scale, grad_scale_dev, M, N, d, RB_DTYPE_BF16,
                     RB_MODE_BF16, nullptr, nullptr, dbias, ws, ws_bytes, stream, dW_bf16, true, m_dev);
}

static int ce_bwd_impl(const void* U, const void* W, const float* bias, float scale, const int64_t* labels,
                       int64_t label_base, const float* lse, float grad_scale, const float* grad_scale_dev,
                       int64_t M, int64_t N, int d, int dtype, int mode, float* dU, float* dW, float* dbias,
                       void* ws, size_t ws_bytes, rb_stream_t stream, void* dW_bf16, bool accumulate, const int* m_dev) {
  DevInfo dv; if (int r = get_dev(dv)) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (int r = check_common(U, W, M, N, d, dtype, mode)) return r;
  if (d > 256) return fail(RB_E_UNSUPPORTED, "CE backward supports d <= 256");
  if (!labels || !lse) return fail(RB_E_ARG, "null pointer");
  if (!ws) return fail(RB_E_WORKSPACE, "workspace required");
  if (dbias && !dW && !dW_bf16) return fail(RB_E_ARG, "dbias is produced by the dW pass: pass dW too");
  Bump b(ws, ws_bytes);
  if (mode == RB_MODE_FP32X3)  // fp32 parity: exact fp32 passes (any sign of scale)
    return ce_bwd_f32(dv, static_cast<const float*>(U), static_cast<const float*>(W), bias, scale, labels, label_base, lse,
                      grad_scale, grad_scale_dev, M, N, d, dU, dW, dbias, b, st, m_dev);
  if (!(scale > 0.f)) return fail(RB_E_UNSUPPORTED, "CE backward needs scale > 0");

  if (dU) {  // no forward accumulator at hand: re-run the fused forward pass, then finish against the GLOBAL lse.
             // (Callers that kept rb_ce_fwd's dU_unnorm use rb_ce_du_finish and pass dU = NULL here.)
    float* rm = b.take<float>(M);
    float* rl = b.take<float>(M);
    float* rll = b.take<float>(M);
    float* du_un = b.take<float>(static_cast<size_t>(M) * d);
    if (!b.ok()) return fail(RB_E_WORKSPACE, "workspace too small: need %zu bytes", b.off);
    if (int r = ce_fwd_pair(dv, U, W, bias, scale, labels, label_base, M, N, d, rm, rl, rll, du_un, b, st, m_dev)) return r;
    const int grid = (int)((M * 32 + 255) / 256);
    ce_du_finish_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(du_un, rm, lse, static_cast<const __nv_bfloat16*>(W), labels,
                                                               label_base, N, grad_scale * scale, grad_scale_dev, (int)M, d, dU, m_dev);
    RB_LAUNCH_CHECK("ce_du_finish_kernel");
  }
  if (dW || dW_bf16) {
    const long long m_pad = ((M + 127) / 128) * 128;
    float* lse2 = b.take<float>(m_pad);
    if (!b.ok()) return fail(RB_E_WORKSPACE, "workspace too small");
    lse2_kernel<<<(int)((m_pad + 255) / 256), 256, 0, st>>>(lse, lse2, (int)M, (int)m_pad, m_dev);
    RB_LAUNCH_CHECK("lse2_kernel");
    return ce_bwd_dw_pair(dv, U, W, bias, scale, labels, label_base, lse2, grad_scale, grad_scale_dev, M, N, d, dW, dbias, b, st,
                          static_cast<__nv_bfloat16*>(dW_bf16), accumulate, m_dev);
  }
  return 0;
}

// ========================================================================= top-K eval
// Exact masked top-K in ONE tensor-core sweep over the catalog plus a short seeding sweep (no dense (B,N), no
// (B, N/128) tile-maximum matrix, no per-element selection work):
//   1. sweep<EPI_TOPK> over a PREFIX of the catalog (max(4K tiles, 2.5 % of the tiles); the whole catalog when it
//      is that small): masked maximum of every (row, 128-item tile)
//   2. tilemax_select:  per row the ladder tau0 = K-th largest clean tile maximum of the prefix (>= K unseen items
//      reach it, so every top-K member does too) and the checkpoints c_k = (K >> k)-th largest
//   3. sweep<EPI_CAND> over the whole catalog: every aligned group of 8 items whose maximum reaches the row's
//      running threshold is appended to the (row, split, warpgroup) sub-list as (maximum, group id); the threshold
//      climbs the ladder as the shared counters show >= K unseen items above a checkpoint (sweep.cuh)
//   4. topk_from_cands: cut at the K-th largest clean group maximum, exact fp32 re-scoring of the ~K groups left,
//      drop seen items, sort, keep K
//   5. rows with an overflowed sub-list (massive ties / catalogs too small for a threshold): exact scan of the row's
//      whole catalog (topk_refine).
// capacity of one candidate sub-list: ~6x the expected share of a sub-list (K (3 + log2(N/prefix)) candidates per
// row in total), a power of two in [32, 512]
static int topk_prefix_tiles(int n_tiles, int K) { return std::min(n_tiles, std::max(4 * K, (n_tiles + 39) / 40)); }
static int topk_candcap(int K, int n_sub, int n_tiles, int prefix_tiles) {
  int phases = 3;
  for (long long t = prefix_tiles; t < n_tiles; t *= 2) ++phases;
  const int want = (6 * K * phases + n_sub - 1) / n_sub;
  int c = 32;
  while (c < want && c < 512) c *= 2;
  return c;
}

// Workspace layout of rb_topk_eval behind the staged operands (one place: the entry, the size query and the
// debug hook below agree by construction).
struct TopkLayout {
  int xt, n0_tiles, n_sub, candcap;
  long long n0;
  Plan p, pp;
  int* crow32; int* col32; float* tmax; RowLadder* ladder; int* cand_cnt; int* overflow; uint2* cand;
  int* fb_count; int* fb_list; unsigned long long* fb_part;
};
static TopkLayout topk_layout(Bump& b, long long B, long long N, int d, int mode, int K, bool seen, long long nnz, int sms) {
  TopkLayout l{};
  l.xt = sweep_xt(mode, d, B);
  l.p = make_plan(B, N, sms, 1 << 20, 8, 128 * l.xt);
  l.n0_tiles = topk_prefix_tiles(l.p.n_strm_tiles, K);
  l.n0 = std::min<long long>(N, 1ll * l.n0_tiles * BN);
  l.pp = make_plan(B, l.n0, sms, 1 << 20, 8, 128 * l.xt);   // the seeding sweep over the prefix
  l.n_sub = 2 * l.p.n_splits;   // one sub-list per (split, tile parity): sweep.cuh, SUBS
  l.candcap = topk_candcap(K, l.n_sub, l.p.n_strm_tiles, l.n0_tiles);
  if (seen) {
    l.crow32 = b.take<int>(B + 1);
    l.col32 = b.take<int>(std::max<long long>(nnz, 1));
  }
  l.tmax = b.take<float>(static_cast<size_t>(B) * l.n0_tiles);
  l.ladder = b.take<RowLadder>(B);
  l.cand_cnt = b.take<int>(static_cast<size_t>(B) * l.n_sub);
  l.overflow = b.take<int>(B);
  l.cand = b.take<uint2>(static_cast<size_t>(B) * l.n_sub * l.candcap);
  l.fb_count = b.take<int>(1);
  l.fb_list = b.take<int>(B);
  l.fb_part = b.take<unsigned long long>(static_cast<size_t>(FB_ROWS) * FB_PARTS * 256);
  return l;
}

// Diagnostics for tests and profiling: where rb_topk_eval leaves its intermediate results in the workspace it was
// given (byte offsets; valid for the same arguments on the same device, bf16 mode or fp32x3 mode alike).
//   out[0] = n_sub, out[1] = cand_cap, out[2] = prefix tiles, out[3] = offset of the ladders (64 B per row),
//   out[4] = offset of cand_cnt (int32 [B][n_sub]), out[5] = offset of the overflow flags (int32 [B]),
//   out[6] = offset of the candidate lists (8 B entries [B][n_sub][cand_cap]), out[7] = bytes needed
extern "C" int rb_topk_debug_layout(int64_t B, int64_t N, int d, int K, int mode, int64_t nnz, int64_t* out) {
  RB_RANGE("rb_topk_debug_layout");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  if (!out) return fail(RB_E_ARG, "null output");
  Bump b(nullptr, ~size_t(0));
  b.take<char>(staged_bytes(B, d, mode) ? staged_bytes(B, d, mode) - 256 : 0);
  b.take<char>(staged_bytes(N, d, mode) ? staged_bytes(N, d, mode) - 256 : 0);
  TopkLayout l = topk_layout(b, B, N, d, mode, K, nnz > 0, nnz, dv.sms);
  out[0] = l.n_sub; out[1] = l.candcap; out[2] = l.n0_tiles;
  out[3] = reinterpret_cast<char*>(l.ladder) - static_cast<char*>(nullptr);
  out[4] = reinterpret_cast<char*>(l.cand_cnt) - static_cast<char*>(nullptr);
  out[5] = reinterpret_cast<char*>(l.overflow) - static_cast<char*>(nullptr);
  out[6] = reinterpret_cast<char*>(l.cand) - static_cast<char*>(nullptr);
  out[7] = static_cast<int64_t>(b.off);
  return 0;
}

extern "C" int rb_topk_eval(const void* U, const void* W, const float* bias, float scale, const int64_t* seen_crow,
                            const int64_t* seen_col, int64_t seen_nnz, int64_t id_base, int64_t B, int64_t N, int d,
                            int dtype, int mode, int K, float* top_vals, int32_t* top_ids, void* ws, size_t ws_bytes,
                            rb_stream_t stream) {
  RB_RANGE("rb_topk_eval");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (int r = check_common(U, W, B, N, d, dtype, mode)) return r;
  if (!top_vals || !top_ids) return fail(RB_E_ARG, "null output");
  if (K < 1 || K > 256) return fail(RB_E_UNSUPPORTED, "K=%d outside [1,256]", K);
  if (!(scale > 0.f)) return fail(RB_E_UNSUPPORTED, "top-K needs scale > 0");
  if (id_base < 0 || id_base + N > 0x7fffffffll) return fail(RB_E_ARG, "id_base + N = %lld does not fit the int32 ids this entry returns", (long long)(id_base + N));
  if (seen_nnz > 0x7fffffffll) return fail(RB_E_ARG, "seen_nnz = %lld does not fit int32 row offsets", (long long)seen_nnz);
  if ((seen_crow == nullptr) != (seen_col == nullptr)) return fail(RB_E_ARG, "seen_crow/seen_col must both be given");
  if (!ws) return fail(RB_E_WORKSPACE, "workspace required");
  Bump b(ws, ws_bytes);
  Operand ou, ow;
  if (int r = stage_operand(U, B, d, mode, b, ou, st)) return r;
  if (int r = stage_operand(W, N, d, mode, b, ow, st)) return r;
  if (seen_crow && seen_nnz < 0) return fail(RB_E_ARG, "seen_nnz < 0");
  const long long nnz = seen_crow ? seen_nnz : 0;
  TopkLayout lay = topk_layout(b, B, N, d, mode, K, seen_crow != nullptr, nnz, dv.sms);
  if (!b.ok()) return fail(RB_E_WORKSPACE, "workspace too small: need %zu bytes", b.off);
  const int xt = lay.xt, n0_tiles = lay.n0_tiles, n_sub = lay.n_sub, candcap = lay.candcap;
  const Plan p = lay.p, pp = lay.pp;
  const long long n0 = lay.n0;
  int* crow32 = lay.crow32; int* col32 = lay.col32;
  float* tmax = lay.tmax; RowLadder* ladder = lay.ladder; int* cand_cnt = lay.cand_cnt; int* overflow = lay.overflow;
  uint2* cand = lay.cand;
  if (seen_crow) {
    const long long n = std::max<long long>(B + 1, nnz);
    csr_local_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(seen_crow, seen_col, id_base, crow32, col32, B, nnz);
    RB_LAUNCH_CHECK("csr_local_kernel");
  }
  CUtensorMap ts, ty;
  if (int r = make_tmap(&ts, ou.ptr, ou.bf16, B, ou.cols, 128)) return r;
  if (int r = make_tmap(&ty, ow.ptr, ow.bf16, N, ow.cols, BN)) return r;
  SweepArgs a{};
  a.n_stat = (int)B; a.d = d; a.scale = scale; a.bias = bias;
  a.seen_crow = crow32; a.seen_col = col32; a.tile_max = tmax;
  a.ladder = ladder; a.k_need = K; a.cand = cand; a.cand_cnt = cand_cnt; a.cand_cap = candcap; a.n_sub = n_sub;
  // 1. seeding sweep over the prefix
  a.n_strm = (int)n0; a.n_stat_tiles = pp.n_stat_tiles; a.n_strm_tiles = pp.n_strm_tiles; a.n_splits = pp.n_splits;
  if (int r = launch_sweep_topk(mode, kc_for(d, mode), ts, ty, a, pp.grid, st, xt)) return r;
  // 2. the ladder
  const int grid_w = static_cast<int>((B * 32 + 127) / 128);
  if (K <= 128) tilemax_select_kernel<4><<<grid_w, 128, 0, st>>>(tmax, n0_tiles, B, K, ladder, lay.fb_count);
  else tilemax_select_kernel<8><<<grid_w, 128, 0, st>>>(tmax, n0_tiles, B, K, ladder, lay.fb_count);
  RB_LAUNCH_CHECK("tilemax_select_kernel");
  // 3. the candidate sweep over the whole catalog
  a.n_strm = (int)N; a.n_stat_tiles = p.n_stat_tiles; a.n_strm_tiles = p.n_strm_tiles; a.n_splits = p.n_splits;
  if (int r = launch_sweep_cand(mode, kc_for(d, mode), ts, ty, a, p.grid, st, xt)) return r;
  // 4. + 5. finish
  const int id_add = static_cast<int>(id_base);
  const int grid_r = static_cast<int>((B + 3) / 4);
  const int n_items = static_cast<int>(N);
  long long n_rows_ll = B;
  auto fallback = [&](auto kern, const void* Up, const void* Wp) -> int {
    int per_sm = 0;
    RB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, FB_THREADS, 0));
    if (per_sm < 1) return fail(RB_E_UNSUPPORTED, "topk_fallback_coop_kernel does not fit on this device");
    const int grid = std::min(per_sm, 2) * dv.sms;
    int dd = d, kk = K, ia = id_add, ni = n_items;
    float sc = scale;
    const int* ovf = overflow;
    void* params[] = {&Up, &Wp, (void*)&bias, &sc, &dd, &n_rows_ll, &ni, (void*)&crow32, (void*)&col32, &kk, &ia, (void*)&top_vals,
                      (void*)&top_ids, (void*)&ovf, (void*)&lay.fb_count, (void*)&lay.fb_list, (void*)&lay.fb_part};
    RB_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(grid), dim3(FB_THREADS), params, 0, st));
    RB_LAUNCH_CHECK("topk_fallback_coop_kernel");
    return 0;
  };
  if (dtype == RB_DTYPE_BF16) {
    const __nv_bfloat16* Ub = static_cast<const __nv_bfloat16*>(U); const __nv_bfloat16* Wb = static_cast<const __nv_bfloat16*>(W);
    if (K <= 128) topk_from_cands_kernel<__nv_bfloat16, 4><<<grid_r, 128, 0, st>>>(Ub, Wb, bias, scale, d, B, (int)N, cand, cand_cnt, n_sub, candcap, crow32, col32, K, id_add, top_vals, top_ids, overflow);
    else topk_from_cands_kernel<__nv_bfloat16, 8><<<grid_r, 128, 0, st>>>(Ub, Wb, bias, scale, d, B, (int)N, cand, cand_cnt, n_sub, candcap, crow32, col32, K, id_add, top_vals, top_ids, overflow);
    RB_LAUNCH_CHECK("topk_from_cands_kernel");
    if (K <= 128) return fallback(topk_fallback_coop_kernel<__nv_bfloat16, 4>, U, W);
    return fallback(topk_fallback_coop_kernel<__nv_bfloat16, 8>, U, W);
  }
  const float* Uf = static_cast<const float*>(U); const float* Wf = static_cast<const float*>(W);
  if (K <= 128) topk_from_cands_kernel<float, 4><<<grid_r, 128, 0, st>>>(Uf, Wf, bias, scale, d, B, (int)N, cand, cand_cnt, n_sub, candcap, crow32, col32, K, id_add, top_vals, top_ids, overflow);
  else topk_from_cands_kernel<float, 8><<<grid_r, 128, 0, st>>>(Uf, Wf, bias, scale, d, B, (int)N, cand, cand_cnt, n_sub, candcap, crow32, col32, K, id_add, top_vals, top_ids, overflow);
  RB_LAUNCH_CHECK("topk_from_cands_kernel");
  if (K <= 128) return fallback(topk_fallback_coop_kernel<float, 4>, U, W);
  return fallback(topk_fallback_coop_kernel<float, 8>, U, W);
}

extern "C" int rb_topk_merge(const float* vals, const int32_t* ids, int R, int64_t B, int K, float* out_vals,
                             int32_t* out_ids, rb_stream_t stream) {
  RB_RANGE("rb_topk_merge");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!vals || !ids || !out_vals || !out_ids) return fail(RB_E_ARG, "null pointer");
  if (R < 1 || B < 1 || K < 1 || K > 256) return fail(RB_E_ARG, "bad shape R=%d B=%lld K=%d", R, (long long)B, K);
  const int grid = static_cast<int>((B * 32 + 127) / 128);
  if (K <= 128) topk_merge_kernel<4><<<grid, 128, 0, st>>>(vals, ids, R, B, K, out_vals, out_ids, B * K);
  else topk_merge_kernel<8><<<grid, 128, 0, st>>>(vals, ids, R, B, K, out_vals, out_ids, B * K);
  RB_LAUNCH_CHECK("topk_merge_kernel");
  return 0;
}
// The same merge reading the lists where ONE all-gather of the per-rank [vals bits | ids] pairs left them:
// packed[R][2][B][K] (32-bit words; plane 0 = float32 values, plane 1 = int32 ids) -- no unpacking copies.
extern "C" int rb_topk_merge_packed(const int32_t* packed, int R, int64_t B, int K, float* out_vals, int32_t* out_ids,
                                    rb_stream_t stream) {
  RB_RANGE("rb_topk_merge_packed");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!packed || !out_vals || !out_ids) return fail(RB_E_ARG, "null pointer");
  if (R < 1 || B < 1 || K < 1 || K > 256) return fail(RB_E_ARG, "bad shape R=%d B=%lld K=%d", R, (long long)B, K);
  const float* vals = reinterpret_cast<const float*>(packed);
  const int32_t* ids = packed + B * K;
  const int grid = static_cast<int>((B * 32 + 127) / 128);
  if (K <= 128) topk_merge_kernel<4><<<grid, 128, 0, st>>>(vals, ids, R, B, K, out_vals, out_ids, 2 * B * K);
  else topk_merge_kernel<8><<<grid, 128, 0, st>>>(vals, ids, R, B, K, out_vals, out_ids, 2 * B * K);
  RB_LAUNCH_CHECK("topk_merge_kernel");
  return 0;
}

extern "C" int rb_topk_hits(const int32_t* top_ids, const int64_t* target_crow, const int64_t* target_col, int64_t B,
                            int K, float* hits, rb_stream_t stream) {
  RB_RANGE("rb_topk_hits");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!top_ids || !target_crow || !hits) return fail(RB_E_ARG, "null pointer");
  if (B < 0 || K < 1) return fail(RB_E_ARG, "bad shape B=%lld K=%d", (long long)B, K);
  if (B == 0) return 0;
  const long long n = B * K;
  topk_hits_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(top_ids, target_crow, target_col, B, K, hits);
  RB_LAUNCH_CHECK("topk_hits_kernel");
  return 0;
}

// Merge of the per-rank row statistics of a row-sharded table (stats[R][3][M] from one all-gather) into the global
// lse and label logit of every query row.
extern "C" int rb_rowstats_merge(const float* stats, int n_ranks, int64_t M, float* lse, float* label_logit, rb_stream_t stream) {
  RB_RANGE("rb_rowstats_merge");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  if (!stats || !lse || !label_logit) return fail(RB_E_ARG, "null pointer");
  if (n_ranks < 1 || M < 0) return fail(RB_E_ARG, "bad shape");
  if (M == 0) return 0;
  rowstats_merge_kernel<<<static_cast<int>((M + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(stats, n_ranks, M, lse, label_logit);
  RB_LAUNCH_CHECK("rowstats_merge_kernel");
  return 0;
}

// Batch means of n_metrics METRIC@k values straight from the ranked ids (kinds: 0 HITRATE, 1 RECALL, 2 PRECISION, 3 NDCG,
// 4 MRR; kinds / ks are HOST arrays).  w[K] = 1/log2(rank+2), w_cum[K] = its running sum (device, float32);
// partial: device scratch of rb_topk_metrics_blocks(B) * 32 doubles; out[n_metrics] float32 on the device.
extern "C" int rb_topk_metrics_blocks(int64_t B) {
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((B + 7) / 8, 148 * 4)));
}
extern "C" int rb_topk_metrics(const int32_t* top_ids, const int64_t* target_crow, const int64_t* target_col, int64_t B, int K,
                               const float* w, const float* w_cum, const int32_t* kinds, const int32_t* ks, int n_metrics,
                               double* partial, float* out, rb_stream_t stream) {
  RB_RANGE("rb_topk_metrics");
  DevInfo dv; if (int r = get_dev(dv)) return r;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!top_ids || !target_crow || !w || !w_cum || !kinds || !ks || !partial || !out) return fail(RB_E_ARG, "null pointer");
  if (B <= 0 || K < 1) return fail(RB_E_ARG, "bad shape B=%lld K=%d", (long long)B, K);
  if (n_metrics < 1 || n_metrics > TM_MAX_METRICS) return fail(RB_E_ARG, "1 <= n_metrics <= %d (got %d)", TM_MAX_METRICS, n_metrics);
  MetricSpec spec{};
  spec.n = n_metrics;
  for (int m = 0; m < n_metrics; ++m) {
    if (kinds[m] < 0 || kinds[m] > 4) return fail(RB_E_ARG, "unknown metric kind %d", kinds[m]);
    if (ks[m] < 1 || ks[m] > K) return fail(RB_E_ARG, "metric %d: k=%d outside [1, %d]", m, ks[m], K);
    spec.kind[m] = kinds[m]; spec.k[m] = ks[m];
  }
  const int blocks = rb_topk_metrics_blocks(B);
  topk_metrics_kernel<<<blocks, 256, 0, st>>>(top_ids, target_crow, target_col, B, K, w, w_cum, spec, partial);
  RB_LAUNCH_CHECK("topk_metrics_kernel");
  topk_metrics_finish_kernel<<<1, 32, 0, st>>>(partial, blocks, n_metrics, B, out);
  RB_LAUNCH_CHECK("topk_metrics_finish_kernel");
  return 0;
}

// ========================================================================== workspace
extern "C" size_t rb_workspace_bytes(int op, int64_t M, int64_t N, int d, int K, int mode, int64_t nnz) {
  int sms = 148;  // B200; refined from the current device when there is one
  { DevInfo dv; if (get_dev(dv) == 0 && dv.sms > 0) sms = dv.sms; }
  size_t need = 4096;
  switch (op) {
    case RB_OP_SCATTER_ADD: return scatter_ws_bytes(nnz, d) + 4096;
    case RB_OP_SCORE_DENSE: return need + staged_bytes(M, d, mode) + staged_bytes(N, d, mode);
    case RB_OP_CE_FWD: {  // the larger of the statistics-only sweep and the fused forward+dU pass
      const int xt = sweep_xt(mode, d, M);
      Plan p = make_plan(M, N, sms, 1 << 20, 8, 128 * xt);
      const size_t m_pad = static_cast<size_t>(p.n_stat_tiles) * 128 * xt;
      const size_t stats = staged_bytes(M, d, mode) + staged_bytes(N, d, mode) + m_pad * 4 + 3 * m_pad * p.n_splits * 4 + 2048;
      return need + std::max(stats, pair_fwd_ws(M, N, d, sms));
    }
    case RB_OP_CE_BWD: {
      if (mode == RB_MODE_FP32X3) return need + f32grad_ws(M, N, d, sms) + scatter_ws_bytes(M, d) + static_cast<size_t>(N) * 4 + 2048;
      size_t n = need + static_cast<size_t>(M) * (d + 3) * 4 + 2048 + pair_fwd_ws(M, N, d, sms);  // dU by recompute
      Plan pw = make_plan(N, M, sms, 64, 8, pair_rows(d));
      n += ((M + 127) / 128) * 128 * 4 + 512;
      n += (pw.n_splits > 1 ? static_cast<size_t>(pw.n_splits) * N * (d + 1) * 4 : 0) + 1024;
      n += static_cast<size_t>(pw.n_stat_tiles) * pair_rows(d) * 4 + 512;  // bias2
      // rb_ce_bwd_dw_bf16: slot map + side table (one split) or an fp32 staging copy of dW (several splits)
      n += static_cast<size_t>(N) * 4 + static_cast<size_t>(M) * (d * 4 + 12) + 2048;   // label owners + side table
      if (pw.n_splits > 1 || M > 16384) n += static_cast<size_t>(N) * d * 4 + 512;        // fp32 staging copy of dW
      n += scatter_ws_bytes(M, d) + 512;
      return n;
    }
    case RB_OP_TOPK_EVAL: {
      Bump b(nullptr, ~size_t(0));
      b.take<char>(staged_bytes(M, d, mode));
      b.take<char>(staged_bytes(N, d, mode));
      topk_layout(b, M, N, d, mode, K, nnz > 0, nnz, sms);
      return need + b.off + 8192;
    }
    default: return 0;
  }
}
