// Launchers of the tcgen05 kernel templates.  Each epilogue / pass is instantiated in its own translation unit
// (inst_*.cu) so that the library builds in parallel; abi.cu plans the work and calls these.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/recboard_b200.h"
#include "sweep.cuh"
#include "pair.cuh"

#ifndef RB_LSE_NWG
#define RB_LSE_NWG 4   // epilogue warpgroups of the statistics sweep with two stationary tiles (2 or 4; sweep.cuh)
#endif
namespace rb {

constexpr int BN = 128;  // streamed tile rows

// error plumbing of abi.cu (thread-local last-error string, launch counter)
int host_fail(int code, const char* msg);
int host_cuda_fail(cudaError_t e, const char* what);
void count_launch();

int launch_sweep_dense(int mode, int kc, const CUtensorMap& ts, const CUtensorMap& ty, const SweepArgs& a, int grid, cudaStream_t st, int xt = 1);
int launch_sweep_lse(int mode, int kc, const CUtensorMap& ts, const CUtensorMap& ty, const SweepArgs& a, int grid, cudaStream_t st, int xt = 1);
int launch_sweep_topk(int mode, int kc, const CUtensorMap& ts, const CUtensorMap& ty, const SweepArgs& a, int grid, cudaStream_t st, int xt = 1);
int launch_sweep_cand(int mode, int kc, const CUtensorMap& ts, const CUtensorMap& ty, const SweepArgs& a, int grid, cudaStream_t st, int xt = 1);
int launch_pair_fwd(int kc, bool bias, const CUtensorMap& ts, const CUtensorMap& ty, const PairArgs& a, int grid, cudaStream_t st);
int launch_pair_dw(int kc, bool bias, const CUtensorMap& ts, const CUtensorMap& ty, const PairArgs& a, int grid, cudaStream_t st);

template <class C>
int launch_sweep_t(const CUtensorMap& ts, const CUtensorMap& ty, const SweepArgs& a, int grid, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(sweep_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  if (e != cudaSuccess) return host_cuda_fail(e, "cudaFuncSetAttribute(sweep_kernel)");
  sweep_kernel<C><<<grid, C::THREADS, C::SMEM_BYTES, st>>>(ts, ty, a);
  count_launch();
  e = cudaGetLastError();
  return e != cudaSuccess ? host_cuda_fail(e, "sweep_kernel") : 0;
}

// NS = pipeline stages (whole 32 KB tiles up to two K-chunks, single 16 KB chunks beyond: SweepCfg::SC),
// chosen so that SMEM stays under 227 KB next to the stationary tile(s)
template <int EPI, bool ROWS>
int launch_sweep(int mode, int kc, const CUtensorMap& ts, const CUtensorMap& ty, const SweepArgs& a, int grid,
                 cudaStream_t st, int xt) {
  if (xt == 2) {
    if constexpr (EPI != EPI_DENSE) {
      // the top-K sweeps are bound by their epilogue's own instruction stream: four epilogue warpgroups (one per S buffer)
      constexpr int NWG = (EPI == EPI_CAND || EPI == EPI_TOPK || (EPI == EPI_LSE && RB_LSE_NWG == 4)) ? 4 : 2;
      if (mode == RB_MODE_BF16) {
        if (a.bias == nullptr) {   // the bias branches compiled out (SweepCfg::MAYBE_BIAS)
          if (kc == 1) return launch_sweep_t<SweepCfg<EPI, DT_BF16, 1, BN, 10, ROWS, 2, NWG, false>>(ts, ty, a, grid, st);
          if (kc == 2) return launch_sweep_t<SweepCfg<EPI, DT_BF16, 2, BN, 4, ROWS, 2, NWG, false>>(ts, ty, a, grid, st);
          if (kc <= 4) return launch_sweep_t<SweepCfg<EPI, DT_BF16, 4, BN, 5, ROWS, 2, NWG, false>>(ts, ty, a, grid, st);
        }
        if (kc == 1) return launch_sweep_t<SweepCfg<EPI, DT_BF16, 1, BN, 10, ROWS, 2, NWG>>(ts, ty, a, grid, st);
        if (kc == 2) return launch_sweep_t<SweepCfg<EPI, DT_BF16, 2, BN, 4, ROWS, 2, NWG>>(ts, ty, a, grid, st);
        if (kc <= 4) return launch_sweep_t<SweepCfg<EPI, DT_BF16, 4, BN, 5, ROWS, 2, NWG>>(ts, ty, a, grid, st);
      } else {
        if (kc == 1) return launch_sweep_t<SweepCfg<EPI, DT_TF32X3, 1, BN, 4, ROWS, 2, NWG>>(ts, ty, a, grid, st);
        if (kc == 2) return launch_sweep_t<SweepCfg<EPI, DT_TF32X3, 2, BN, 5, ROWS, 2, NWG>>(ts, ty, a, grid, st);
      }
    }
    return host_fail(RB_E_UNSUPPORTED, "unsupported feature width for two stationary tiles");
  }
  if (mode == RB_MODE_BF16) {
    if (kc == 1) return launch_sweep_t<SweepCfg<EPI, DT_BF16, 1, BN, 10, ROWS>>(ts, ty, a, grid, st);
    if (kc == 2) return launch_sweep_t<SweepCfg<EPI, DT_BF16, 2, BN, 5, ROWS>>(ts, ty, a, grid, st);
    if (kc <= 4) return launch_sweep_t<SweepCfg<EPI, DT_BF16, 4, BN, 8, ROWS>>(ts, ty, a, grid, st);
  } else {
    if (kc == 1) return launch_sweep_t<SweepCfg<EPI, DT_TF32X3, 1, BN, 5, ROWS>>(ts, ty, a, grid, st);
    if (kc == 2) return launch_sweep_t<SweepCfg<EPI, DT_TF32X3, 2, BN, 8, ROWS>>(ts, ty, a, grid, st);
    if (kc == 4) return launch_sweep_t<SweepCfg<EPI, DT_TF32X3, 4, BN, 5, ROWS>>(ts, ty, a, grid, st);
  }
  return host_fail(RB_E_UNSUPPORTED, "unsupported feature width for this mode");
}

template <class C>
int launch_pair_t(const CUtensorMap& ts, const CUtensorMap& ty, const PairArgs& a, int grid, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(pair_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  if (e != cudaSuccess) return host_cuda_fail(e, "cudaFuncSetAttribute(pair_kernel)");
  pair_kernel<C><<<grid, PAIR_THREADS, C::SMEM_BYTES, st>>>(ts, ty, a);
  count_launch();
  e = cudaGetLastError();
  return e != cudaSuccess ? host_cuda_fail(e, "pair_kernel") : 0;
}
template <int PASS>
int launch_pair(int kc, bool bias, const CUtensorMap& ts, const CUtensorMap& ty, const PairArgs& a, int grid, cudaStream_t st) {
  if (kc == 1) {
    if (bias) return launch_pair_t<PairCfg<PASS, 1, 6, true>>(ts, ty, a, grid, st);
    return launch_pair_t<PairCfg<PASS, 1, 6, false>>(ts, ty, a, grid, st);
  }
  if (kc == 2) {
    if (bias) return launch_pair_t<PairCfg<PASS, 2, 4, true>>(ts, ty, a, grid, st);
    return launch_pair_t<PairCfg<PASS, 2, 4, false>>(ts, ty, a, grid, st);
  }
  if (kc <= 4) {  // 128 < d <= 256: the d-split variant (one stationary tile, two output column halves; pair.cuh)
    if (bias) return launch_pair_t<PairCfg<PASS, 2, 2, true, true>>(ts, ty, a, grid, st);
    return launch_pair_t<PairCfg<PASS, 2, 2, false, true>>(ts, ty, a, grid, st);
  }
  return host_fail(RB_E_UNSUPPORTED, "the fused CE passes support d <= 256");
}

}  // namespace rb
