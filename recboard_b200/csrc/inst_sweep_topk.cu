// One translation unit per epilogue / pass of the tcgen05 kernel templates (parallel build; see launch.cuh).
#include "launch.cuh"
namespace rb {
int launch_sweep_topk(int mode, int kc, const CUtensorMap& ts, const CUtensorMap& ty, const SweepArgs& a, int grid, cudaStream_t st, int xt) { return launch_sweep<EPI_TOPK, true>(mode, kc, ts, ty, a, grid, st, xt); }
}  // namespace rb
