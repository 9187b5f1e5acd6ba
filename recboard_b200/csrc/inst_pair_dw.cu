// One translation unit per epilogue / pass of the tcgen05 kernel templates (parallel build; see launch.cuh).
#include "launch.cuh"
namespace rb {
int launch_pair_dw(int kc, bool bias, const CUtensorMap& ts, const CUtensorMap& ty, const PairArgs& a, int grid, cudaStream_t st) { return launch_pair<PASS_DW>(kc, bias, ts, ty, a, grid, st); }
}  // namespace rb
