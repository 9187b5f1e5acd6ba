#!/bin/bash
# Builds the stand-alone micro-benchmarks under tests/ (sm_100a) into tests/_probe/ (git-ignored; the
# binaries travel to the GPU box with the gpurun snapshot).  usage: bash tools/build_probe.sh
set -e
cd "$(dirname "$0")/.."
mkdir -p tests/_probe
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -o tests/_probe/probe_cta2 tests/probe_cta2.cu "$@"
echo built tests/_probe/probe_cta2
$NVCC -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -o tests/_probe/probe_cta2_tma tests/probe_cta2_tma.cu "$@"
echo built tests/_probe/probe_cta2_tma
$NVCC -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -o tests/_probe/probe_cta2_sweep tests/probe_cta2_sweep.cu "$@"
echo built tests/_probe/probe_cta2_sweep
