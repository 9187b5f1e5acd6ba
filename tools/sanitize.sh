#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (small shapes only: the tools slow kernels down 10-100x).
#   tools/sanitize.sh <tag>   -> gpurun_out/<tag>_{memcheck,racecheck,synccheck}.log
tag=${1:-san}
mkdir -p gpurun_out
SMALL='not full_size and not config2 and not config3 and not 1_000_001 and not 5_000_000 and not 40_000 and not 70_000 and not 50_000 and not 17000'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SMALL" > gpurun_out/${tag}_${tool}.full.log 2>&1
  echo "exit=$?" >> gpurun_out/${tag}_${tool}.full.log
  grep -E "passed|failed|ERROR SUMMARY|exit=|RACECHECK SUMMARY|hazard|Error" gpurun_out/${tag}_${tool}.full.log | sort | uniq -c | sort -rn | head -25 > gpurun_out/${tag}_${tool}.log
  tail -c 3000 gpurun_out/${tag}_${tool}.full.log > gpurun_out/${tag}_${tool}.tail.log
  rm -f gpurun_out/${tag}_${tool}.full.log
done
