#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list of one pass of every sweep.
# usage: tools/gpu_round.sh <tag> [full]   (outputs under gpurun_out/<tag>_*)
tag=${1:-run}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python tests/prof_target.py > gpurun_out/${tag}_ncu.log 2>&1
# launch list of the bench command itself (2 timed steps, no pre-warm: the kernels' SHARES of a step, not a bench value)
RB_BENCH_PREWARM=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${tag}_launches.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows[-60:]:
    print(r[4][:90], r[-1])
PY
if [ "$2" == "full" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pair_kernel|sweep_kernel' -c 6 -o gpurun_out/${tag}_prof python tests/prof_target.py 1000000 dU,dW,topk > gpurun_out/${tag}_ncu_full.log 2>&1
  tail -3 gpurun_out/${tag}_ncu_full.log
fi
