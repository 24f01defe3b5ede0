#!/bin/bash
# Final-style one-GPU visit: every GPU test, smoke, the default bench line, the ensemble and frame benches, the reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/t_all.log 2>&1; echo "exit $? all gpu tests"; tail -3 gpurun_out/t_all.log | cut -c1-400
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "exit $? smoke"; tail -1 gpurun_out/smoke.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "exit $? bench"
tail -1 gpurun_out/bench.log | python -c "
import json, sys
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'ms_per_step_spread_rank0', 'clocks', 'gpu_launches')})
print(d['e2e']['value'], d['roofline']['frac'], d['roofline_gemm']['frac'])
print(d['other_configs'])"
timeout 300 python scripts/bench_ensemble.py > gpurun_out/bench_ensemble.log 2>&1; echo "exit $? ensemble"; tail -1 gpurun_out/bench_ensemble.log | cut -c1-900
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "exit $? ref"; tail -1 gpurun_out/bench_ref.log | cut -c1-400
if [ -n "$PROFILE" ]; then
  PCMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-aux --cpu-windows 512 --no-graph"
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $PCMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:tcn_layer_fwd_bf16 -s 20 -c 2 -f -o gpurun_out/prof_tcn_bf16 python scripts/bench_ensemble.py --videos 512 --reps 1 > gpurun_out/ncu_tcn_bf16.log 2>&1; echo "ncu tcn bf16 exit $?"
fi
exit 0
