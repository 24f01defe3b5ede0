#!/bin/bash
# Round-2 GPU visit: every GPU test file in its own process (a trapped kernel must not poison the rest), smoke, bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/bf16_flips.jsonl
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name: $*"; timeout "${TMO:-900}" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n "${TAIL:-6}" gpurun_out/$name.log | cut -c1-${CUT:-1500}; }
if [ -z "$SKIP_TESTS" ]; then
for f in ${TESTS:-test_gpu_heads test_gpu_kernels test_gpu_fixed_weights test_gpu_models test_gpu_tcn test_gpu_ensemble}; do
  run t_$f python -m pytest tests/$f.py -m gpu -q -x --timeout 600 -s
done
run smoke python __graft_entry__.py --smoke
fi
[ -n "$SANITIZE" ] && TMO=1200 run sanitize compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_heads.py -m gpu -q -x -k "not 8192 and not 513 and not 600" --timeout 1100
[ -n "$BREAKDOWN" ] && TAIL=24 CUT=200 run step_breakdown python scripts/step_breakdown.py
[ -n "$LSTM_BENCH" ] && CUT=2000 TAIL=1 run bench_lstm python scripts/bench_lstm.py
if [ -z "$SKIP_BENCH" ]; then
CUT=4000 TAIL=2 run bench python bench.py --steps 20 --warmup 3 ${BENCH_ARGS}
TAIL=2 CUT=1500 run bench_ref python bench.py --impl reference --steps 4 --warmup 1
fi
if [ -n "$PROFILE" ]; then
  PCMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-aux --no-graph"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $PCMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
  for k in $PROFILE_KERNELS; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s ${NCU_SKIP:-6} -c ${NCU_COUNT:-3} -f -o gpurun_out/prof_$k $PCMD > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit $?"
  done
fi
exit 0
