#!/bin/bash
# Short GPU visit: a pytest selection ($K), then the bench (graph mode) and optionally an ncu launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name: $*"; timeout "${TMO:-300}" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n "${TAIL:-15}" gpurun_out/$name.log | cut -c1-${CUT:-600}; }
[ -n "$K" ] && run q_tests python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py -m gpu -q -s -k "$K" --timeout 200
[ -n "$BENCH" ] && CUT=3000 TAIL=2 run q_bench python bench.py --steps 10 --warmup 3 $BENCH
if [ -n "$LAUNCHES" ]; then
  PCMD="python bench.py --steps 2 --warmup 3 --no-e2e --cpu-windows 512 --no-graph"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_q.csv $PCMD > gpurun_out/ncu_launches_q.log 2>&1
  echo "ncu launches exit $?"
fi
exit 0
