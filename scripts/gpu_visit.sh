#!/bin/bash
# Generic short visit: optional pytest selection, bench without the auxiliary configs, step breakdown, ncu launch list.
#   K="expr" FILES="tests/..." BENCH=1 BREAKDOWN=1 LAUNCHES=1 EXTRA="cmd" bash scripts/gpu_visit.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name: $*"; timeout "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n "${TAIL:-8}" gpurun_out/$name.log | cut -c1-${CUT:-1200}; }
[ -n "$K" ] && run v_tests python -m pytest ${FILES:-tests/test_gpu_kernels.py tests/test_gpu_models.py} -m gpu -q -x -k "$K" --timeout 300
[ -n "$BREAKDOWN" ] && TAIL=11 CUT=200 run step_breakdown env ONLY_NO_PREFETCH=1 python scripts/step_breakdown.py
[ -n "$BENCH" ] && CUT=6000 TAIL=1 run v_bench python bench.py --steps 20 --warmup 3 --no-aux $BENCH_ARGS
if [ -n "$LAUNCHES" ]; then
  PCMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-aux --no-graph"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $PCMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
fi
[ -n "$EXTRA" ] && CUT=3000 TAIL=${ETAIL:-8} run v_extra bash -c "$EXTRA"
exit 0
