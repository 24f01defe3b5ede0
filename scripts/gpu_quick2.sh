#!/bin/bash
# Short visit: a pytest selection ($K over $FILES), the in-graph step breakdown, a short bench without the auxiliary configs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name: $*"; timeout "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n "${TAIL:-8}" gpurun_out/$name.log | cut -c1-${CUT:-1200}; }
[ -n "$K" ] && run q_tests python -m pytest ${FILES:-tests/test_gpu_kernels.py tests/test_gpu_models.py} -m gpu -q -x -k "$K" --timeout 300
[ -n "$BREAKDOWN" ] && TAIL=22 CUT=200 run step_breakdown python scripts/step_breakdown.py
[ -n "$BENCH" ] && CUT=4000 TAIL=1 run q_bench python bench.py --steps 20 --warmup 3 --no-aux $BENCH_ARGS
[ -n "$EXTRA" ] && CUT=3000 TAIL=6 run q_extra bash -c "$EXTRA"
exit 0
