#!/bin/bash
# N-GPU visit (round 2): replica check through the public API, then the weak-scaling bench at N and at 1 on the same box.
cd "$(dirname "$0")/.."
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_multi.txt 2>&1
run() { name=$1; shift; echo "=== $name: $*"; timeout "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; grep '^{' gpurun_out/$name.log | tail -1 | cut -c1-${CUT:-500}; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
[ -z "$SKIP_CHECK" ] && CUT=1500 run dp_window_check_$N $TR scripts/dp_window_check.py
run mg_graph_$N $TR bench.py --gpus $N --steps 20 --warmup 3 --no-ensemble
B200MED_PEER_EXCHANGE=0 run mg_graph_nccl_$N $TR bench.py --gpus $N --steps 20 --warmup 3 --no-aux --no-e2e
[ -z "$SKIP_SINGLE" ] && run sg_graph python bench.py --gpus 1 --steps 20 --warmup 3 --no-aux --no-e2e
exit 0
