#!/bin/bash
# 8-GPU confirmation of the final build: the driver's own commands (reference arm under torchrun, then the full bench).
cd "$(dirname "$0")/.."
N=${N:-8}
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name: $*"; S=$(date +%s); timeout "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name) in $(( $(date +%s) - S )) s"; grep '^{' gpurun_out/$name.log | tail -1 | cut -c1-${CUT:-400}; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
CUT=300 run mg_ref_$N $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1
CUT=300 run mg_full_$N $TR bench.py --gpus $N --steps 20 --warmup 3
exit 0
