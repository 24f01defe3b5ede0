#!/bin/bash
# N-GPU check of the final build: replica / kernel check through the public API, then the driver's own command.
cd "$(dirname "$0")/.."
N=${N:-4}
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name: $*"; S=$(date +%s); timeout "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name) in $(( $(date +%s) - S )) s"; grep '^{' gpurun_out/$name.log | tail -1 | cut -c1-${CUT:-400}; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
CUT=2500 B200MED_REQUIRE_PEER=1 run dp_window_check_$N $TR scripts/dp_window_check.py
CUT=1200 run mg_full_$N $TR bench.py --gpus $N --steps 20 --warmup 3
exit 0
