#!/bin/bash
# N-GPU visit: data-parallel bench (weak scaling), eager then graph.
cd "$(dirname "$0")/.."
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_multi.txt 2>&1
run() { name=$1; shift; echo "=== $name: $*"; timeout "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; grep '^{' gpurun_out/$name.log | tail -1 | cut -c1-400; tail -3 gpurun_out/$name.log | cut -c1-300; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run mg_eager_$N $TR bench.py --gpus $N --steps 10 --warmup 3 --no-graph --no-e2e --cpu-windows 512
run mg_graph_$N $TR bench.py --gpus $N --steps 10 --warmup 3 --cpu-windows 512
run mg_ref_$N $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1
run sg_graph python bench.py --gpus 1 --steps 10 --warmup 3 --cpu-windows 512
