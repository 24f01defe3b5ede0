#!/bin/bash
# One GPU-box visit: parity tests in isolated processes (a trapped kernel must not poison the rest), then the bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name: $*"; timeout "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n "${TAIL:-15}" gpurun_out/$name.log | cut -c1-${CUT:-3000}; }
run t_kernels python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not tcgen05" -x --timeout 300
run t_tcgen05 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tcgen05" --timeout 120
run t_models_fp32 python -m pytest tests/test_gpu_models.py -m gpu -q -k "not bf16" --timeout 600
run t_models_bf16 python -m pytest tests/test_gpu_models.py -m gpu -q -k "bf16" --timeout 300
run smoke python __graft_entry__.py --smoke
TAIL=2 run bench_fp32 python bench.py --steps 5 --warmup 3 --precision fp32 --batch 2048 --videos 256 --cpu-windows 1024
TAIL=2 run bench_cudnn python bench.py --steps 10 --warmup 3 --lstm-impl cudnn --no-graph --no-e2e --cpu-windows 512
TAIL=2 run bench_per_step python bench.py --steps 10 --warmup 3 --lstm-impl b200_per_step --no-e2e --cpu-windows 512
TAIL=2 run bench_eager python bench.py --steps 10 --warmup 3 --no-graph --no-e2e --cpu-windows 512
TAIL=2 run bench python bench.py --steps 20 --warmup 3
TAIL=2 run bench_ref python bench.py --impl reference --steps 4 --warmup 1
if [ -n "$PROFILE" ]; then
  # launch list (device time per launch; compare SHARES) and full captures of the dominant kernels
  PCMD="python bench.py --steps 2 --warmup 3 --no-e2e --cpu-windows 512 --no-graph"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $PCMD > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  ncu --set full --clock-control none --import-source on -k regex:gather_norm -s 6 -c 2 -f -o gpurun_out/prof_gather $PCMD > gpurun_out/ncu_gather.log 2>&1
  echo "ncu gather exit $?"
  ncu --set full --clock-control none --import-source on -k regex:lstm_rec -s 6 -c 6 -f -o gpurun_out/prof_rec $PCMD > gpurun_out/ncu_rec.log 2>&1
  echo "ncu rec exit $?"
  ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05_kernel -s 20 -c 1 -f -o gpurun_out/prof_gemm $PCMD > gpurun_out/ncu_gemm_step.log 2>&1
  echo "ncu gemm exit $?"
fi
exit 0
