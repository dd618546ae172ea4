#!/bin/bash
# Frame-sharded runs on N GPUs of one box (gpurun --gpus N -- 'bash tools/evidence_multigpu.sh N'): the two-GPU NCCL
# parity test, then bench.py config 4 and config 4 + sharded optimize_feature.  8 ranks take ~110 s per bench line.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02_${N}gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -x -k "nccl" 2>&1 | tail -12 > gpurun_out/r02_${N}gpu_nccl_test.txt
for wl in config4 config4opt; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 15 --warmup 3 --workload $wl > gpurun_out/r02_bench_${wl}_${N}gpu.json 2> gpurun_out/r02_bench_${wl}_${N}gpu.err
  grep '^{' gpurun_out/r02_bench_${wl}_${N}gpu.json | cut -c1-260
done
