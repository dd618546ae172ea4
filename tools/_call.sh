mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02_8gpu.txt 2>&1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 15 --warmup 3 > gpurun_out/r02_bench_config4_8gpu.json 2> gpurun_out/r02_bench_config4_8gpu.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 15 --warmup 3 > gpurun_out/r02_bench_config4_4gpu.json 2> gpurun_out/r02_bench_config4_4gpu.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 15 --warmup 3 --workload config4opt > gpurun_out/r02_bench_config4opt_8gpu.json 2> gpurun_out/r02_bench_config4opt_8gpu.err
for f in config4_8gpu config4_4gpu config4opt_8gpu; do echo == $f; grep '^{' gpurun_out/r02_bench_$f.json | cut -c1-260; tail -2 gpurun_out/r02_bench_$f.err | cut -c1-300; done
