mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02_2gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -x -k "nccl" 2>&1 | tail -12 > gpurun_out/r02_2gpu_nccl_test.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 15 --warmup 3 > gpurun_out/r02_bench_config4_2gpu.json 2> gpurun_out/r02_bench_config4_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 15 --warmup 3 --workload config4opt > gpurun_out/r02_bench_config4opt_2gpu.json 2> gpurun_out/r02_bench_config4opt_2gpu.err
timeout 400 python bench.py --workload config4 --steps 15 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_config4_1gpu.json 2> gpurun_out/r02_bench_config4_1gpu.err
timeout 500 python bench.py --workload config4opt --steps 15 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_config4opt_1gpu.json 2> gpurun_out/r02_bench_config4opt_1gpu.err
tail -5 gpurun_out/r02_2gpu_nccl_test.txt | cut -c1-300; for f in config4_2gpu config4opt_2gpu config4_1gpu config4opt_1gpu; do echo == $f; cut -c1-330 gpurun_out/r02_bench_$f.json; tail -3 gpurun_out/r02_bench_$f.err | cut -c1-300; done
