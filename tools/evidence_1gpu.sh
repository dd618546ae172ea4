#!/bin/bash
# The round-end evidence run on ONE B200 (tools/gpu.sh 3000 'bash tools/evidence_1gpu.sh'): GPU tests, attention
# micro-benchmark with torch SDPA beside it, bench lines (config 2 / 3), ncu launch list of one step, `ncu --set full` of
# every hot-path kernel summarised on the box (the .ncu-rep of 30 kernels exceeds what gpurun copies back), and one
# source-level capture of the dominant kernel.  Everything lands in gpurun_out/r02_*; copy what is to be judged to profiles/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_gputests.txt
timeout 300 python tools/bench_attn.py default FRESCO_ATTN_WIDE=1 FRESCO_ATTN_WIDE=2 FRESCO_ATTN_WIDE=0 --sdpa > gpurun_out/r02_attn_microbench_final.jsonl 2>&1
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 500 python bench.py --workload config3 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_config3.json 2> gpurun_out/r02_bench_config3.err
timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --profile-mode --steps 1 --warmup 0 > gpurun_out/r02_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1
gzip -f gpurun_out/r02_launches.csv
PROF_ITERS=1 timeout 700 ncu --set full --clock-control none -k "regex:fresco_attn|temporal_attn|warp_chain|warp_loss|gram|kv_compact|adam|adain|gmflow|tile_gemm|dilate|project" -c 30 -f -o gpurun_out/r02_kernels python tools/prof_kernels.py > gpurun_out/r02_ncu.log 2>&1
python tools/ncu_table.py gpurun_out/r02_kernels.ncu-rep > gpurun_out/r02_kernels_ncu.txt 2>&1
ncu -i gpurun_out/r02_kernels.ncu-rep --page details > gpurun_out/r02_kernels_ncu_details.txt 2>&1
gzip -f gpurun_out/r02_kernels_ncu_details.txt
rm -f gpurun_out/r02_kernels.ncu-rep
PROF_ITERS=1 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:fresco_attn" -c 1 -f -o gpurun_out/r02_attn_final python tools/prof_kernels.py > gpurun_out/r02_ncu_attn_final.log 2>&1
du -sh gpurun_out; tail -3 gpurun_out/r02_gputests.txt | cut -c1-300; cat gpurun_out/r02_attn_microbench_final.jsonl | cut -c1-400; cut -c1-400 gpurun_out/r02_bench.json
