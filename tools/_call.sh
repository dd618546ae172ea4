mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_c1_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -k "not WIDE3" 2>&1 | tail -25 > gpurun_out/r02_gputests.txt
if grep -q "failed\|error" gpurun_out/r02_gputests.txt; then
  echo "ELECT BUILD FAILED TESTS: lane0 build for the rest of this call" | tee gpurun_out/r02_fallback.txt
  FRESCO_B200_LIB=$PWD/fresco_b200/libfresco_b200_lane0.so timeout 900 python -m pytest tests -m gpu -q -x -k "not WIDE3" 2>&1 | tail -25 > gpurun_out/r02_gputests_lane0.txt
  export FRESCO_B200_LIB=$PWD/fresco_b200/libfresco_b200_lane0.so
fi
timeout 300 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -k "WIDE3" 2>&1 | tail -25 > gpurun_out/r02_duo_parity.txt
timeout 200 python tools/bench_attn.py default FRESCO_ATTN_WIDE=0 FRESCO_ATTN_WIDE=2 FRESCO_ATTN_WIDE=4 --sdpa > gpurun_out/r02_attn_microbench.jsonl 2>&1
timeout 200 python tools/bench_attn.py FRESCO_ATTN_WIDE=3 FRESCO_ATTN_WIDE=3,FRESCO_ATTN_POLY=4 FRESCO_ATTN_WIDE=3,FRESCO_ATTN_POLY=8 FRESCO_ATTN_WIDE=0,FRESCO_ATTN_POLY=4 > gpurun_out/r02_attn_microbench_duo.jsonl 2>&1
FRESCO_B200_LIB=$PWD/fresco_b200/libfresco_b200_lane0.so timeout 200 python tools/bench_attn.py default FRESCO_ATTN_WIDE=0 FRESCO_ATTN_WIDE=3 > gpurun_out/r02_attn_microbench_lane0.jsonl 2>&1
timeout 500 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --profile-mode --steps 1 --warmup 1 > gpurun_out/r02_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_summary.txt 2>&1
gzip -f gpurun_out/r02_launches.csv
PROF_ITERS=1 timeout 700 ncu --set full --clock-control none -k "regex:fresco_attn|temporal_attn|warp_|gram|kv_compact|adam|adain|gmflow|dilate|project" -c 40 -f -o gpurun_out/r02_kernels python tools/prof_kernels.py > gpurun_out/r02_ncu.log 2>&1
python tools/ncu_table.py gpurun_out/r02_kernels.ncu-rep > gpurun_out/r02_kernels_ncu.txt 2>&1
ncu -i gpurun_out/r02_kernels.ncu-rep --page details > gpurun_out/r02_kernels_ncu_details.txt 2>&1
gzip -f gpurun_out/r02_kernels_ncu_details.txt
rm -f gpurun_out/r02_kernels.ncu-rep
FRESCO_ATTN_WIDE=3 PROF_ITERS=1 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:fresco_attn" -c 2 -f -o gpurun_out/r02_attn_duo python tools/prof_kernels.py > gpurun_out/r02_ncu_duo.log 2>&1
FRESCO_ATTN_WIDE=0 PROF_ITERS=1 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:fresco_attn" -c 1 -f -o gpurun_out/r02_attn_pipe python tools/prof_kernels.py > gpurun_out/r02_ncu_pipe.log 2>&1
du -sh gpurun_out; cat gpurun_out/r02_fallback.txt 2>/dev/null; tail -5 gpurun_out/r02_gputests.txt | cut -c1-300; tail -8 gpurun_out/r02_duo_parity.txt | cut -c1-300; cat gpurun_out/r02_attn_microbench.jsonl gpurun_out/r02_attn_microbench_duo.jsonl gpurun_out/r02_attn_microbench_lane0.jsonl | cut -c1-600; cut -c1-1200 gpurun_out/r02_bench.json; tail -3 gpurun_out/r02_bench.err; cat gpurun_out/r02_kernels_ncu.txt | cut -c1-250; ls -la gpurun_out/
