mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -x -k "WIDE6" 2>&1 | tail -8 > gpurun_out/r02_twin16_parity.txt
timeout 400 python tools/bench_attn.py default FRESCO_ATTN_WIDE=6 FRESCO_ATTN_WIDE=6,FRESCO_ATTN_POLY=0 FRESCO_ATTN_WIDE=6,FRESCO_ATTN_POLY=8 FRESCO_ATTN_WIDE=5 FRESCO_ATTN_WIDE=4 > gpurun_out/r02_attn_microbench_twin16.jsonl 2>&1
FRESCO_ATTN_WIDE=6 PROF_ITERS=1 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:fresco_attn" -c 1 -f -o gpurun_out/r02_attn_twin16 python tools/prof_kernels.py > gpurun_out/r02_ncu_twin16.log 2>&1
tail -4 gpurun_out/r02_twin16_parity.txt | cut -c1-300; cat gpurun_out/r02_attn_microbench_twin16.jsonl | cut -c1-400
