mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -x -k "WIDE5" 2>&1 | tail -15 > gpurun_out/r02_twin_parity.txt
timeout 300 python tools/bench_attn.py default FRESCO_ATTN_WIDE=5 FRESCO_ATTN_WIDE=5,FRESCO_ATTN_POLY=4 FRESCO_ATTN_WIDE=5,FRESCO_ATTN_POLY=8 FRESCO_ATTN_WIDE=3 FRESCO_ATTN_WIDE=3,FRESCO_ATTN_POLY=4 FRESCO_ATTN_WIDE=0,FRESCO_ATTN_POLY=4 --sdpa > gpurun_out/r02_attn_microbench_twin.jsonl 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -k "not WIDE5" 2>&1 | tail -15 > gpurun_out/r02_gputests.txt
FRESCO_ATTN_WIDE=5 PROF_ITERS=1 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:fresco_attn" -c 1 -f -o gpurun_out/r02_attn_twin python tools/prof_kernels.py > gpurun_out/r02_ncu_twin.log 2>&1
tail -6 gpurun_out/r02_twin_parity.txt | cut -c1-300; cat gpurun_out/r02_attn_microbench_twin.jsonl | cut -c1-420; tail -6 gpurun_out/r02_gputests.txt | cut -c1-300; du -sh gpurun_out
