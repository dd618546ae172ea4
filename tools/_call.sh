mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -x -k "attention or processor or level_a" 2>&1 | tail -8 > gpurun_out/r02_twin3_parity.txt
timeout 300 python tools/bench_attn.py default FRESCO_ATTN_POLY=0 FRESCO_ATTN_POLY=8 FRESCO_ATTN_POLY=6 FRESCO_ATTN_POLY=5 FRESCO_ATTN_POLY=3 > gpurun_out/r02_attn_microbench_twin3.jsonl 2>&1
FRESCO_B200_LIB=$PWD/fresco_b200/libfresco_b200_nofold.so timeout 200 python tools/bench_attn.py default FRESCO_ATTN_POLY=0 > gpurun_out/r02_attn_microbench_twin3_nofold.jsonl 2>&1
PROF_ITERS=1 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:fresco_attn" -c 1 -f -o gpurun_out/r02_attn_twin3 python tools/prof_kernels.py > gpurun_out/r02_ncu_twin3.log 2>&1
tail -4 gpurun_out/r02_twin3_parity.txt | cut -c1-300; cat gpurun_out/r02_attn_microbench_twin3.jsonl gpurun_out/r02_attn_microbench_twin3_nofold.jsonl | cut -c1-330
