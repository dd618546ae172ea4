mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity_r2.py -q -x -k "attention_variants or level_a or processor" 2>&1 | tail -30 > gpurun_out/r02_c8_attn_parity.txt
timeout 200 python tools/bench_attn.py default FRESCO_ATTN_WIDE=0 FRESCO_ATTN_WIDE=2 FRESCO_ATTN_WIDE=4 --sdpa > gpurun_out/r02_c8_bench_attn.jsonl 2>&1
PROF_ITERS=1 timeout 700 ncu --set full --clock-control none --import-source on -k "regex:fresco_attn|temporal_attn|warp_|gram|kv_compact|adam|adain|gmflow|dilate|project" -c 60 -f -o gpurun_out/r02_kernels python tools/prof_kernels.py > gpurun_out/r02_c8_ncu.log 2>&1
tail -5 gpurun_out/r02_c8_attn_parity.txt | cut -c1-300; cat gpurun_out/r02_c8_bench_attn.jsonl; tail -5 gpurun_out/r02_c8_ncu.log; ls -la gpurun_out/r02_kernels.ncu-rep
