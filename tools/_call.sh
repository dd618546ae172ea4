mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_r2.py -q -k "attention_variants or level_a" 2>&1 | tail -40 > gpurun_out/r02_c4_attn_parity.txt
timeout 300 python tools/bench_attn.py FRESCO_ATTN_WIDE=0 FRESCO_ATTN_WIDE=2 FRESCO_ATTN_WIDE=4 FRESCO_ATTN_WIDE=2,FRESCO_ATTN_POLY=4 FRESCO_ATTN_WIDE=4,FRESCO_ATTN_POLY=4 > gpurun_out/r02_c4_bench_attn.jsonl 2>&1
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity_r2.py::test_attention_variants_all_head_dims 2>&1 | tail -60 > gpurun_out/r02_c4_pytest_all.txt
timeout 300 python tools/bench_opt.py > gpurun_out/r02_c4_bench_opt.txt 2>&1
tail -5 gpurun_out/r02_c4_attn_parity.txt; cat gpurun_out/r02_c4_bench_attn.jsonl; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02_c4_pytest_all.txt; tail -6 gpurun_out/r02_c4_bench_opt.txt | cut -c1-900
