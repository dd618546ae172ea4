mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_c1_smi.txt
timeout 600 python -m pytest tests/test_gpu_parity_r2.py -q -k "attention_variants or level_a" 2>&1 | tail -40 > gpurun_out/r02_c1_attn_parity.txt
timeout 300 python tools/bench_attn.py --sdpa FRESCO_ATTN_WIDE=0 FRESCO_ATTN_WIDE=1 FRESCO_ATTN_WIDE=1,FRESCO_ATTN_POLY=8 FRESCO_ATTN_WIDE=1,FRESCO_ATTN_POLY=4 FRESCO_ATTN_WIDE=0,FRESCO_ATTN_POLY=4 > gpurun_out/r02_c1_bench_attn.jsonl 2>&1
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity_r2.py::test_attention_variants_all_head_dims 2>&1 | tail -60 > gpurun_out/r02_c1_pytest_all.txt
tail -5 gpurun_out/r02_c1_attn_parity.txt; cat gpurun_out/r02_c1_bench_attn.jsonl; tail -15 gpurun_out/r02_c1_pytest_all.txt
