#!/bin/bash
# First GPU call of round 2 (run through tools/gpu.sh):  tools/gpu.sh 400 'bash tools/round2_first_call.sh'
#   1. parity of the experimental wide attention kernel (and every other knob combination)
#   2. golden set B (N = 3, 64 x 96, head_dim 80) through the kernels
#   3. level-B attention timing, wide off / on
# Everything is wrapped in its own timeout (the kernels trap on a lost barrier arrival instead of hanging).
mkdir -p gpurun_out
timeout 90 python tools/check_attn_knobs.py --wide 2>&1 | tail -30 | tee gpurun_out/r02_knob_parity.txt
timeout 90 python tools/check_set_b_gpu.py 2>&1 | tail -20 | tee gpurun_out/r02_set_b_gpu.txt
timeout 120 python tools/env_sweep.py FRESCO_ATTN_WIDE=0 FRESCO_ATTN_WIDE=1 2>&1 | tee gpurun_out/r02_wide_sweep.jsonl
