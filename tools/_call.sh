mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity_r2.py::test_attention_variants_all_head_dims 2>&1 | tail -60 > gpurun_out/r02_c3_pytest_all.txt
timeout 900 python bench.py --steps 15 --warmup 3 > gpurun_out/r02_c3_bench.json 2> gpurun_out/r02_c3_bench.err
timeout 600 python bench.py --steps 15 --warmup 3 --workload config3 --no-cpu-baseline --no-extras > gpurun_out/r02_c3_bench_config3.json 2> gpurun_out/r02_c3_bench_config3.err
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02_c3_pytest_all.txt; tail -3 gpurun_out/r02_c3_bench.err; python - <<'PY'
import json
for f in ("gpurun_out/r02_c3_bench.json", "gpurun_out/r02_c3_bench_config3.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
    print(" roofline", json.dumps(d.get("roofline"))[:600])
    for k, v in d.get("kernels", {}).items():
        print("  %-34s %-6s ms %-8s ach %-8s frac %-6s n %s" % (k, v["bound"], v["ms"], v["achieved"], v["frac"], v["launches"]))
    print(" eager", json.dumps(d.get("gpu_eager_baseline"))[:1200])
    print(" cpu", json.dumps(d.get("cpu_baseline"))[:300])
PY
