mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 15 --warmup 3 > gpurun_out/r02_c6_bench_g2.json 2> gpurun_out/r02_c6_bench_g2.err
grep -n "Error\|error" gpurun_out/r02_c6_bench_g2.err | grep -v "torch/distributed" | head -5
python - <<'PY'
import json
for f in ("gpurun_out/r02_c6_bench_g2.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["config"]["execution"], d.get("sharded_check"), "launches", d["gpu_launches"])
    for k, v in d.get("kernels", {}).items():
        print("  %-34s %-6s ms %-8s ach %-8s frac %-6s n %s" % (k, v["bound"], v["ms"], v["achieved"], v["frac"], v["launches"]))
PY
