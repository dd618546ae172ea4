#!/usr/bin/env python
"""bench.py -- denoise-steps/sec of the FRESCO hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl fresco|reference]

A "step" is one UNet forward over an N=8-keyframe 512x512 batch (CFG => 16 samples) of an
SD-1.5-shaped random-init fp16 UNet with the FRESCO hooks installed through the reference's
own plug-in surface (apply_FRESCO_attn / apply_FRESCO_opt -> pipe.unet(...)):
  * FRESCO attention on the 6 decoder self-attention layers of up_blocks.2/3
    (cross-frame always; spatial-guided on schedule step 0; temporal-guided while t >= 350),
  * warp_tensor background fusion on the 4 decoder features on the optimisation steps
    (BASELINE configs[1]: "full FRESCO attn + flow warp", optimisation off).
Steps walk the reference's 15-step schedule cyclically.

Output: ONE JSON line (rank 0).  `value` = device-resident inputs; `e2e` = same call with
pinned-host inputs/outputs copied inside the timed region; `roofline` = the dominant kernel
(cross-frame attention, level B) timed per launch with CUDA events; `cpu_baseline` = the CPU
oracle on a bounded sample.  With --gpus N > 1 (torchrun) every rank runs an independent
8-keyframe batch (the reference's own batching unit): weak scaling, no data-path collective.
Other workloads (never the headline): --workload config3 adds the feature optimisation,
--workload config4 frame-shards ONE N=16 batch over the ranks (K/V all-gather per layer,
strong scaling).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FRAMES = 8
RES = 512
CHUNKS = 2
TIMESTEPS = [950 - 50 * i for i in range(20)]          # DDPMScheduler.set_timesteps(20), "leading"
WARMUP_STEPS = 5                                        # SDEdit warm-up: run timesteps[5:]
SCHEDULE = TIMESTEPS[WARMUP_STEPS:]                     # 15 steps: 700 .. 0
OPT_STEPS = TIMESTEPS[WARMUP_STEPS:15]                  # optimisation / warp active on t in 700..250
METRIC = "denoise-steps/sec"
UNIT = "steps/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))),
                "tflops_burst": float(d.get("bf16_tflops", 1590.0)), "hbm_gbs": float(d.get("hbm_gbs", 6650.0)),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed
    `ncu --set full` capture of this same workload (profiles/r01_attn_traffic.json); None if absent."""
    p = os.path.join(ROOT, "profiles", "r01_attn_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return {"bytes_per_launch": d["traffic_bytes_per_launch"], "unit": "B", "source": d["source"]}
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.lines = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# workload construction (fresco arm)
# --------------------------------------------------------------------------------------------
class Workload:
    def __init__(self, device, seed=0, n_frames=N_FRAMES, res=RES, optimise=False, shard=None):
        """shard = (world, rank): frame-sharded batch (config 4); n_frames is then the GLOBAL frame count and this
        rank feeds frames [rank*n/world, (rank+1)*n/world) of both CFG chunks through the UNet."""
        from fresco_b200 import diffusion_hacked as dh
        from fresco_b200 import flow_utils as fu
        from fresco_b200.harness import synth
        from fresco_b200.harness.sd15_unet import FakePipe, SD15UNet
        self.dh = dh
        self.device = device
        self.n = n_frames
        torch.manual_seed(seed)
        unet = SD15UNet().to(device=device, dtype=torch.float16).eval()
        self.pipe = FakePipe(unet)
        self.shard = shard
        if shard is not None:
            world, rk = shard
            self.lo, self.hi = rk * n_frames // world, (rk + 1) * n_frames // world
        else:
            self.lo, self.hi = 0, n_frames
        n_local = self.hi - self.lo
        self.proc = dh.apply_FRESCO_attn(self.pipe, shard=None if shard is None else (shard[0], shard[1], None))
        ctrl = self.proc.controller
        ctrl.disable_controller()
        dh.disable_FRESCO_opt(self.pipe)
        lat = res // 8
        g = torch.Generator().manual_seed(seed + 1)
        self.latents_host = torch.randn(n_frames, 4, lat, lat, generator=g)[self.lo:self.hi].half().pin_memory()
        self.prompt_host = torch.randn(CHUNKS * n_local, 77, 768, generator=g).half().pin_memory()
        self.out_host = torch.empty(CHUNKS * n_local, 4, lat, lat, dtype=torch.float16).pin_memory()
        self.latents = self.latents_host.to(device)
        self.prompt = self.prompt_host.to(device)
        # per-batch FRESCO parameters (what get_flow_and_interframe_paras produces, with synthetic flow)
        self.flows, self.occs = synth.smooth_flows(n_frames, res, res, device, seed=seed + 2)
        imgs = synth.frames(n_frames, res, res, device, seed + 3)
        self.saliency = synth.saliency(n_frames, device, seed + 4)
        self.attn_mask = dh.cross_frame_attn_masks(self.occs[1])
        fm, bm, im = [], [], []
        for scale in (8.0, 16.0):
            f, b, m = fu.get_mapping_ind(self.flows[1], self.occs[1], imgs, scale=scale)
            fm.append(f), bm.append(b), im.append(m)
        self.interattn_paras = {"fwd_mappings": fm, "bwd_mappings": bm, "interattn_masks": im}
        self.kv_len = {int(m.shape[1]): int(m.sum().item()) for m in self.attn_mask}
        # reference pass: store the 6 decoder self-attention inputs (get_intraframe_paras, store=True)
        ctrl.clear_store()
        ctrl.enable_store()
        with torch.no_grad():
            self.pipe.unet(torch.cat([self.latents] * 2), TIMESTEPS[-1], encoder_hidden_states=self.prompt,
                           return_dict=False)
        ctrl.disable_store()
        if shard is not None:
            return                         # config 4: attention only (warp_tensor is a sequential frame chain)
        if optimise:
            # BASELINE configs[2]: FRESCO feature optimisation (20 Adam iterations, temporal + Gram-L1 loss) on the 4
            # decoder features on the optimisation steps, Gram targets from the reference pass (get_intraframe_paras)
            with torch.no_grad():
                feats = self.pipe.unet(torch.cat([self.latents] * 2), TIMESTEPS[-1], encoder_hidden_states=self.prompt,
                                       return_dict=False)[1:]
            self.gram = dh.gram_targets(feats)
            dh.apply_FRESCO_opt(self.pipe, steps=OPT_STEPS, flows=self.flows, occs=self.occs,
                                correlation_matrix=self.gram, intra_weight=1e2, iters=20, optimize_temporal=True,
                                saliency=self.saliency)
        else:
            # BASELINE configs[1]: warp-only decoder fusion on the optimisation steps (optimize_feature early-outs:
            # no Gram targets, optimize_temporal=False)
            dh.apply_FRESCO_opt(self.pipe, steps=OPT_STEPS, flows=self.flows, occs=self.occs, correlation_matrix=[],
                                optimize_temporal=False, saliency=self.saliency)
    def set_schedule_state(self, i):
        """flag toggles of pipe_FRESCO.inference (src/pipe_FRESCO.py:171-174)"""
        ctrl = self.proc.controller
        if i == 0:
            ctrl.enable_controller(interattn_paras=self.interattn_paras, attn_mask=self.attn_mask)
        if i >= 1:
            ctrl.disable_intraattn()
        if SCHEDULE[i] < 350:
            ctrl.disable_interattn()

    @torch.no_grad()
    def step(self, k, host_io=False):
        i = k % len(SCHEDULE)
        self.set_schedule_state(i)
        if host_io:
            lat = self.latents_host.to(self.device, non_blocking=True)
            prompt = self.prompt_host.to(self.device, non_blocking=True)
        else:
            lat, prompt = self.latents, self.prompt
        out = self.pipe.unet(torch.cat([lat] * 2), SCHEDULE[i], encoder_hidden_states=prompt, return_dict=False)[0]
        if host_io:
            self.out_host.copy_(out, non_blocking=True)
        return out


def dist_setup(gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    return world, rank, local


def timed_region(wl, steps, warmup, host_io, world):
    import torch.distributed as dist
    for k in range(warmup):
        wl.step(k, host_io)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        wl.step(k, host_io)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


# --------------------------------------------------------------------------------------------
# CPU baseline / reference arm (oracle port; see oracle/ref_step.py)
# --------------------------------------------------------------------------------------------
def cpu_reference_line(args, as_reference_arm):
    from oracle import ref_step
    cores = ref_step.pick_threads(os.cpu_count() or 1)
    steps = args.steps if as_reference_arm else 1
    warm = args.warmup if as_reference_arm else 0
    res = ref_step.run(n_full=N_FRAMES, n_sample=args.cpu_sample_frames, res=RES, schedule=SCHEDULE,
                       opt_steps=OPT_STEPS, steps=steps, warmup=warm, max_seconds=args.cpu_budget_s)
    cb = {"value": res["steps_per_s"], "unit": UNIT, "cores": cores, "kind": "port", "sample": res["sample"]}
    return cb, res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=15)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fresco", choices=["fresco", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-frames", type=int, default=2)
    ap.add_argument("--cpu-budget-s", type=float, default=150.0)
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "config4"],
                    help="config2 (default, the headline): FRESCO attention + warp fusion; config3: + feature "
                         "optimisation (apply_FRESCO_opt with Gram targets, 20 Adam iterations on 10 of 15 steps); "
                         "config4: ONE N=16 keyframe batch frame-sharded over the ranks, K/V all-gather per FRESCO "
                         "layer (strong scaling; attention only, no warp fusion)")
    ap.add_argument("--profile-mode", action="store_true",
                    help="for ncu captures only: 1 warm-up + --steps, no e2e / cpu baseline; never a bench value")
    args = ap.parse_args()
    if args.impl == "fresco" and not args.profile_mode:
        args.warmup = max(args.warmup, 3)

    if args.impl == "reference" and int(os.environ.get("RANK", "0")) != 0:
        return                                   # under torchrun only rank 0 runs (and prints) the CPU arm
    world, rank, local = (1, 0, 0) if args.impl == "reference" else dist_setup(args.gpus)
    config = {"workload": "N=8 keyframes 512x512 (CFG batch 16), SD1.5-shaped random-init fp16 UNet, FRESCO "
                          "attention (cross-frame + spatial step 0 + temporal t>=350) on 6 decoder layers + "
                          "warp_tensor fusion on 4 decoder features; 15-step DDPM schedule t=700..0 walked cyclically",
              "frames": N_FRAMES, "resolution": RES, "parallelism": "replicas x%d (one 8-keyframe batch per GPU)" % world,
              "l2": "working set (1.7 GB weights + activations) >> 126 MB L2, no explicit flush"}

    if args.impl == "reference":
        config["parallelism"] = "host cores only (rank 0); n_gpus echoes --gpus"
        cb, res = cpu_reference_line(args, True)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": res["steps"], "warmup": res["warmup"], "ms_per_step": 1000.0 / cb["value"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "detail": res["detail"]}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl fresco needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    from fresco_b200 import _lib, ops
    _lib.lib()
    if args.workload == "config4":
        # every rank must build the same per-batch parameters (same seed); only the frame slice differs
        wl = Workload(device, seed=0, n_frames=16, shard=(world, rank))
        config.update({"workload": "ONE batch of N=16 keyframes 512x512 (CFG batch 32) frame-sharded over the ranks, SD1.5-"
                       "shaped random-init fp16 UNet, FRESCO attention on 6 decoder layers with one NCCL all-gather of "
                       "the compacted K/V per layer (+ q/k/out gathers while temporal-guided attention is on)",
                       "frames": 16, "parallelism": "frame-sharded x%d" % world})
    else:
        wl = Workload(device, seed=rank, optimise=args.workload == "config3")
    if args.workload == "config3":
        config["workload"] += "; + optimize_feature (20 Adam iters, temporal + Gram-L1) on 4 decoder features, 10 of 15 steps"

    # ---- value: inputs resident in HBM
    sampler = ClockSampler(local)
    ops.PROFILE = []
    launches0 = _lib.launch_count()
    sampler.start()
    ms = timed_region(wl, args.steps, args.warmup, False, world)
    clocks = sampler.stop()
    warm_launch_entries = None
    launches = _lib.launch_count() - launches0
    prof = ops.PROFILE
    ops.PROFILE = None
    # launches counted above include the warm-up steps: rescale to the timed steps only
    gpu_launches = int(round(launches * args.steps / float(args.steps + args.warmup)))
    if args.profile_mode:
        print(json.dumps({"profile_mode": True, "ms_per_step": ms / args.steps, "note": "not a bench value"}))
        return
    # ---- e2e: pinned-host inputs / outputs copied inside the timed region
    ms_e2e = timed_region(wl, args.steps, 1, True, world)

    jobs = 1 if args.workload == "config4" else world      # config 4: all ranks work on the same batch
    value = jobs * args.steps / (ms / 1000.0)
    e2e_value = jobs * args.steps / (ms_e2e / 1000.0)
    h2d = wl.latents_host.numel() * 2 + wl.prompt_host.numel() * 2
    d2h = wl.out_host.numel() * 2

    # ---- roofline of the dominant kernel: cross-frame attention at level B (L=4096, d=40)
    peaks = load_peaks()
    L_b = (RES // 8) ** 2
    n_q = 2 * (wl.hi - wl.lo)
    by_tag = {}
    for tag, work, a, b in prof:
        by_tag.setdefault(tag, []).append((work, a.elapsed_time(b)))
    dom_tag = "attn_d40_L%d_Lk%d" % (L_b, wl.kv_len[L_b])
    roof = None
    share = None
    if dom_tag in by_tag:
        ent = by_tag[dom_tag]
        avg_ms = sum(t for _, t in ent) / len(ent)
        flops = ent[0][0]
        ach = flops / (avg_ms * 1e-3) / 1e12
        total_attn_ms = sum(t for e in by_tag.values() for _, t in e)
        share = sum(t for _, t in ent) / ((args.steps + args.warmup) / args.steps * ms)
        roof = {"kernel": "fresco_attn_kernel<40> (cross-frame, L=%d, Lk=%d, B=%d, 8 heads)" % (L_b, wl.kv_len[L_b], n_q),
                "bound": "tensor", "achieved": round(ach, 2), "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": round(ach / peaks["tflops"], 4), "traffic": load_traffic(), "peak_source": peaks["source"],
                "algorithmic_flops_per_launch": flops, "avg_launch_ms": round(avg_ms, 4),
                "launches_timed": len(ent), "share_of_step_time": round(share, 4),
                "all_attention_ms_per_step": round(total_attn_ms / (args.steps + args.warmup), 3)}

    line = {"metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "strong" if args.workload == "config4" else "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": round(e2e_value, 4), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": gpu_launches, "roofline": roof,
            "kv_len": wl.kv_len}
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            try:
                cb, _ = cpu_reference_line(args, False)
                line["cpu_baseline"] = cb
            except Exception as e:  # the baseline is a reported extra; never hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": "failed: %r" % (e,)}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
