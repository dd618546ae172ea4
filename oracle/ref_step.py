"""TEST / BENCH INFRASTRUCTURE ONLY -- the CPU arm of bench.py.

Runs the denoise step of bench.py's workload with the *oracle* (the CPU restatement of the
reference's FRESCO path, oracle/fresco_oracle.py) on the host cores, fp32, all threads:
the same SD-1.5-shaped harness UNet (random init), the oracle attention on the 6 decoder
self-attention layers and the oracle warp_tensor on the 4 decoder features.

A full N=8 step takes minutes on a CPU, so each step is a BOUNDED SAMPLE: ``n_sample`` of the
``n_full`` keyframes.  The three components are timed separately and scaled to the full
workload by their algorithmic work (UNet body and warp: linear in frames; attention:
4*B*L*Lk*C with the real K/V lengths of the N=8 masks).  bench.py reports the result as the
``cpu_baseline`` / ``--impl reference`` line and says what the sample was.
"""
from __future__ import annotations

import time
from typing import Dict, List

import torch

from . import fresco_oracle as O


class _Timer:
    def __init__(self):
        self.t: Dict[str, float] = {}

    def add(self, key, dt):
        self.t[key] = self.t.get(key, 0.0) + dt

    def reset(self):
        self.t = {}


class OracleProcessor:
    """diffusers-style attn processor that evaluates oracle.fresco_attention."""

    def __init__(self, state, timer, layer_stride=1):
        self.s = state
        self.timer = timer
        self.ref_index = 0
        self.layer_stride = layer_stride      # evaluate 1 of `layer_stride` FRESCO layers per level (time scaled back)
        self.calls = {}

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        if encoder_hidden_states is not None:
            from fresco_b200.harness.sd15_unet import PlainProcessor
            return PlainProcessor()(attn, hidden_states, encoder_hidden_states, attention_mask)
        s = self.s
        L = hidden_states.shape[1]
        c = self.calls.get(L, 0)
        self.calls[L] = c + 1
        if c % self.layer_stride != 0:
            if s["intra"]:
                self.ref_index += 1
            return attn.to_out[0](attn.to_v(hidden_states))        # skipped layer of the bounded sample
        t0 = time.perf_counter()
        ref = None
        if s["intra"]:
            ref = s["ref"][self.ref_index % len(s["ref"])]
            self.ref_index += 1
        out = O.fresco_attention(
            hidden_states, attn.to_q.weight, attn.to_k.weight, attn.to_v.weight, attn.to_out[0].weight,
            attn.to_out[0].bias, attn.heads, use_cfattn=True, attn_masks=s["attn_masks"],
            use_intraattn=s["intra"], ref_hidden=ref, use_interattn=s["inter"], fwd_mappings=s["fwd"],
            bwd_mappings=s["bwd"], interattn_masks=s["imask"])
        self.timer.add("attn_L%d" % L, (time.perf_counter() - t0) * self.layer_stride)
        return out


def _attn_flops(B, L, Lk, C, intra):
    return 4.0 * B * L * Lk * C + (4.0 * B * L * L * C if intra else 0.0)


@torch.no_grad()
def run(n_full: int, n_sample: int, res: int, schedule: List[int], opt_steps: List[int], steps: int, warmup: int,
        max_seconds: float = 150.0, layer_stride: int = 3):
    from fresco_b200.harness.sd15_unet import SD15UNet      # scaffolding UNet body (not the product path)
    torch.manual_seed(0)
    unet = SD15UNet().float().eval()
    timer = _Timer()
    lat = res // 8
    n = n_sample
    flows, occs = O.synth_flows(n, res, res, seed=2)
    imgs = torch.rand(n, 3, res, res) * 2 - 1
    sal = torch.rand(n, 1, 256, 256)
    masks = O.cross_frame_masks(occs[1])
    fwd, bwd, imask = [], [], []
    for sc in (8.0, 16.0):
        f, b, m = O.mapping_ind(flows[1], occs[1], imgs, sc)
        fwd.append(f), bwd.append(b), imask.append(m)
    # K/V lengths of the FULL workload (masks only; cheap)
    _, occs_full = O.synth_flows(n_full, res, res, seed=2)
    masks_full = O.cross_frame_masks(occs_full[1])
    lk_full = {int(m.shape[1]): int(m.sum()) for m in masks_full}
    lk_samp = {int(m.shape[1]): int(m.sum()) for m in masks}
    state = {"attn_masks": masks, "fwd": fwd, "bwd": bwd, "imask": imask, "intra": False, "inter": False,
             "ref": [torch.randn(2 * n, (lat // 2) ** 2, 640)] * 3 + [torch.randn(2 * n, lat ** 2, 320)] * 3}
    proc = OracleProcessor(state, timer, layer_stride)
    table = {}
    for name in unet.attn_processors:
        table[name] = proc if name.startswith(("up_blocks.2", "up_blocks.3")) else unet.attn_processors[name]
    unet.set_attn_processor(table)
    warp_on = {"on": False}

    def pre_hook(module, a, kw):
        if not warp_on["on"]:
            return None
        t0 = time.perf_counter()
        kw = dict(kw)
        kw["hidden_states"] = O.warp_tensor(kw["hidden_states"], flows, occs, sal, 2)
        timer.add("warp", time.perf_counter() - t0)
        return a, kw

    for blk in unet.up_blocks:
        blk.register_forward_pre_hook(pre_hook, with_kwargs=True)
    latents = torch.randn(n, 4, lat, lat)
    prompt = torch.randn(2 * n, 77, 768)
    ratio_frames = n_full / float(n_sample)
    scaled_times, raw_times, detail = [], [], []
    t_begin = time.perf_counter()
    done = 0
    for k in range(warmup + steps):
        i = k % len(schedule)
        state["intra"] = (i == 0)
        state["inter"] = schedule[i] >= 350
        proc.ref_index = 0
        proc.calls = {}
        warp_on["on"] = schedule[i] in opt_steps
        timer.reset()
        t0 = time.perf_counter()
        unet(torch.cat([latents] * 2), schedule[i], encoder_hidden_states=prompt, return_dict=False)
        total = time.perf_counter() - t0
        comp = dict(timer.t)
        body = total - comp.get("warp", 0.0) - sum(v for kk, v in comp.items() if kk.startswith("attn_L")) / layer_stride
        scaled = body * ratio_frames + comp.get("warp", 0.0) * ratio_frames
        for key, t in comp.items():
            if key.startswith("attn_L"):
                L = int(key[6:])
                C = 320 if L == lat * lat else 640
                r = _attn_flops(2 * n_full, L, lk_full[L], C, state["intra"]) / _attn_flops(2 * n, L, lk_samp[L], C,
                                                                                             state["intra"])
                scaled += t * r
        if k >= warmup:
            scaled_times.append(scaled)
            raw_times.append(total)
            detail.append({"t": schedule[i], "sample_s": round(total, 3), "scaled_s": round(scaled, 3),
                           "body_s": round(body, 3), **{kk: round(v, 3) for kk, v in comp.items()}})
            done += 1
        if time.perf_counter() - t_begin > max_seconds and done >= 1:
            break
    mean_scaled = sum(scaled_times) / len(scaled_times)
    return {"steps_per_s": 1.0 / mean_scaled, "steps": done, "warmup": min(warmup, k),
            "raw_ms_per_step": 1000.0 * sum(raw_times) / len(raw_times),
            "sample": ("%d of %d keyframes per step at 512x512 (CFG batch %d), 1 of %d FRESCO attention layers per "
                       "level evaluated (time x%d), oracle fp32 on %d threads, %d step(s); "
                       "component times scaled to N=%d by algorithmic work (UNet body, warp: x%.1f; attention: "
                       "4*B*L*Lk*C with Lk=%s); mean raw sample step %.2f s" %
                       (n_sample, n_full, 2 * n_sample, layer_stride, layer_stride, torch.get_num_threads(), done,
                        n_full, ratio_frames,
                        lk_full, sum(raw_times) / len(raw_times))),
            "detail": detail}
