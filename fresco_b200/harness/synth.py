"""Seeded synthetic inputs for tests / bench (scaffolding): smooth optical flows with
realistic occlusion rates (white-noise flow gives ~98 % occlusion and degenerates every
mask), frames, saliency, prompt embeddings."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import flow_utils


def smooth_flows(n: int, H: int, W: int, device, seed: int = 0, mag: float = 12.0):
    """Bidirectional flows for the n wrap-around frame pairs + occlusion masks from the
    forward/backward consistency check (gmflow/geometry.py:75-96)."""
    g = torch.Generator().manual_seed(seed)
    coarse = torch.randn(n, 2, 6, 6, generator=g) * mag * (H / 512.0)
    fwd = F.interpolate(coarse, size=(H, W), mode="bicubic", align_corners=False).to(device).contiguous()
    bwd = -flow_utils.flow_warp(fwd, -fwd)
    jitter = F.interpolate(torch.randn(n, 2, 6, 6, generator=g), size=(H, W), mode="bicubic").to(device)
    bwd = (bwd + 0.3 * jitter).contiguous()
    fwd_occ, bwd_occ = flow_utils.forward_backward_consistency_check(fwd, bwd)
    return [fwd, bwd], [fwd_occ, bwd_occ]


def frames(n: int, H: int, W: int, device, seed: int = 1):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(n, 3, H, W, generator=g) * 2 - 1).to(device)


def saliency(n: int, device, seed: int = 2, size: int = 256):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 1, size, size, generator=g).to(device)
