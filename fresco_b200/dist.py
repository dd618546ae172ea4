"""Frame-sharded FRESCO attention for one multi-GPU box (NCCL over NVLink / NVSwitch).

The reference has no collective on this path; this module is the B200 design for
BASELINE config 4: the N keyframes of ONE batch are split contiguously over G ranks
(rank r owns frames [r*N/G, (r+1)*N/G) of both CFG chunks) and the UNet body runs as plain
replicas.  Only FRESCO attention couples frames:

* cross-frame attention (src/diffusion_hacked.py:225-247, :303-305): every query frame attends the
  same K/V set = all tokens of frame 0 + the occluded tokens of frames 1..N-1.  Each rank compacts
  the selected rows of ITS frames, and ONE all-gather per layer (K and V packed in one padded buffer;
  per-rank counts are known on the host from the per-batch mask) gives every rank the full set in
  the reference's row-major (frame, token) order -- rank order == frame order.
* temporal-guided attention (:320-367) mixes all N frames along each trajectory: a second
  all-gather (q, k, attention output) while it is enabled (t >= 350), then each rank keeps its frames.
* spatial-guided attention (:257-288) is per frame: local.

Compute goes through a small backend object (default: the CUDA kernels in ``ops``); the CPU
tests inject a torch backend to exercise the exchange / ordering logic under gloo.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def frame_range(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    if n_frames % world != 0:
        raise ValueError(f"{n_frames} frames do not split evenly over {world} ranks")
    per = n_frames // world
    return rank * per, (rank + 1) * per


def _all_gather_stacked(x: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """all-gather equal-shaped tensors; returns [world, *x.shape] (rank-major)"""
    x = x.contiguous()
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x, group=group)
    return out.view((world,) + tuple(x.shape))


class _OpsBackend:
    """the product path: hand-written sm_100a kernels"""

    def kv_compact(self, k, v, idx, chunks):
        from . import ops
        return ops.kv_compact(k, v, idx, chunks)

    def attn_fwd(self, q, k, v, heads, q_per_kv, softmax_scale, diag_bias=0.0):
        from . import ops
        return ops.attn_fwd(q, k, v, heads, q_per_kv, softmax_scale=softmax_scale, diag_bias=diag_bias)

    def temporal_attn_fwd(self, q, k, v, fwd_map, traj_mask, chunks, heads, scale):
        from . import ops
        return ops.temporal_attn_fwd(q, k, v, fwd_map, traj_mask, chunks, heads, scale)


class ShardedKVExchange:
    """Per-batch plan + per-layer all-gather of the compacted cross-frame K/V."""

    def __init__(self, attn_mask: torch.Tensor, world: int, rank: int, group=None):
        n_frames, tokens = attn_mask.shape
        self.world, self.rank, self.group = world, rank, group
        self.lo, self.hi = frame_range(n_frames, world, rank)
        counts = [int(attn_mask[slice(*frame_range(n_frames, world, r))].sum().item()) for r in range(world)]
        self.counts = counts
        self.max_count = max(max(counts), 1)
        self.total = sum(counts)
        local = attn_mask[self.lo:self.hi].reshape(-1)
        self.local_idx = torch.nonzero(local, as_tuple=False).reshape(-1).to(torch.int32).contiguous()

    def exchange(self, k_local: torch.Tensor, v_local: torch.Tensor, chunks: int, backend) -> Tuple[torch.Tensor, torch.Tensor]:
        """k_local, v_local [chunks * N_local, L, C] -> k_sel, v_sel [chunks, Lk_global, C]"""
        C = k_local.shape[-1]
        cnt = self.counts[self.rank]
        send = torch.zeros(chunks, self.max_count, 2 * C, dtype=k_local.dtype, device=k_local.device)
        if cnt > 0:
            ks, vs = backend.kv_compact(k_local, v_local, self.local_idx, chunks)
            send[:, :cnt, :C] = ks
            send[:, :cnt, C:] = vs
        if self.world == 1:
            gathered = send.unsqueeze(0)
        else:
            gathered = _all_gather_stacked(send, self.world, self.group)
        parts = [gathered[r, :, :self.counts[r]] for r in range(self.world)]       # rank order == frame order
        full = torch.cat(parts, dim=1)
        return full[..., :C].contiguous(), full[..., C:].contiguous()


def _gather_frames(x_local: torch.Tensor, chunks: int, world: int, group=None) -> torch.Tensor:
    """[chunks * N_local, ...] on every rank -> [chunks * N, ...] in the reference's (b f) order"""
    if world == 1:
        return x_local
    out = _all_gather_stacked(x_local, world, group)
    n_local = x_local.shape[0] // chunks
    out = out.view(world, chunks, n_local, *x_local.shape[1:]).transpose(0, 1)
    return out.reshape(chunks * world * n_local, *x_local.shape[1:]).contiguous()


class ShardedFRESCOAttention:
    """FRESCO self-attention of one layer on a frame shard (the sharded counterpart of
    ``FRESCOAttnProcessor2_0._fresco_self_attention``)."""

    def __init__(self, controller, world: int, rank: int, chunks: int = 2, group=None, backend=None):
        self.ctrl, self.world, self.rank, self.chunks, self.group = controller, world, rank, chunks, group
        self.backend = backend if backend is not None else _OpsBackend()
        self._plans = {}

    def _plan(self, tokens: int) -> Optional[ShardedKVExchange]:
        """Exchange plan for the mask with ``shape[1] == tokens``; rebuilt whenever the controller holds a different
        mask tensor (enable_cfattn(new_mask) for the next keyframe batch), validated by identity."""
        chosen = None
        if self.ctrl.attn_mask is not None:
            for m in self.ctrl.attn_mask:
                if m.shape[1] == tokens:
                    chosen = m
        hit = self._plans.get(tokens)
        if hit is None or hit[0] is not chosen:
            plan = None if chosen is None else ShardedKVExchange(chosen, self.world, self.rank, self.group)
            hit = (chosen, plan)
            self._plans[tokens] = hit
        return hit[1]

    def __call__(self, q, k, v, heads: int, ref_q=None, ref_k=None):
        """q, k, v: [chunks * N_local, L, C] projections of the local frames; ref_q / ref_k: projections of the
        stored reference features (spatial-guided attention) or None.  Returns [chunks * N_local, L, C]."""
        ctrl, be, chunks = self.ctrl, self.backend, self.chunks
        B, L, C = q.shape
        n_local = B // chunks
        d = C // heads
        inv = 1.0 / math.sqrt(d)
        q_main = q
        if ctrl.use_intraattn and ref_q is not None:
            q_main = be.attn_fwd(ref_q, ref_k, q, heads, 1, ctrl.intraattn_scale_factor * inv, float(ctrl.intraattn_bias))
        if ctrl.use_cfattn:
            plan = self._plan(L)
            if plan is None:
                # no mask at this resolution: everybody attends global frame 0 (owned by rank 0)
                k0 = k.view(chunks, n_local, L, C)[:, 0].contiguous()
                v0 = v.view(chunks, n_local, L, C)[:, 0].contiguous()
                if self.world > 1:
                    kv = torch.stack([k0, v0])
                    src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
                    dist.broadcast(kv, src=src, group=self.group)
                    k0, v0 = kv[0], kv[1]
                k_att, v_att = k0, v0
            else:
                k_att, v_att = plan.exchange(k, v, chunks, be)
            out = be.attn_fwd(q_main, k_att, v_att, heads, n_local, inv)
        else:
            out = be.attn_fwd(q_main, k, v, heads, 1, inv)
        if ctrl.use_interattn:
            traj = ctrl.trajectories(L)
            if traj is None:
                raise RuntimeError(f"temporal-guided attention enabled but no mapping matches {L} tokens")
            fwd_map, traj_mask = traj
            qg = _gather_frames(q, chunks, self.world, self.group)
            kg = _gather_frames(k, chunks, self.world, self.group)
            og = _gather_frames(out, chunks, self.world, self.group)
            full = be.temporal_attn_fwd(qg, kg, og, fwd_map, traj_mask, chunks, heads, ctrl.interattn_scale_factor * inv)
            n = n_local * self.world
            out = full.view(chunks, n, L, C)[:, self.rank * n_local:(self.rank + 1) * n_local].reshape(B, L, C).contiguous()
        return out
