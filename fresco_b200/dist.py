"""Frame-sharded FRESCO attention for one multi-GPU box (NCCL over NVLink / NVSwitch).

The reference has no collective on this path; this module is the B200 design for BASELINE config 4: the
N keyframes of ONE batch are split contiguously over G ranks (rank r owns frames [r*N/G, (r+1)*N/G) of both CFG
chunks) and the UNet body runs as plain replicas.  Only FRESCO attention couples frames:

* cross-frame attention (src/diffusion_hacked.py:225-247, :303-305): every query frame attends the same K/V set = all
  tokens of frame 0 + the occluded tokens of frames 1..N-1.  Each rank compacts the selected rows of ITS frames straight
  into the send buffer (K | V side by side in one 2C-wide row), ONE all-gather per layer (padded to the largest
  per-rank count, known on the host from the per-batch mask) delivers every rank's rows, one row-gather drops the
  padding, and the attention kernel reads K and V out of the interleaved buffer through strided TMA maps -- the
  reference's row-major (frame, token) order is kept because rank order == frame order.
* temporal-guided attention (:320-367) mixes all N frames along each flow trajectory.  It is sharded by TRAJECTORY:
  rank s owns trajectories [s*L/G, (s+1)*L/G); one all-to-all sends it the q | k | attention-output rows those
  trajectories visit in every rank's frames (1/G of the data instead of an all-gather of all of it), it runs the
  N x N attention for its trajectories only (1/G of the work instead of all of it on every rank), and a second
  all-to-all returns the result rows to the ranks that own the frames.
* spatial-guided attention (:257-288) is per frame: local.

Compute goes through a small backend object (default: the CUDA kernels in ``ops``); the CPU tests inject a torch
backend to exercise the exchange / ordering logic under gloo.  Nothing here synchronises with the host after the
per-batch plans exist, so a whole UNet step (collectives included) can be captured in a CUDA graph.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def frame_range(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    if n_frames % world != 0:
        raise ValueError(f"{n_frames} frames do not split evenly over {world} ranks")
    per = n_frames // world
    return rank * per, (rank + 1) * per


class RingComm:
    """Neighbour exchange of the frame ring (exchange 3): ``shift(send, recv, +1)`` sends ``send`` to rank+1
    and fills ``recv`` from rank-1 (``-1``: the other way round).  One batched NCCL send/recv pair on the current
    stream: no host synchronisation, capturable in a CUDA graph.  Tests inject an object with the same method."""

    def __init__(self, world: int, rank: int, group=None):
        self.world, self.rank, self.group = world, rank, group

    def _peer(self, r: int) -> int:
        r %= self.world
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def shift(self, send: torch.Tensor, recv: torch.Tensor, direction: int) -> None:
        dst, src = self._peer(self.rank + direction), self._peer(self.rank - direction)
        reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, send, dst, self.group),
                                       dist.P2POp(dist.irecv, recv, src, self.group)])
        for r in reqs:
            r.wait()

    def all_reduce_sum(self, t: torch.Tensor) -> None:
        dist.all_reduce(t, group=self.group)


class _OpsBackend:
    """the product path: hand-written sm_100a kernels"""

    def kv_compact_packed(self, k, v, idx, kv_out, chunks):
        from . import ops
        return ops.kv_compact_packed(k, v, idx, kv_out, chunks)

    def rows_gather(self, src, idx, dst, dst_col=0):
        from . import ops
        return ops.rows_gather(src, idx, dst, dst_col)

    def rows_scatter(self, src, idx, dst):
        from . import ops
        return ops.rows_scatter(src, idx, dst)

    def attn_fwd(self, q, k, v, heads, q_per_kv, softmax_scale, diag_bias=0.0):
        from . import ops
        return ops.attn_fwd(q, k, v, heads, q_per_kv, softmax_scale=softmax_scale, diag_bias=diag_bias)

    def attn_fwd_kv_packed(self, q, kv, heads, q_per_kv, softmax_scale):
        from . import ops
        return ops.attn_fwd_kv_packed(q, kv, heads, q_per_kv, softmax_scale)

    def temporal_attn_fwd(self, q, k, v, fwd_map, traj_mask, chunks, heads, scale):
        from . import ops
        return ops.temporal_attn_fwd(q, k, v, fwd_map, traj_mask, chunks, heads, scale)

    def temporal_attn_fwd_packed(self, qkv, fwd_map, traj_mask, chunks, heads, scale):
        from . import ops
        return ops.temporal_attn_fwd_packed(qkv, fwd_map, traj_mask, chunks, heads, scale)


class ShardedKVExchange:
    """Per-batch plan + per-layer all-gather of the compacted cross-frame K/V (exchange 1)."""

    def __init__(self, attn_mask: torch.Tensor, world: int, rank: int, chunks: int, group=None):
        n_frames, tokens = attn_mask.shape
        self.world, self.rank, self.group, self.chunks = world, rank, group, chunks
        self.lo, self.hi = frame_range(n_frames, world, rank)
        per_frame = attn_mask.sum(1).tolist()                               # one host sync per batch and resolution
        n_local = self.hi - self.lo
        counts = [int(sum(per_frame[r * n_local:(r + 1) * n_local])) for r in range(world)]
        self.counts = counts
        self.max_count = max(max(counts), 1)
        self.total = sum(counts)
        dev = attn_mask.device
        local = attn_mask[self.lo:self.hi].reshape(-1)
        self.local_idx = torch.nonzero(local, as_tuple=False).reshape(-1).to(torch.int32).contiguous()
        # row (b, i) of the dense result  <-  row ((r*chunks + b)*max_count + j) of the gathered buffer
        rows = []
        for b in range(chunks):
            for r in range(world):
                rows.append((r * chunks + b) * self.max_count + torch.arange(counts[r], dtype=torch.int64))
        self.unpad_idx = torch.cat(rows).to(torch.int32).to(dev).contiguous()
        self._bufs = {}

    def _buffers(self, C: int, like: torch.Tensor):
        key = (C, like.dtype)
        if key not in self._bufs:
            send = torch.empty(self.chunks, self.max_count, 2 * C, dtype=like.dtype, device=like.device)
            gathered = torch.empty(self.world, self.chunks, self.max_count, 2 * C, dtype=like.dtype, device=like.device)
            self._bufs[key] = (send, gathered)
        return self._bufs[key]

    def exchange(self, k_local: torch.Tensor, v_local: torch.Tensor, backend) -> torch.Tensor:
        """k_local, v_local [chunks * N_local, L, C] -> kv [chunks, Lk_global, 2C] (K | V per row)"""
        C = k_local.shape[-1]
        send, gathered = self._buffers(C, k_local)
        if self.counts[self.rank] > 0:
            backend.kv_compact_packed(k_local, v_local, self.local_idx, send, self.chunks)
        if self.world == 1:
            return send                                                     # max_count == total: nothing to drop
        dist.all_gather_into_tensor(gathered.view(-1, self.max_count, 2 * C), send, group=self.group)   # rank-major concat
        kv = torch.empty(self.chunks, self.total, 2 * C, dtype=k_local.dtype, device=k_local.device)
        backend.rows_gather(gathered.view(-1, 2 * C), self.unpad_idx, kv.view(-1, 2 * C))
        return kv


class ShardedTrajectoryExchange:
    """Per-batch plan of the trajectory-sharded temporal-guided attention (exchange 2)."""

    def __init__(self, fwd_map: torch.Tensor, traj_mask: torch.Tensor, world: int, rank: int, chunks: int, group=None):
        """fwd_map int64 [N, L] (trajectory p visits token fwd_map[f, p] of frame f); traj_mask uint8 [L, N, N]"""
        N, L = fwd_map.shape
        if L % world != 0:
            raise ValueError(f"{L} trajectories do not split evenly over {world} ranks")
        self.world, self.rank, self.group, self.chunks = world, rank, group, chunks
        self.N, self.L, self.Lg = N, L, L // world
        lo, hi = frame_range(N, world, rank)
        n_local = hi - lo
        self.n_local = n_local
        dev = fwd_map.device
        # send row (s, j, b, t)  <-  local row (b * n_local + j) * L + fwd_map[lo + j, s * Lg + t]
        fm = fwd_map[lo:hi].view(n_local, world, self.Lg).permute(1, 0, 2)                    # [s, j, t]
        brow = (torch.arange(chunks, device=dev)[None, :] * n_local + torch.arange(n_local, device=dev)[:, None]) * L   # [j, b]
        idx = fm[:, :, None, :] + brow[None, :, :, None]                                     # [s, j, b, t]
        self.row_idx = idx.reshape(-1).to(torch.int32).contiguous()
        self.identity_map = torch.arange(chunks * self.Lg, device=dev, dtype=torch.int64)[None].expand(N, -1).contiguous()
        mine = traj_mask[rank * self.Lg:(rank + 1) * self.Lg]                                 # [Lg, N, N]
        self.mask = mine[None].expand(chunks, -1, -1, -1).reshape(chunks * self.Lg, N, N).contiguous()
        self._bufs = {}

    def _buffers(self, C: int, like: torch.Tensor):
        key = (C, like.dtype)
        if key not in self._bufs:
            n = self.world * self.n_local * self.chunks * self.Lg
            mk = lambda w: torch.empty(n, w, dtype=like.dtype, device=like.device)
            self._bufs[key] = (mk(3 * C), mk(3 * C), mk(C))
        return self._bufs[key]

    def run(self, q, k, attn_out, heads: int, scale: float, backend) -> torch.Tensor:
        """q, k, attn_out [chunks * N_local, L, C] (local frames) -> temporal-guided output, same shape"""
        B, L, C = q.shape
        send, recv, back = self._buffers(C, q)
        for col, t in enumerate((q, k, attn_out)):
            backend.rows_gather(t.view(-1, C), self.row_idx, send, col * C)
        dist.all_to_all_single(recv, send, group=self.group)
        # recv rows are ordered (source rank r, its frame j, chunk b, trajectory t) = (global frame, chunk * Lg + t)
        res = backend.temporal_attn_fwd_packed(recv.view(self.N, self.chunks * self.Lg, 3 * C), self.identity_map, self.mask,
                                               1, heads, scale)
        dist.all_to_all_single(back, res.view(-1, C), group=self.group)
        out = torch.empty_like(attn_out)
        backend.rows_scatter(back, self.row_idx, out.view(-1, C))
        return out


class ShardedFRESCOAttention:
    """FRESCO self-attention of one layer on a frame shard (the sharded counterpart of
    ``FRESCOAttnProcessor2_0._fresco_self_attention``)."""

    def __init__(self, controller, world: int, rank: int, chunks: int = 2, group=None, backend=None):
        self.ctrl, self.world, self.rank, self.chunks, self.group = controller, world, rank, chunks, group
        self.backend = backend if backend is not None else _OpsBackend()
        self._plans = {}
        self._traj_plans = {}

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
            plan = None if chosen is None else ShardedKVExchange(chosen, self.world, self.rank, self.chunks, self.group)
            hit = (chosen, plan)
            self._plans[tokens] = hit
        return hit[1]

    def _traj_plan(self, tokens: int) -> ShardedTrajectoryExchange:
        traj = self.ctrl.trajectories(tokens)
        if traj is None:
            raise RuntimeError(f"temporal-guided attention enabled but no mapping matches {tokens} tokens")
        hit = self._traj_plans.get(tokens)
        if hit is None or hit[0] is not traj[0]:
            hit = (traj[0], ShardedTrajectoryExchange(traj[0], traj[1], self.world, self.rank, self.chunks, self.group))
            self._traj_plans[tokens] = hit
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
                out = be.attn_fwd(q_main, k0, v0, heads, n_local, inv)
            else:
                kv = plan.exchange(k, v, be)
                out = be.attn_fwd_kv_packed(q_main, kv, heads, n_local, inv)
        else:
            out = be.attn_fwd(q_main, k, v, heads, 1, inv)
        if ctrl.use_interattn:
            scale = ctrl.interattn_scale_factor * inv
            if self.world == 1:
                traj = ctrl.trajectories(L)
                if traj is None:
                    raise RuntimeError(f"temporal-guided attention enabled but no mapping matches {L} tokens")
                out = be.temporal_attn_fwd(q, k, out, traj[0], traj[1], chunks, heads, scale)
            else:
                out = self._traj_plan(L).run(q, k, out, heads, scale, be)
        return out
