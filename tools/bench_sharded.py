"""torchrun --nproc-per-node G tools/bench_sharded.py : BASELINE config 4 building block.
N=16 keyframes, level-B FRESCO attention layer (L=4096, C=320, 8 heads), frames sharded over G GPUs with
one K/V all-gather per layer.  Checks the sharded result against the unsharded kernels on rank 0 and
prints per-layer time (max over ranks, CUDA events)."""
import json
import math
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fresco_b200 import diffusion_hacked as dh, ops  # noqa: E402
from fresco_b200.dist import ShardedFRESCOAttention, frame_range  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    N, chunks, L, heads, C = 16, 2, 4096, 8, 320
    g = torch.Generator().manual_seed(0)            # identical on every rank
    q = torch.randn(chunks * N, L, C, generator=g).half().cuda()
    k = torch.randn(chunks * N, L, C, generator=g).half().cuda()
    v = torch.randn(chunks * N, L, C, generator=g).half().cuda()
    mask = torch.rand(N, L, generator=g) > 0.71
    mask[0] = True
    mask = mask.cuda()
    ctrl = dh.AttentionControl()
    ctrl.enable_cfattn([mask])
    lo, hi = frame_range(N, world, rank)
    sel = torch.cat([torch.arange(c * N + lo, c * N + hi) for c in range(chunks)]).cuda()
    ql, kl, vl = q[sel].contiguous(), k[sel].contiguous(), v[sel].contiguous()
    sharded = ShardedFRESCOAttention(ctrl, world, rank, chunks)
    out = sharded(ql, kl, vl, heads)
    # reference: the unsharded kernels on the full batch
    idx = ctrl.kv_indices(L)
    ks, vs = ops.kv_compact(k, v, idx, chunks)
    full = ops.attn_fwd(q, ks, vs, heads, N, softmax_scale=1 / math.sqrt(C // heads))
    err = (out.float() - full[sel].float()).abs().max().item()
    for _ in range(3):
        sharded(ql, kl, vl, heads)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        sharded(ql, kl, vl, heads)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        lk = int(mask.sum().item())
        print(json.dumps({"gpus": world, "frames": N, "kv_len": lk, "layer_ms": round(ms.item(), 4),
                          "max_abs_err_vs_unsharded": err,
                          "allgather_bytes_per_rank": chunks * sharded._plan(L).max_count * 2 * C * 2,
                          "tflops_aggregate": round(4.0 * chunks * N * L * lk * C / ms.item() / 1e9, 1)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
