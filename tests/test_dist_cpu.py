"""World-size-2 gloo test of the frame-sharded attention exchange (host logic of fresco_b200/dist.py).
Compute is injected as a torch backend (the CUDA kernels need a GPU); the result of the two ranks,
concatenated, must equal the single-process oracle on the full batch."""
import math
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class TorchBackend:
    """torch stand-ins with the semantics of the C entry points (fresco_kv_compact_packed, fresco_rows_gather /
    _scatter, fresco_attn_fwd[_kv_strided], fresco_temporal_attn_fwd[_strided])"""

    def kv_compact_packed(self, k, v, idx, kv_out, chunks):
        B, L, C = k.shape
        rows = (B // chunks) * L
        n = idx.numel()
        kv_out[:, :n, :C] = k.reshape(chunks, rows, C)[:, idx.long()]
        kv_out[:, :n, C:] = v.reshape(chunks, rows, C)[:, idx.long()]
        kv_out[:, n:] = float("nan")                       # padding rows must never be read
        return kv_out

    def rows_gather(self, src, idx, dst, dst_col=0):
        dst[:, dst_col:dst_col + src.shape[-1]] = src[idx.long()]
        return dst

    def rows_scatter(self, src, idx, dst):
        dst[idx.long()] = src
        return dst

    def attn_fwd(self, q, k, v, heads, q_per_kv, softmax_scale, diag_bias=0.0):
        B, L, C = q.shape
        d = C // heads
        k = k.repeat_interleave(q_per_kv, 0)
        v = v.repeat_interleave(q_per_kv, 0)
        qh, kh, vh = (t.view(t.shape[0], -1, heads, d).transpose(1, 2) for t in (q, k, v))
        s = qh @ kh.transpose(-1, -2) * softmax_scale
        if diag_bias:
            s = s + torch.eye(L, kh.shape[2]) * diag_bias
        return (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, L, C)

    def attn_fwd_kv_packed(self, q, kv, heads, q_per_kv, softmax_scale):
        C = q.shape[-1]
        return self.attn_fwd(q, kv[..., :C].contiguous(), kv[..., C:].contiguous(), heads, q_per_kv, softmax_scale)

    def temporal_attn_fwd_packed(self, qkv, fwd_map, traj_mask, chunks, heads, scale):
        C = qkv.shape[-1] // 3
        return self.temporal_attn_fwd(qkv[..., :C].contiguous(), qkv[..., C:2 * C].contiguous(), qkv[..., 2 * C:].contiguous(),
                                      fwd_map, traj_mask, chunks, heads, scale)

    def temporal_attn_fwd(self, q, k, v, fwd_map, traj_mask, chunks, heads, scale):
        B, L, C = q.shape
        N = B // chunks
        d = C // heads
        gi = fwd_map[None, :, :, None, None].expand(chunks, N, L, heads, d)
        qt = torch.gather(q.view(chunks, N, L, heads, d), 2, gi)
        kt = torch.gather(k.view(chunks, N, L, heads, d), 2, gi)
        vt = torch.gather(v.view(chunks, N, L, heads, d), 2, gi)
        s = torch.einsum("bfphd,bgphd->bphfg", qt, kt) * scale
        s = s.masked_fill(~traj_mask.bool()[None, :, None], float("-inf"))
        o = torch.einsum("bphfg,bgphd->bfphd", torch.softmax(s, -1), vt)
        out = torch.empty_like(o)
        out.scatter_(2, gi, o)
        return out.reshape(B, L, C)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fresco_b200 import diffusion_hacked as dh
    from fresco_b200.dist import ShardedFRESCOAttention, frame_range
    from oracle import fresco_oracle as O
    torch.manual_seed(0)
    N, chunks, L, heads, C = 4, 2, 64, 2, 16
    x = torch.randn(chunks * N, L, C)
    ref = torch.randn(chunks * N, L, C)
    w = [torch.randn(C, C) * 0.3 for _ in range(4)]
    bo = torch.randn(C) * 0.1
    masks = [torch.rand(N, L) > 0.6]
    masks[0][0] = True
    perm = torch.stack([torch.randperm(L) for _ in range(N)])
    perm[0] = torch.arange(L)
    fm = perm[:, None, :]
    bm = torch.argsort(perm, dim=1)[:, None, :]
    im = (torch.rand(L, 1, N, N) > 0.3) | torch.eye(N, dtype=torch.bool)[None, None]
    for flags in range(8):
        cf, intra, inter = bool(flags & 1), bool(flags & 2), bool(flags & 4)
        want = O.fresco_attention(x, w[0], w[1], w[2], w[3], bo, heads, use_cfattn=cf, attn_masks=masks,
                                  use_intraattn=intra, ref_hidden=ref, use_interattn=inter, fwd_mappings=[fm],
                                  bwd_mappings=[bm], interattn_masks=[im])
        ctrl = dh.AttentionControl()
        if intra:
            ctrl.stored_attn["decoder_attn"] = [ref]
            ctrl.enable_intraattn()
        if inter:
            ctrl.enable_interattn({"fwd_mappings": [fm], "bwd_mappings": [bm], "interattn_masks": [im]})
        if cf:
            ctrl.enable_cfattn(masks)
        lo, hi = frame_range(N, world, rank)
        sel = torch.cat([torch.arange(c * N + lo, c * N + hi) for c in range(chunks)])
        xl, rl = x[sel], ref[sel]
        sharded = ShardedFRESCOAttention(ctrl, world, rank, chunks, backend=TorchBackend())
        a = sharded(xl @ w[0].t(), xl @ w[1].t(), xl @ w[2].t(), heads, ref_q=rl @ w[0].t(), ref_k=rl @ w[1].t())
        got = a @ w[3].t() + bo
        err = (got - want[sel]).abs().max().item()
        assert err < 1e-4, (flags, rank, err)
    if rank == 0:
        ret.put("ok")
    dist.destroy_process_group()


def test_sharded_attention_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == "ok"


def test_frame_range():
    sys.path.insert(0, ROOT)
    from fresco_b200.dist import frame_range
    assert [frame_range(16, 4, r) for r in range(4)] == [(0, 4), (4, 8), (8, 12), (12, 16)]


# ------------------------------------------------------------------------------------------------ sharded inference loop
class _StubVAE:
    class config:
        scaling_factor = 0.5

    def encode(self, x):
        class D:
            def __init__(self, v):
                self.latent_dist = self
                self.v = v

            def sample(self):
                return self.v
        return D(torch.nn.functional.avg_pool2d(x, 8)[:, :1].repeat(1, 4, 1, 1))


class _StubUNet(torch.nn.Module):
    """per-frame 'UNet': every output frame depends on its own input frame only, like the real UNet body"""

    class config:
        in_channels = 4

    def __init__(self):
        super().__init__()
        self.dtype = torch.float32

    def forward(self, x, t, encoder_hidden_states=None, **kw):
        return (torch.tanh(x * 0.7 + 0.001 * float(t)) + 0.1 * encoder_hidden_states.mean((1, 2))[:, None, None, None],)


class _StubPipe:
    def __init__(self, ac):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from test_cpu_host import _Sched
        self.scheduler = _Sched(ac)
        self.vae = _StubVAE()
        self.unet = _StubUNet()
        self._execution_device = torch.device("cpu")

    def prepare_latents(self, B, C, H, W, dtype, device, generator, latents=None):
        return torch.randn(B, C, H // 8, W // 8, generator=generator, dtype=dtype)

    def progress_bar(self, total=None):
        import contextlib
        return contextlib.nullcontext()


class _Ctrl:
    class controller:
        @staticmethod
        def disable_intraattn():
            pass

        @staticmethod
        def disable_interattn():
            pass


def _inference_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    from fresco_b200 import pipe_FRESCO as pf
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_cpu_host import _TorchStepOps
    pf.ops = _TorchStepOps
    ac = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "ddpm_step.npz"))["alphas_cumprod"])
    N = 4
    g = torch.Generator().manual_seed(1)
    imgs = torch.rand(N, 3, 64, 64, generator=g)
    pe = torch.randn(2 * N, 5, 8, generator=g)
    timesteps = [950 - 50 * i for i in range(20)]
    ok = True
    for repeat in (False, True):
        rec_full, rec_sh = [], []
        for propagation in (False, True):
            kw = dict(num_warmup_steps=14, use_controlnet=False, seed=3, repeat_noise=repeat, bg_smoothing_steps=[],
                      propagation_mode=propagation)
            full = pf.inference(_StubPipe(ac), None, _Ctrl, imgs, pe, None, timesteps, record_latents=rec_full, **kw)
            mine = pf.inference(_StubPipe(ac), None, _Ctrl, imgs, pe, None, timesteps, record_latents=rec_sh,
                                shard=(world, rank, None), **kw)
            lo, hi = rank * N // world, (rank + 1) * N // world
            ok = ok and torch.equal(mine, full[lo:hi]) and len(rec_full) == len(rec_sh)
            ok = ok and all(torch.equal(a, b) for a, b in zip(rec_full, rec_sh))
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


def test_sharded_inference_loop_world2_gloo():
    """fresco_b200.pipe_FRESCO.inference frame-sharded over 2 ranks == unsharded, bit for bit: shared noise streams,
    repeat_noise, record_latents (first batch and propagation mode)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_inference_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


# ------------------------------------------------------------------------------------------------ channel-resharded warp chain
def _warp_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fresco_b200 import flow_utils as fu
    from oracle import fresco_oracle as O
    from test_cpu_host import _TorchWarpOps
    fu.ops = _TorchWarpOps
    N, chunks, C, h = 4, 2, 6, 8
    flows, occs = O.synth_flows(N, 64, 64, seed=5, mag=5.0)
    sal = torch.rand(N, 1, 32, 32, generator=torch.Generator().manual_seed(2))
    feat = torch.randn(chunks * N, C, h, h, generator=torch.Generator().manual_seed(3))
    want = O.warp_tensor(feat, flows, occs, sal, chunks)
    lo, hi = rank * N // world, (rank + 1) * N // world
    sel = torch.cat([torch.arange(c * N + lo, c * N + hi) for c in range(chunks)])
    got = fu.warp_tensor(feat[sel].contiguous(), flows, occs, sal, chunks, shard=(world, rank, None))
    ok = (got - want[sel]).abs().max().item() < 1e-5
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


def test_sharded_warp_tensor_world2_gloo():
    """warp_tensor on a frame-sharded batch: re-shard by channel (all-to-all), chain, all-to-all back == the oracle on the
    full batch."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_warp_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


# ------------------------------------------------------------------------------------------------ ring-halo feature optimisation
class TorchOptBackend:
    """torch stand-ins for the compute steps of optimize_feature (fresco_warp_loss_fwd_bwd[_halo], the Gram-L1 kernels,
    fresco_adam_step, fresco_adain), written from the oracle's formulas"""

    def __init__(self, O):
        self.O = O

    def temporal_prepare(self, flows, occs, h, lo, hi, total):
        ff, bf, fo, bo = self.O._resize_flow_occ(flows, occs, h / flows[0].shape[2])
        return {"ff": ff[lo:hi], "bf": bf[lo:hi], "mf": (1 - fo)[lo:hi], "mb": (1 - bo)[lo:hi]}

    def temporal(self, cs, prep, grad, loss_acc, halo_cs=None, halo_grad=None, total_frames=None):
        O = self.O
        b, n, c, h, w = cs.shape
        rep = lambda t: t.repeat(b, 1, 1, 1)
        ff, bf, mf, mb = rep(prep["ff"]), rep(prep["bf"]), rep(prep["mf"]), rep(prep["mb"])
        if halo_cs is None:
            loss, g = O.temporal_loss_and_grad(cs, ff, bf, mf, mb)
        else:
            ext = torch.cat([cs, halo_cs[:, None]], 1)
            c1 = ext[:, :n].reshape(b * n, c, h, w)
            c2 = ext[:, 1:].reshape(b * n, c, h, w)
            r1 = c2 - O.flow_warp(c1, bf)
            r2 = c1 - O.flow_warp(c2, ff)
            k = 2.0 / (b * total_frames * c * h * w)
            loss = ((r1 * mb).abs() + (r2 * mf).abs()).sum() * k
            s1, s2 = torch.sign(r1) * mb * k, torch.sign(r2) * mf * k
            g = (s2 - O.flow_warp_adjoint(s1, bf)).view(b, n, c, h, w).clone()
            g2 = (s1 - O.flow_warp_adjoint(s2, ff)).view(b, n, c, h, w)
            g[:, 1:] += g2[:, :-1]
            halo_grad.copy_(g2[:, -1])
        grad.copy_(g)
        if loss_acc is not None:
            loss_acc += loss

    def spatial(self, cs_bcl, target, weight, grad_bcl, loss_acc):
        B, C, L = cs_bcl.shape
        loss, g = self.O.spatial_loss_and_grad(cs_bcl.reshape(1, B, C, L, 1), target, weight)
        grad_bcl += g.reshape(B, C, L)
        if loss_acc is not None:
            loss_acc += loss

    def adam(self, cs, grad, m, v, it):
        m.mul_(0.9).add_(grad, alpha=0.1)
        v.mul_(0.999).addcmul_(grad, grad, value=0.001)
        cs.sub_((0.2 / (1 - 0.9 ** it)) * m / (v.sqrt() / math.sqrt(1 - 0.999 ** it) + 1e-8))

    def adain(self, cs_bchw, sample):
        return self.O.adain(cs_bchw.to(sample.dtype), sample)


def _optimize_worker(rank, world, port, ret, subgroups=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fresco_b200 import diffusion_hacked as dh
    from oracle import fresco_oracle as O
    group, seed = None, 0
    if subgroups:
        # two independent sharded batches on one box: ranks {0,1} and {2,3}; the second group does not contain global
        # rank 0, so every peer index has to go through the group -> global mapping
        groups = [dist.new_group([0, 1]), dist.new_group([2, 3])]
        group, seed = groups[rank // 2], rank // 2
        rank, world = rank % 2, 2
    N, chunks, C, h = 4, 2, 6, 8
    flows, occs = O.synth_flows(N, 64, 64, seed=11 + seed, mag=5.0)
    g = torch.Generator().manual_seed(4 + seed)
    base = torch.randn(chunks, 1, C, h, h, generator=g)
    feat = (base + 0.5 * torch.randn(chunks, N, C, h, h, generator=g)).reshape(chunks * N, C, h, h)
    ref = feat + 0.3 * torch.randn(feat.shape, generator=g)
    X = ref.reshape(chunks * N, C, h * h).transpose(1, 2)
    Xh = X / (X ** 2).sum(2, keepdim=True) ** 0.5
    target = torch.bmm(Xh, Xh.transpose(1, 2))
    lo, hi = rank * N // world, (rank + 1) * N // world
    sel = torch.cat([torch.arange(c * N + lo, c * N + hi) for c in range(chunks)])
    ok = True
    be = TorchOptBackend(O)
    for temporal, spatial in ((True, True), (True, False), (False, True)):
        corr = [target] if spatial else []
        corr_l = [target[sel]] if spatial else []
        tr_full, tr_mine = dh.OptimizeTrace(), dh.OptimizeTrace()
        full = dh.optimize_feature(feat, flows, occs, corr, iters=6, optimize_temporal=temporal, trace=tr_full, backend=be)
        want = O.optimize_feature(feat, flows, occs, corr, iters=6, optimize_temporal=temporal)
        mine = dh.optimize_feature(feat[sel].contiguous(), flows, occs, corr_l, iters=6, optimize_temporal=temporal,
                                   trace=tr_mine, shard=(world, rank, group), backend=be)
        if not subgroups:       # (one fixed seed: two fp32 implementations of 6 Adam steps may part ways on other data)
            ok = ok and (full - want).abs().max().item() < 1e-4                # the host loop itself against the oracle
        ok = ok and (mine - full[sel]).abs().max().item() < 1e-4               # sharded == unsharded
        ok = ok and all(abs(a - b) < 1e-4 * abs(a) for a, b in zip(tr_full.losses, tr_mine.losses))
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(int(t.item()))
    dist.destroy_process_group()


def test_sharded_optimize_feature_in_subgroups_gloo():
    """Two sharded batches side by side (process groups {0,1} and {2,3} of a world of 4): RingComm must address its
    neighbours through the group (the second group does not contain global rank 0)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_optimize_worker, args=(r, 4, port, ret, True)) for r in range(4)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_optimize_feature_gloo(world):
    """optimize_feature on a frame-sharded batch (ring halo of one boundary frame per Adam iteration, SURVEY 8e exchange
    3): all ranks together == the unsharded call == the oracle, loss curves included.  World 4 leaves ONE frame per rank
    (every pair crosses a rank boundary; previous and next neighbour differ)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_optimize_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1
