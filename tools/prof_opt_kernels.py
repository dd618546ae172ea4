"""Launch each optimisation kernel once at the layer-3 shape ([16,640,64,64]) -- target for ncu captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fresco_b200 import diffusion_hacked as dh, ops  # noqa: E402
from fresco_b200.flow_utils import adjoint_csr  # noqa: E402
from fresco_b200.harness import synth  # noqa: E402

N, res, C, h = 8, 512, 640, 64
dev = torch.device("cuda")
flows, occs = synth.smooth_flows(N, res, res, dev, seed=2)
g = torch.Generator().manual_seed(h)
sample = torch.randn(2 * N, C, h, h, generator=g).half().to(dev)
ref = (sample.float() + 0.5 * torch.randn(2 * N, C, h, h, generator=g).to(dev)).half()
target = dh.gram_targets([ref])[0]
cs = sample.float().reshape(2, N, C, h, h).contiguous()
B, L = 2 * N, h * h
_, ff, bf, fo, bo = dh.resize_flows_occs(flows, occs, h)
adj = adjoint_csr(flows, occs, h)
kf, kb = (1 - fo).reshape(N, h, h).contiguous(), (1 - bo).reshape(N, h, h).contiguous()
gr = torch.empty_like(cs)
for _ in range(2):
    xhat, norms = ops.gram_normalize(cs.view(B, C, L))
    tsign = ops.gram_sign(xhat, target, 100.0)
    ops.gram_grad(tsign, xhat, norms, gr.view(B, C, L), 100.0)
    ops.warp_loss_fwd_bwd(cs, ff, bf, kf, kb, gr, adjoint=adj)
    m = torch.zeros_like(cs)
    v = torch.zeros_like(cs)
    ops.adam_step(cs, gr, m, v, 1)
torch.cuda.synchronize()
print("done")
