"""optimize_feature (BASELINE config 3 building block) at the four decoder levels of SD1.5 @512x512, N=8:
20 Adam iterations of temporal + normalised-Gram L1 loss.  Prints per-call time, per-kernel roofline numbers."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fresco_b200 import diffusion_hacked as dh, ops  # noqa: E402
from fresco_b200.harness import synth  # noqa: E402


def timeit(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    N, res = 8, 512
    dev = torch.device("cuda")
    flows, occs = synth.smooth_flows(N, res, res, dev, seed=2)
    total = 0.0
    for C, h in [(1280, 8), (1280, 16), (1280, 32), (640, 64)]:
        L = h * h
        g = torch.Generator().manual_seed(h)
        sample = torch.randn(2 * N, C, h, h, generator=g).half().to(dev)
        ref = (sample.float() + 0.5 * torch.randn(2 * N, C, h, h, generator=g).to(dev)).half()
        target = dh.gram_targets([ref])[0]
        ms = timeit(lambda: dh.optimize_feature(sample, flows, occs, correlation_matrix=[target], iters=20), 2)
        total += ms
        # kernel-level numbers for one iteration
        cs = sample.float().reshape(2, N, C, h, h).contiguous()
        B = 2 * N
        xhat, norms = ops.gram_normalize(cs.view(B, C, L))
        dense = target.dense()
        t_sign_dense = timeit(lambda: ops.gram_sign(xhat, dense, 100.0), 5)
        del dense
        t_sign = timeit(lambda: ops.gram_sign_ref(xhat, target.yhat, 100.0), 5)
        tsign = ops.gram_sign_ref(xhat, target.yhat, 100.0)
        grad = torch.zeros(B, C, L, device=dev)
        t_grad = timeit(lambda: ops.gram_grad(tsign, xhat, norms, grad, 100.0), 5)
        t_norm = timeit(lambda: ops.gram_normalize(cs.view(B, C, L)), 5)
        _, ff, bf, fo, bo = dh.resize_flows_occs(flows, occs, h)
        gr = torch.empty_like(cs)
        from fresco_b200.flow_utils import adjoint_csr
        adj = adjoint_csr(flows, occs, h)
        kf, kb = (1 - fo).reshape(N, h, h).contiguous(), (1 - bo).reshape(N, h, h).contiguous()
        t_warp = timeit(lambda: ops.warp_loss_fwd_bwd(cs, ff, bf, kf, kb, gr, adjoint=adj), 5)
        m = torch.zeros_like(cs)
        v = torch.zeros_like(cs)
        t_adam = timeit(lambda: ops.adam_step(cs, gr, m, v, 1), 5)
        flops_sign = 2.0 * B * L * L * C
        flops_grad = 2.0 * B * L * L * C
        print(json.dumps({
            "layer": "[%d,%d,%d,%d]" % (B, C, h, h), "optimize_feature_20it_ms": round(ms, 3),
            "gram_sign_ms": round(t_sign, 4), "gram_sign_tflops": round(flops_sign / t_sign / 1e9, 1),
            "gram_sign_dense_target_ms (round-1 kernel)": round(t_sign_dense, 4),
            "gram_grad_ms": round(t_grad, 4), "gram_grad_tflops": round(flops_grad / t_grad / 1e9, 1),
            "gram_normalize_ms": round(t_norm, 4), "warp_loss_ms": round(t_warp, 4),
            "warp_loss_GBps(4 passes)": round(4.0 * cs.numel() * 4 / t_warp / 1e6, 1),
            "adam_ms": round(t_adam, 4), "adam_GBps(7 passes)": round(7.0 * cs.numel() * 4 / t_adam / 1e6, 1)}))
    print(json.dumps({"optimize_feature_all_4_layers_ms": round(total, 2)}))


if __name__ == "__main__":
    main()
