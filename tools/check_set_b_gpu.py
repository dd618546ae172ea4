"""Round-2 to-do, runnable as is: the -m gpu counterpart of golden set B (N = 3 frames, 64 x 96 plane, head_dim 80;
tests/golden/set_b.npz, generated from the reference).  Prints error / bound per check; once it has been seen green on
a B200 the checks move into tests/test_gpu_kernels.py."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fresco_b200 import diffusion_hacked as dh, flow_utils as fu   # noqa: E402

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "set_b.npz"))


def T(a, device="cuda"):
    return torch.from_numpy(np.asarray(a)).to(device)


class FakeAttn(torch.nn.Module):
    def __init__(self):
        super().__init__()
        c = g["wq"].shape[0]
        self.heads = int(g["heads"])
        self.spatial_norm = self.group_norm = None
        self.norm_cross = self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.to_q, self.to_k, self.to_v = (torch.nn.Linear(c, c, bias=False) for _ in range(3))
        self.to_out = torch.nn.ModuleList([torch.nn.Linear(c, c), torch.nn.Dropout(0.0)])
        with torch.no_grad():
            for lin, k in ((self.to_q, "wq"), (self.to_k, "wk"), (self.to_v, "wv"), (self.to_out[0], "wo")):
                lin.weight.copy_(T(g[k], "cpu"))
            self.to_out[0].bias.copy_(T(g["bo"], "cpu"))


ok = True


def report(name, err, bound):
    global ok
    ok &= err < bound
    print(f"{name}: {err:.3e} (bound {bound:.3e}) {'OK' if err < bound else 'FAIL'}")


# mapping / trajectory mask: bit-exact
fm, bm, mask = fu.get_mapping_ind(T(g["bwd"]), T(g["bwd_occ"]), T(g["imgs"]), scale=8.0)
report("mapping mismatches", float((fm.cpu() != T(g["fwd_map"], "cpu")).sum() + (bm.cpu() != T(g["bwd_map"], "cpu")).sum()
                                   + (mask.cpu() != T(g["inter_mask"], "cpu")).sum()), 0.5)

# processor, four mode combinations
attn = FakeAttn().cuda().half()
x, ref_hidden = T(g["x"]).half(), T(g["ref_hidden"]).half()
masks = [T(g[f"attn_mask{i}"]) for i in range(3)]
paras = {"fwd_mappings": [T(g["fwd_map"])], "bwd_mappings": [T(g["bwd_map"])], "interattn_masks": [T(g["inter_mask"])]}
for flags in (0, 1, 6, 7):
    ctrl = dh.AttentionControl()
    proc = dh.FRESCOAttnProcessor2_0(2, ctrl)
    if flags & 2:
        ctrl.stored_attn["decoder_attn"] = [ref_hidden.clone()]
        ctrl.enable_intraattn()
    if flags & 4:
        ctrl.enable_interattn(paras)
    if flags & 1:
        ctrl.enable_cfattn(masks)
    with torch.no_grad():
        out = proc(attn, x.clone()).float().cpu()
    ref = T(g[f"out_{flags}"], "cpu")
    report(f"processor flags={flags}", (out - ref).abs().max().item(), 1e-2 * ref.abs().max().item())

# warp_tensor (decoder feature, image with the dilation path)
flows, occs, sal = [T(g["fwd"]), T(g["bwd"])], [T(g["fwd_occ"]), T(g["bwd_occ"])], T(g["saliency"])
report("warp_tensor feat", (fu.warp_tensor(T(g["sample_feat"]), flows, occs, sal, 2).cpu() - T(g["out_feat"], "cpu")).abs().max().item(), 2e-5)
report("warp_tensor img", (fu.warp_tensor(T(g["sample_img"]), flows, occs, sal, 1).cpu() - T(g["out_img"], "cpu")).abs().max().item(), 2e-5)

# optimize_feature: loss curves, 1-iteration output by fraction of elements
for tag, iters in (("full1", 1), ("full3", 3)):
    tr = dh.OptimizeTrace()
    out = dh.optimize_feature(T(g["opt_sample"]), flows, occs, correlation_matrix=[T(g["opt_target"])], intra_weight=1e2,
                              iters=iters, trace=tr)
    ref_l = g[f"opt_{tag}_losses"]
    report(f"optimize {tag} loss curve rel", float(np.abs(np.array(tr.losses) / ref_l - 1).max()), 1e-2)
    diff = (out.cpu() - T(g[f"opt_{tag}_out"], "cpu")).abs()
    if iters == 1:
        report("optimize full1 fraction of elements off by a sign flip", (diff > 2e-3).float().mean().item(), 0.03)
    else:
        report("optimize full3 rel-mean", (diff.mean() / T(g[f"opt_{tag}_out"], "cpu").abs().mean()).item(), 0.1)
print("ALL OK" if ok else "FAILURES")
