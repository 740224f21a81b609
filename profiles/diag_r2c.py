"""Round-2 diagnostic c: which tensor-core kernel family carries the backward error?  One encoder-0 stack at
16 x 500; gradients in tf32x3 with individual families forced back to fp32 (crk_debug_tc_disable), all against the
all-fp32 CUDA-core result (verified against the oracle at <= 5e-5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crank_b200 import lib as L
from crank_b200.parallel_wavegan.models import ParallelWaveGANGenerator
from tests.util import rel_err

torch.manual_seed(3)
B, T = 16, 500
net = ParallelWaveGANGenerator(in_channels=80, out_channels=64, kernel_size=5, layers=8, stacks=4, aux_channels=0,
                               upsample_conditional_features=False).cuda()
x0 = torch.randn(B, T, 80, device="cuda")
dy_kinds = {"randn": torch.randn(B, T, 64, device="cuda"),
            "small smooth": 1e-4 * torch.cumsum(torch.randn(B, T, 64, device="cuda"), 1) / 20}


def run(prec, mask, dy):
    L.set_precision(prec)
    L.check(L.lib().crk_debug_tc_disable(mask), "mask")
    net.zero_grad(set_to_none=True)
    x = x0.clone().requires_grad_(True)
    y = net.forward_cl(x)
    y.backward(dy)
    torch.cuda.synchronize()
    g = {k: v.detach().clone() for k, v in net.named_conv_grads().items()}
    L.check(L.lib().crk_debug_tc_disable(0), "mask")
    return y.detach().clone(), x.grad.detach().clone(), g


for name, dy in dy_kinds.items():
    y0, dx0, g0 = run("fp32", 0, dy)
    print(f"== dy = {name}")
    for label, mask in (("all TC", 0), ("fwd fp32", 1), ("conv/dgrad fp32", 2), ("wgrad fp32", 4), ("gate-bwd fp32", 8),
                        ("fwd+gate fp32", 9), ("only fwd TC", 14), ("only dgrad TC", 13), ("only wgrad TC", 11), ("only gate TC", 7)):
        y, dx, g = run("tf32x3", mask, dy)
        rows = sorted(((rel_err(g[k], g0[k]), k) for k in g0 if g0[k].abs().max() > 0), reverse=True)
        wv = [r for r in rows if r[1].endswith("weight_v")]
        print(f"  {label:18s} y {rel_err(y, y0):.1e}  dx {rel_err(dx, dx0):.1e}  worst grad {rows[0][0]:.1e} ({rows[0][1]})  worst weight_v {wv[0][0]:.1e} ({wv[0][1]})")
    # depth profile of the conv weight gradients, all TC
    y, dx, g = run("tf32x3", 0, dy)
    print("  per layer conv.weight_v:", " ".join(f"{rel_err(g[f'conv_layers.{l}.conv.weight_v'], g0[f'conv_layers.{l}.conv.weight_v']):.1e}" for l in range(8)))
L.set_precision("tf32x3")
