"""Round-2 diagnostic d: the saved-activation buffer of the tensor-core forward vs the fp32 forward, section by section."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crank_b200 import lib as L
from crank_b200.parallel_wavegan.models import ParallelWaveGANGenerator

torch.manual_seed(3)
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 500
net = ParallelWaveGANGenerator(in_channels=80, out_channels=64, kernel_size=5, layers=8, stacks=4, aux_channels=0,
                               upsample_conditional_features=False).cuda()
cfg = net.cfg
x = torch.randn(B, T, 80, device="cuda")
F = B * T
n_act = L.lib().crk_wavenet_act_floats(C.byref(cfg), B, T)


def fwd(prec, enable=0):
    L.set_precision(prec)
    L.check(L.lib().crk_debug_opt_enable(enable), "en")
    net._weff, net._weff_key = None, None
    weff = net.effective_weights()
    y = torch.empty(B, T, 64, device="cuda")
    act = torch.zeros(n_act, device="cuda")
    L.call("crk_wavenet_fwd", C.byref(cfg), L.ptr(weff), L.ptr(x), 80, None, 0, None, L.ptr(y), 64, L.ptr(act), B, T)
    torch.cuda.synchronize()
    L.check(L.lib().crk_debug_opt_enable(0), "en")
    return y, act


y0, a0 = fwd("fp32")
for label, prec, en in (("tf32x3 k_resblock_fwd_tc", "tf32x3", 0), ("tf32x3 k_resblock_fwd_pt", "tf32x3", 1)):
    y1, a1 = fwd(prec, en)
    print(f"== {label}: y max abs diff {(y1 - y0).abs().max().item():.2e} (max |y| {y0.abs().max().item():.2e})")
    h0, h1 = a0[: 9 * F * 64].view(9, F, 64), a1[: 9 * F * 64].view(9, F, 64)
    print("  h per layer  max abs diff:", " ".join(f"{(h1[l] - h0[l]).abs().max().item():.1e}" for l in range(9)))
    t0_, t1_ = a0[9 * F * 64: 9 * F * 64 + 8 * F * 128].view(8, B, T, 128), a1[9 * F * 64: 9 * F * 64 + 8 * F * 128].view(8, B, T, 128)
    for l in range(8):
        d = (t1_[l] - t0_[l]).abs()
        bad = (d > 1e-4)
        rows = bad.any(-1)                       # (B, T)
        where = rows.nonzero()
        print(f"  tasb layer {l}: max abs diff {d.max().item():.2e}, elements > 1e-4: {int(bad.sum())}, frames affected {int(rows.sum())}"
              + (f", e.g. (b,t) {where[:6].tolist()}, t%128 of affected: {sorted(set((where[:, 1] % 128).tolist()))[:12]}" if len(where) else ""))
L.set_precision("tf32x3")
