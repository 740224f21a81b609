"""VQ argmin timing: fp32 CUDA-core kernel vs tensor-core kernel (+exact re-score), 64 x 500 frames, K=512."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crank_b200 import lib as L, ops

g = torch.Generator().manual_seed(0)
x = torch.randn(64, 500, 64, generator=g).cuda()
F = 64 * 500
for kind in ("spread", "trained-like", "fresh-init"):
    W = torch.randn(512, 64, generator=g)
    if kind == "fresh-init":
        W = (torch.rand(512, 64, generator=g) * 2 - 1) / 512
    if kind == "trained-like":      # ~80 live codes near the data, the rest dead (|w| ~ 1e5), as after EMA updates
        W[80:] *= 1e5
    W = W.cuda()
    for mode in ("fp32", "tf32x3"):
        L.set_precision(mode)
        for _ in range(3):
            ops.VQFn.apply(x, W)
        L.check(L.lib().crk_timing_enable(5))
        for _ in range(10):
            ops.VQFn.apply(x, W)
        import ctypes
        cnt, tot = ctypes.c_int(), ctypes.c_float()
        L.check(L.lib().crk_timing_read(ctypes.byref(cnt), ctypes.byref(tot)))
        L.lib().crk_timing_enable(0)
        us = 1e3 * tot.value / cnt.value
        print(f"{kind:13s} {mode:7s}: {us:7.1f} us/launch  algorithmic {520.0 * F / (us * 1e-6) / 1e9:7.1f} GB/s  ({2.0*F*64*512/(us*1e-6)/1e12:.1f} TFLOP/s)")
