"""Per-phase clock64() breakdown of the tensor-core residual-block forward kernel (perf debugging)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crank_b200 import lib as L
from crank_b200.parallel_wavegan.models import ParallelWaveGANGenerator

for prec in ("tf32x3", "tf32"):
    L.set_precision(prec)
    net = ParallelWaveGANGenerator(in_channels=80, out_channels=64, kernel_size=5, layers=8, stacks=4, aux_channels=0,
                                   upsample_conditional_features=False).cuda()
    B, T = 64, 500
    x = torch.randn(B, T, 80, device="cuda")
    ntiles = B * ((T + 127) // 128)
    buf = torch.zeros(ntiles * 8, dtype=torch.int64, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            net.forward_cl(x)
        L.check(L.lib().crk_debug_timestamps(buf.data_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); net.forward_cl(x); e1.record(); torch.cuda.synchronize()
        L.check(L.lib().crk_debug_timestamps(None))
    t = buf.view(ntiles, 8)[:, :6].double().cpu()
    d = (t[:, 1:] - t[:, :-1]).mean(0)
    names = ["stageX", "gemm1(taps,TMA)", "epi1(gate)", "gemm2", "epi2"]
    print(prec, "stack fwd ms", e0.elapsed_time(e1), "| last block per-CTA cycles:", {n: int(v) for n, v in zip(names, d)},
          "total", int((t[:, 5] - t[:, 0]).mean()), "span(all CTAs)", int(t[:, 5].max() - t[:, 0].min()))
