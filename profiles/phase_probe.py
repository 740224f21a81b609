"""Per-phase clock64() breakdown of the tensor-core kernels (perf debugging): encoder-0 stack, 64 x 500 frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from crank_b200 import lib as L
from crank_b200.parallel_wavegan.models import ParallelWaveGANGenerator

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
L.set_precision(prec)
net = ParallelWaveGANGenerator(in_channels=80, out_channels=64, kernel_size=5, layers=8, stacks=4, aux_channels=0,
                               upsample_conditional_features=False).cuda()
B, T = 64, 500
x = torch.randn(B, T, 80, device="cuda", requires_grad=True)
buf = torch.zeros(1024 * 16, dtype=torch.int64, device="cuda")


def run():
    y = net.forward_cl(x)
    y.square().mean().backward()
    torch.cuda.synchronize()


for _ in range(2):
    run()
# launch indices inside one fwd+bwd:  fwd TC: blocks 0..7 (idx 3 = k5 dil 2); conv TC: fwd first(0) head(1,2), bwd: head dgrads(3,4),
# then per layer dgrad...; wgrad TC: head(0,1) then per layer: Wos(2), conv(3), ...
PT = bool(int(os.environ.get("CRANK_B200_OPT_ENABLE", "0")) & 1)     # persistent pipelined forward: worker-side stamps
sel = {"resblock_fwd k5 d2": (1, 3, ["stage X0,X1", "wait G1(0)", "E1(0)", "wait G2(0)", "E2(0)", "remaining tiles"], 7) if PT else
                             (1, 3, ["stageX", "gemm1+TMA", "epi1", "gemm2+TaSb", "epi2"], 6),
       "conv dgrad k5 (K128,N64)": (3, 5, ["stage seg0,1", "-> acc(0) ready", "epilogue(0)+stage seg3", "tile 1", "(s5-s4)", "(s6-s5)"], 7) if (int(os.environ.get("CRANK_B200_OPT_ENABLE", "0")) & 2) else
                                   (3, 5, ["stageA", "mma+TMA", "tmem->smem", "coalesced epilogue"], 5),
       "gate backward (K128,N64,k1)": (4, 3, ["stage dH,dS->GOS", "mma+TMA", "tmem->smem", "gate' epilogue"], 5),
       "wgrad conv k5": (2, 3, ["issue loads+wait", "store G,X0", "taps(tile0)", "other tiles", "wait last", "epilogue"], 7),
       "wgrad out|skip k1 (N=128 G cols, Cin 64)": (2, 2, ["issue loads+wait", "store G,X0", "taps(tile0)", "other tiles", "wait last", "epilogue"], 7)}
for name, (kid, idx, labels, n) in sel.items():
    buf.zero_()
    L.check(L.lib().crk_debug_timestamps(buf.data_ptr(), kid, idx))
    run()
    L.check(L.lib().crk_debug_timestamps(None, 0, 0))
    full = buf.view(-1, 16).double().cpu()
    full = full[full[:, 0] > 0]
    t = full[:, :n]
    d = (t[:, 1:] - t[:, :-1]).mean(0)
    diag = {"resblock_fwd k5 d2": ["issuer wait for weights"],
            "conv dgrad k5 (K128,N64)": ["issuer wait for weights", "producer wait for free slot"],
            "wgrad out|skip k1 (N=128 G cols, Cin 64)": [],
            "gate backward (K128,N64,k1)": ["issuer wait for weights", "producer wait for free slot"],
            "wgrad conv k5": ["slot waits", "transposition", "barrier", "tile: wait prev MMAs", "tile: G store", "tile: raw store",
                              "tile: issue next loads", "tile: barrier"]}[name]
    print("   diag (cycles, CTA mean):", {l: int(full[:, 8 + i].mean()) for i, l in enumerate(diag)})
    if n == 7 and "conv" in name:
        print("   raw stamps (CTA mean, relative to s0):", [int(v) for v in (t - t[:, :1]).mean(0)])
    print(prec, name, "CTAs", len(t), {l: int(v) for l, v in zip(labels, d)}, "total", int((t[:, n - 1] - t[:, 0]).mean()),
          "kernel span", int(t[:, n - 1].max() - t[:, 0].min()))
