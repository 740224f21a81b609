"""Bit-identity of two LSGAN steps with the weight-gradient side stream on vs off (same kernels, only the stream assignment
and the per-layer operand buffers differ): any ordering bug between the side stream and the main stream shows up as a
difference.  Shapes of profiles/dp_equiv.py (16 utterances x 300 frames, ragged masks) and the bench shape (64 x 500)."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from crank_b200 import lib as L
from profiles.dp_equiv import build, state
from crank_b200.synthetic import make_batch, to_device

dev = torch.device("cuda", 0)
for B, T in ((16, 300), (64, 500)):
    res = []
    for mask in (0, 512, 0):
        L.check(L.lib().crk_debug_opt_disable(mask), "opt")
        P = build("lsgan", 14, dev)
        for i in range(2):
            random.seed(100 + i)
            P.train(to_device(make_batch(B, T, 14, seed=70 + i, ragged=True), dev), "train")
        torch.cuda.synchronize()
        res.append(state(P))
    L.lib().crk_debug_opt_disable(0)
    for name, a, b in (("side on vs off", res[0], res[1]), ("side on vs on (rerun)", res[0], res[2])):
        bad = [k for k in a if not torch.equal(a[k], b[k])]
        worst = max(((a[k] - b[k]).abs().max().item() / max(a[k].abs().max().item(), 1e-30), k) for k in a)
        print(f"{B}x{T} {name}: {len(bad)} of {len(a)} tensors differ; worst rel diff {worst[0]:.2e} ({worst[1]})")
