"""Round-2 diagnostic: where does the 16x500 vqvae step diverge from the oracle after the first G update?
(a) per-tensor G gradients pre-Adam, (b) the speaker-adversarial pass with the ORACLE's updated parameters."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_gpu_baseline_shapes import _pair
from crank_b200 import lib as L
from crank_b200.synthetic import clone_batch, make_batch, to_device
from tests.util import rel_err

L.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32x3")
kind, S, B, T = "vqvae", 12, 16, 500
conf, om, pm, O, P = _pair(kind, S)
batch = make_batch(B, T, S, seed=0, ragged=True)
for k in om:
    pm[k].load_state_dict(om[k].state_dict())
# (a) gradients of the generator loss, no optimizer step
import types
def no_step(self, *a, **k):
    return None
for opt in list(O.optimizer.values()) + list(P.optimizer.values()):
    opt.step = types.MethodType(no_step, opt)
random.seed(100); ov = O.train(clone_batch(batch), "train")
random.seed(100); pv = P.train(to_device(clone_batch(batch), "cuda"), "train")
print("losses without any optimizer step:")
for k in sorted(ov):
    if ov[k]:
        print(f"  {k:16s} oracle {ov[k]:.6f} product {pv[k]:.6f} rel {abs(pv[k]-ov[k])/abs(ov[k]):.2e}")
# gradients left in .grad by the last backward of each sub-model
G_o, G_p = om["G"], pm["G"]
rows = []
for lst in ("encoders", "decoders"):
    for n in range(2):
        po_, oo_ = getattr(G_p, lst)[n], getattr(G_o, lst)[n]
        pg = po_.named_conv_grads()
        for name, prm in oo_.named_parameters():
            if prm.grad is None or name not in pg:
                continue
            den = prm.grad.abs().max().item()
            rows.append((rel_err(pg[name], prm.grad) if den > 0 else pg[name].abs().max().item(), f"{lst}.{n}.{name}", den))
rows.sort(reverse=True)
print("worst G gradient tensors (rel err, name, max|g|):")
for r in rows[:12]:
    print(f"  {r[0]:.2e} {r[1]} {r[2]:.2e}")
if hasattr(G_p, "spkr_embedding") and G_p.spkr_embedding.weight.grad is not None:
    print("  spkr_embedding", rel_err(G_p.spkr_embedding.weight.grad, G_o.spkr_embedding.weight.grad))
