"""Round-2 diagnostic b: per-loss-term gradient parity of the generator at 16 x 500 (ragged), product vs oracle."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_gpu_baseline_shapes import _pair
from crank_b200 import lib as L
from crank_b200.synthetic import clone_batch, make_batch, to_device
from tests.util import rel_err

L.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32x3")
B, T = int(sys.argv[2]) if len(sys.argv) > 2 else 16, int(sys.argv[3]) if len(sys.argv) > 3 else 500
ragged = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
kind, S = "vqvae", 12
conf, om, pm, O, P = _pair(kind, S)
batch = make_batch(B, T, S, seed=0, ragged=ragged)
warm = make_batch(B, T, S, seed=7, ragged=ragged)
with torch.no_grad():
    dh_, sp_ = O._dec_h(clone_batch(warm))
    for _ in range(3):
        om["G"].forward(warm["in_feats"], None, dh_, spkrvec=sp_)
for k in om:
    pm[k].load_state_dict(om[k].state_dict())


def grads(mods):
    out = {}
    G = mods["G"]
    for lst in ("encoders", "decoders"):
        for n in range(2):
            net = getattr(G, lst)[n]
            if hasattr(net, "named_conv_grads"):
                for k, v in net.named_conv_grads().items():
                    out[f"{lst}.{n}.{k}"] = v.detach().cpu().clone()
            else:
                for k, prm in net.named_parameters():
                    if prm.grad is not None:
                        out[f"{lst}.{n}.{k}"] = prm.grad.detach().clone()
    if G.spkr_embedding.weight.grad is not None:
        out["spkr_embedding"] = G.spkr_embedding.weight.grad.detach().cpu().clone()
    return out


terms = ["G_l1", "G_mse", "G_stft", "G_commit0", "G_commit1", "G_spkradv_org", "G"]
for term in terms:
    res = []
    for side in ("oracle", "product"):
        mods = om if side == "oracle" else pm
        for m in mods.values():
            m.zero_grad(set_to_none=True)
        if side == "oracle":
            sd = {k: {n: v.clone() for n, v in om[k].state_dict().items()} for k in om}
            b = clone_batch(batch)
            dec_h, spk = O._dec_h(b)
            o = om["G"].forward(b["in_feats"], O._enc_h(b), dec_h, spkrvec=spk)
            loss = {"G": 0.0}
            O._vqvae_loss(b, o, loss)
            O._spkradv_loss(b, o, loss)
            loss[term].backward()
            for k in om:
                om[k].load_state_dict(sd[k])
        else:
            sd = {k: {n: v.clone() for n, v in pm[k].state_dict().items()} for k in pm}
            b = to_device(clone_batch(batch), "cuda")
            dec_h, spk = P._get_dec_h(b)
            o = pm["G"].forward(b["in_feats"], P._get_enc_h(b), dec_h, spkrvec=spk)
            loss = P.calculate_vqvae_loss(b, o, P._get_loss_dict())
            loss = P.calculate_spkradv_loss(b, o, loss)
            loss[term].backward()
            for k in pm:
                pm[k].load_state_dict(sd[k])
        res.append((float(loss[term]), grads(mods), [q.detach().cpu().clone() for q in o["qidx"]]))
    (lo, go, qo), (lp, gp, qp) = res
    print("   qidx mismatches:", [int((a != b).sum()) for a, b in zip(qo, qp)])
    rows = sorted(((rel_err(gp[k], go[k]) if go[k].abs().max() > 0 else float(gp[k].abs().max()), k, float(go[k].abs().max()), float(gp[k].abs().max()))
                   for k in go if k in gp), reverse=True)
    missing = [k for k in go if k not in gp]
    print(f"{term}: loss oracle {lo:.6f} product {lp:.6f}; {len(rows)} tensors, worst:")
    for r in rows[:4]:
        print(f"    {r[0]:.2e} {r[1]}  max|g| oracle {r[2]:.2e} product {r[3]:.2e}")
    if missing:
        print("    missing in product:", missing[:5])
