"""GPU parity AT THE BASELINE SHAPES against the CPU oracle (round-1 verdict, parity item 1a).

BASELINE.json configs 2-4 with the shipped production arithmetic (3xTF32 tcgen05):
  config 2  VCC2018 trainer_type=vqvae,  16 utts x 500 frames, 12 speakers
  config 3  VCC2020 trainer_type=lsgan,  64 utts x 500 frames, 14 speakers (the bench workload)
            + the same shapes with trainer_type=vqvae
  config 4  VCC2020 trainer_type=cyclegan, 16 utts x 500 frames
The oracle (`oracle/crank_port.py`, pinned bit-identical to the live reference trainers) runs the same step on the
host cores: a 64 x 500 LSGAN step takes a couple of seconds there.

Protocol: the oracle warms the EMA codebooks with three generator passes, its state is
loaded into the product, then
  * one more generator pass on both: VQ code indices must be IDENTICAL; should a frame differ it has to be a
    floating-point near-tie of the reference's own distance expression (<= 8 ulp), and at most 2 per stack;
  * one full train step on both: every loss key within 1e-4 relative (the north star's tolerance).
Discriminator dropout is 0 (torch's Philox stream cannot be reproduced); masks are ragged.
"""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _W:
    def add_scalar(self, *a, **k):
        pass

    def flush(self):
        pass

    def close(self):
        pass


def _pair(kind, S, seed=1234):
    from crank_b200.conf import vcc2020_conf
    from crank_b200.net.trainer import TrainerWrapper, get_criterion, get_model, get_optimizer, get_scheduler
    from crank_b200.synthetic import spkr_dict
    from oracle import crank_port as cp

    conf = vcc2020_conf(trainer_type=kind, n_steps_gan_start=-1, n_steps_cycle_start=-1, discriminator_dropout=0.0)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    om = cp.build_models(conf, S)
    pm = get_model(conf, S, device="cuda")
    O = cp.OracleTrainer(kind, om, cp.build_optimizers(conf, om), conf)
    opt = get_optimizer(conf, pm)
    P = TrainerWrapper(kind, model=pm, optimizer=opt, criterion=get_criterion(conf), dataloader={"spkrs": spkr_dict(S)},
                       writer={"train": _W(), "dev": _W()}, expdir="/tmp/exp", conf=conf, feat_conf=conf["feature"],
                       scheduler=get_scheduler(conf, opt), scaler=None, resume=0, device="cuda", n_jobs=1)
    P.tqdm.close()
    return conf, om, pm, O, P


def _near_tie(q, x_flat, idx_ref, idx_ours):
    """both indices minimise the reference's fp32 distance expression to within 8 ulp (vqvae2.py:340-344)"""
    w = q.embedding.weight.detach()
    d = (w.pow(2).sum(1)[None, :] - 2 * x_flat @ w.t()) + x_flat.pow(2).sum(1, keepdim=True)
    a, b = d[0, idx_ref].item(), d[0, idx_ours].item()
    ulp = np.spacing(np.float32(max(abs(a), abs(b))))
    return abs(a - b) <= 8 * float(ulp)


@pytest.mark.parametrize("kind,S,B,T", [
    ("vqvae", 12, 16, 500),       # BASELINE config 2
    ("lsgan", 14, 64, 500),       # BASELINE config 3 (bench workload)
    ("vqvae", 14, 64, 500),
    ("cyclegan", 14, 16, 500),    # BASELINE config 4
])
def test_baseline_shape_step_matches_cpu_oracle(kind, S, B, T):
    from crank_b200 import lib as L
    from crank_b200.synthetic import clone_batch, make_batch, to_device

    L.set_precision("tf32x3")
    conf, om, pm, O, P = _pair(kind, S)
    batch = make_batch(B, T, S, seed=0, ragged=True)
    warm = make_batch(B, T, S, seed=7, ragged=True)

    # ---- EMA-warmed codebooks on the oracle, then identical state on both sides ----
    with torch.no_grad():
        dec_h, spk = O._dec_h(clone_batch(warm))
        for _ in range(3):
            om["G"].forward(warm["in_feats"], None, dec_h, spkrvec=spk)
    for k in om:
        pm[k].load_state_dict(om[k].state_dict())

    # ---- one generator pass: VQ indices ----
    qin = []
    hooks = [q.register_forward_hook(lambda m, inp, out, qin=qin: qin.append(inp[0].detach().clone()))
             for q in om["G"].quantizers]
    with torch.no_grad():
        dec_h, spk = O._dec_h(clone_batch(batch))
        oo = om["G"].forward(batch["in_feats"], None, dec_h, spkrvec=spk)
        bd = to_device(clone_batch(batch), "cuda")
        dec_hp, spkp = P._get_dec_h(bd)
        po = pm["G"].forward(bd["in_feats"], None, dec_hp, spkrvec=spkp)
    for h in hooks:
        h.remove()
    # decode() runs the top stack first: qin = [stack n-1, ..., stack 0]; inputs are (B, 64, T)
    nst = conf["n_vq_stacks"]
    for n in range(nst):
        ours, ref = po["qidx"][n].cpu(), oo["qidx"][n]
        bad = (ours != ref).nonzero()
        print(f"{kind} {B}x{T}: stack {n}: {len(bad)} of {ref.numel()} VQ indices differ from the oracle")
        assert len(bad) <= 2, f"qidx{n}: {len(bad)} mismatches"
        x = qin[nst - 1 - n].transpose(1, 2)                  # (B, T, 64)
        for b_, t_ in bad.tolist():
            assert _near_tie(om["G"].quantizers[n], x[b_, t_][None, :], int(ref[b_, t_]), int(ours[b_, t_])), \
                f"qidx{n}[{b_},{t_}] differs and is not a floating-point near-tie"
    e = ((po["decoded"].cpu().double() - oo["decoded"].double()).abs().max() / oo["decoded"].double().abs().max()).item()
    print(f"{kind} {B}x{T}: decoded rel err {e:.2e}")
    assert e <= 1e-4
    # identical EMA state again before the train step (a near-tie frame would have moved one code's mean slightly)
    for k in om:
        pm[k].load_state_dict(om[k].state_dict())

    # ---- one full train step ----
    random.seed(100)
    ov = O.train(clone_batch(batch), "train")
    random.seed(100)
    pv = P.train(to_device(clone_batch(batch), "cuda"), "train")
    assert set(ov) == set(pv), set(ov) ^ set(pv)
    worst = 0.0
    for k in sorted(ov):
        ref = ov[k]
        err = abs(pv[k] - ref) / max(abs(ref), 1e-12) if ref != 0 else abs(pv[k])
        worst = max(worst, err)
        assert err <= 1e-4, f"{kind} {B}x{T} loss {k}: product {pv[k]} vs oracle {ref} (rel {err:.2e})"
    print(f"{kind} {B}x{T}: {len(ov)} loss keys, worst rel err {worst:.2e}")


def _generator_gradient_errors(precision, tc_disable):
    """rel err (max-norm, per tensor) of every generator-loss gradient vs the oracle at BASELINE config 2's shape."""
    from crank_b200 import lib as L
    from crank_b200.synthetic import clone_batch, make_batch, to_device
    from tests.util import rel_err

    L.set_precision(precision)
    kind, S, B, T = "vqvae", 12, 16, 500
    conf, om, pm, O, P = _pair(kind, S)
    batch = make_batch(B, T, S, seed=0, ragged=True)
    warm = make_batch(B, T, S, seed=7, ragged=True)
    # EMA-warmed codebooks: with the fresh-init codebook U(-1/512, 1/512) a single near-tie frame whose
    # index flips moves gradient elements by ~2e-3 -- measured on the oracle itself, fp32 against float64.
    with torch.no_grad():
        dec_h, spk = O._dec_h(clone_batch(warm))
        for _ in range(3):
            om["G"].forward(warm["in_feats"], None, dec_h, spkrvec=spk)
    for k in om:
        pm[k].load_state_dict(om[k].state_dict())

    b = clone_batch(batch)
    dec_h, spk = O._dec_h(b)
    o = om["G"].forward(b["in_feats"], O._enc_h(b), dec_h, spkrvec=spk)
    lo = {"G": 0.0}
    O._vqvae_loss(b, o, lo)
    O._spkradv_loss(b, o, lo)
    lo["G"].backward()

    L.check(L.lib().crk_debug_tc_disable(tc_disable), "tc_disable")
    try:
        bp = to_device(clone_batch(batch), "cuda")
        dec_hp, spkp = P._get_dec_h(bp)
        po = pm["G"].forward(bp["in_feats"], P._get_enc_h(bp), dec_hp, spkrvec=spkp)
        lp = P.calculate_vqvae_loss(bp, po, P._get_loss_dict())
        lp = P.calculate_spkradv_loss(bp, po, lp)
        lp["G"].backward()
    finally:
        L.check(L.lib().crk_debug_tc_disable(0), "tc_disable")
        L.set_precision("tf32x3")
    assert abs(float(lp["G"]) - float(lo["G"])) <= 1e-4 * abs(float(lo["G"]))
    for n in range(conf["n_vq_stacks"]):
        assert torch.equal(po["qidx"][n].cpu(), o["qidx"][n]), f"qidx{n} differs: the gradient comparison would be moot"
    rows = []
    for lst in ("encoders", "decoders"):
        for s in range(conf["n_vq_stacks"]):
            pg = getattr(pm["G"], lst)[s].named_conv_grads()
            for name, prm in getattr(om["G"], lst)[s].named_parameters():
                assert name in pg, f"{lst}.{s}.{name}: no gradient in the product"
                if prm.grad is None:
                    continue
                if prm.grad.abs().max().item() < 1e-10:
                    assert pg[name].abs().max().item() < 1e-7, f"{lst}.{s}.{name} should be zero"
                    continue
                rows.append((rel_err(pg[name], prm.grad), f"{lst}.{s}.{name}"))
    rows.append((rel_err(pm["G"].spkr_embedding.weight.grad, om["G"].spkr_embedding.weight.grad), "spkr_embedding.weight"))
    rows.sort(reverse=True)
    return rows


@pytest.mark.parametrize("mode", ["fp32 kernels", "tensor-core backward on the fp32 forward", "all tensor cores"])
def test_generator_gradients_match_cpu_oracle_before_adam(mode):
    """Every gradient tensor of the generator loss (l1 + stft + commit + speaker-adversarial through the GRL) at
    BASELINE config 2's shape (16 x 500, ragged), BEFORE any optimizer step (round-1 verdict, weak item 4).

    What the three modes pin, and why the last bound is looser (measured: profiles/diag_r2c.py, profiles/diag_r2b.py):
      * fp32 CUDA-core kernels: every tensor <= 1e-4 of its largest element (measured <= 5e-5);
      * the tensor-core BACKWARD families (dgrad, gate backward, wgrad in 3xTF32) running on the activations the fp32
        forward saved: the same 1e-4 -- their arithmetic reproduces the fp32 kernels to ~1e-5;
      * everything on tensor cores: the forward's own rounding noise (~1e-6 relative with the round-to-nearest split)
        moves a handful of the ~4 M ReLU / LeakyReLU inputs of a pass across zero; each such element switches its
        derivative between 0 and 1 -- a discontinuity of the loss gradient itself, which any two implementations with
        different rounding hit at different elements -- and one flipped element of an 8 000-frame reduction moves a
        weight-gradient column by ~1e-2 of its norm, i.e. ~1e-3 of the tensor's largest element.  So: 5e-3 per tensor,
        and 1e-4 for the median tensor."""
    import statistics

    prec, mask = {"fp32 kernels": ("fp32", 0), "tensor-core backward on the fp32 forward": ("tf32x3", 1),
                  "all tensor cores": ("tf32x3", 0)}[mode]
    rows = _generator_gradient_errors(prec, mask)
    med = statistics.median(e for e, _ in rows)
    print(f"{mode}: {len(rows)} gradient tensors, median rel err {med:.2e}; worst: " + ", ".join(f"{n} {e:.1e}" for e, n in rows[:6]))
    if mode == "all tensor cores":
        assert rows[0][0] <= 5e-3, rows[0]
        assert med <= 1e-4, med
    else:
        assert rows[0][0] <= 1e-4, rows[0]
