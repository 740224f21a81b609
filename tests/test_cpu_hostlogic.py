"""Host logic of the product (trainers, VQVAE2 / Quantizer orchestration, flat-parameter state dicts, FusedAdam
plumbing, frozen / no-grad passes) for every recipe switch, WITHOUT a GPU: the C-ABI ops are swapped for torch-CPU
stand-ins (tests/cpu_emulation.py, test infrastructure) and the product's train steps are compared with the oracle
port, which is itself bit-identical to the unmodified reference trainers (tests/test_cpu_oracle.py).
The kernels behind the ops are verified separately on the GPU (`-m gpu`)."""
import random

import numpy as np
import pytest
import torch

from tests.cpu_emulation import emulated_ops

S = 5


class _W:
    def add_scalar(self, *a, **k): pass
    def flush(self): pass
    def close(self): pass


def _run(kind, overrides, T, steps=2):
    from crank_b200.conf import vcc2020_conf
    from crank_b200.net.trainer import TrainerWrapper, get_criterion, get_model, get_optimizer, get_scheduler
    from crank_b200.synthetic import clone_batch, make_batch, spkr_dict
    from oracle import crank_port as cp

    conf = vcc2020_conf(trainer_type=kind, n_steps_gan_start=-1, discriminator_dropout=0.0, **overrides)
    torch.manual_seed(11)
    om = cp.build_models(conf, S)
    O = cp.OracleTrainer(kind, om, cp.build_optimizers(conf, om), conf)
    with emulated_ops():
        pm = get_model(conf, S, device="cpu")
        for k in om:
            pm[k].load_state_dict(om[k].state_dict())
        opt = get_optimizer(conf, pm)
        P = TrainerWrapper(kind, model=pm, optimizer=opt, criterion=get_criterion(conf),
                           dataloader={"spkrs": spkr_dict(S)}, writer={"train": _W(), "dev": _W()},
                           expdir="/tmp/crank_b200_hostlogic", conf=conf, feat_conf=conf["feature"],
                           scheduler=get_scheduler(conf, opt), scaler=None, resume=0, device="cpu", n_jobs=1)
        P.tqdm.close()
        worst = 0.0
        for step in range(steps):
            b = make_batch(2, T, S, seed=20 + step, ragged=True)
            random.seed(step)
            ov = O.train(clone_batch(b), "train")
            random.seed(step)
            pv = P.train(clone_batch(b), "train")
            assert set(ov) <= set(pv), sorted(set(ov) - set(pv))
            for k, ref in ov.items():
                err = abs(pv[k] - ref) / max(abs(ref), 1e-3)
                worst = max(worst, err)
                # the stand-ins are not bit-identical to the oracle's modules, and Adam turns rounding-level
                # gradient components into +-lr updates, so quantities evaluated after an update inside the step
                # agree to ~1e-4; a wrong mask / weight / missing term shows up at 1e-2 .. 1
                assert err <= 3e-4, f"{kind} {overrides} step {step}: {k}: product {pv[k]} vs oracle {ref}"
        # every parameter / buffer after the steps, through the reference-keyed state dict
        for k in om:
            sd_o, sd_p = om[k].state_dict(), pm[k].state_dict()
            assert set(sd_o) == set(sd_p), (k, sorted(set(sd_o) ^ set(sd_p))[:6])
            for name, t in sd_o.items():
                d = (sd_p[name].double() - t.double()).abs().max().item()
                scale = t.double().abs().max().item()
                # Adam normalises the update: a gradient component that is ~0 moves by +-lr whatever its rounding, so
                # single elements may differ by up to 2 * lr * steps = 8e-4 in absolute terms (weights are ~0.3)
                assert d <= max(5e-4 * scale, 1e-3), f"{kind} {overrides}: {k}.{name} differs by {d:.2e} (max |ref| {scale:.2e})"
    return worst


@pytest.mark.parametrize("kind", ["vqvae", "lsgan", "cyclegan", "stargan"])
def test_product_train_steps_match_oracle_with_emulated_ops(kind):
    _run(kind, {}, T=96)


@pytest.mark.parametrize("kind,overrides", [
    ("lsgan", dict(causal=True, causal_size=4)),
    ("cyclegan", dict(causal=True, causal_size=4)),
    ("vqvae", dict(n_vq_stacks=1)),
    ("lsgan", dict(use_spkr_embedding=False)),
    ("cyclegan", dict(encoder_f0=True)),
    ("lsgan", dict(ema_flag=False)),
    ("cyclegan", dict(acgan_flag=True)),
    ("lsgan", dict(acgan_flag=True)),
    ("stargan", dict(cvadv_flag=True)),
    ("lsgan", dict(encoder_detach=True)),
    ("lsgan", dict(train_first="G")),
    ("vqvae", dict(use_cyclic_training=True, n_steps_cycle_start=-1)),
])
def test_product_host_logic_on_config_variants(kind, overrides):
    _run(kind, overrides, T=176 if overrides.get("causal") else 96, steps=1 if overrides.get("causal") else 2)


def test_three_vq_stacks_fail_loudly():
    """n_vq_stacks = 3 makes the bottom decoder's input 192 channels wide; the kernels take at most 128, and the
    product says so at construction instead of computing something else."""
    from crank_b200.conf import vcc2020_conf
    from crank_b200.net.trainer import get_model

    with emulated_ops():
        with pytest.raises(NotImplementedError, match="128"):
            get_model(vcc2020_conf(trainer_type="vqvae", n_vq_stacks=3), S, device="cpu")


def test_emulation_is_test_only_and_restored():
    from crank_b200 import lib, ops
    from crank_b200.parallel_wavegan import models

    before = (models.WavenetFn, ops.VQFn, ops.MaskedLossFn, lib.require_cuda)
    with emulated_ops():
        assert models.WavenetFn is not before[0]
    assert (models.WavenetFn, ops.VQFn, ops.MaskedLossFn, lib.require_cuda) == before
    with pytest.raises(lib.CrkError):
        ops.masked_l1_mse(torch.zeros(1, 2, 3), 0.0)          # the product still refuses CPU tensors
