"""Whole-step CUDA graph on hardware (round-1 verdict item 3, ADVICE medium):
a replayed step must be BIT-IDENTICAL to the eager step -- loss values, every parameter, the Adam moments and the
VQ EMA buffers -- over several replays, i.e. the capture contains every weight-norm refresh, all-reduce-free update and
EMA kernel of the step and nothing reads a stale eager buffer."""
import copy
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


def _trainer(kind, S, seed=1234):
    from crank_b200.conf import vcc2020_conf
    from crank_b200.net.trainer import TrainerWrapper, get_criterion, get_model, get_optimizer, get_scheduler
    from crank_b200.synthetic import spkr_dict

    # dropout 0: eager and replayed steps would otherwise consume different Philox offsets
    conf = vcc2020_conf(trainer_type=kind, n_steps_gan_start=-1, n_steps_cycle_start=-1, discriminator_dropout=0.0)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    pm = get_model(conf, S, device="cuda")
    opt = get_optimizer(conf, pm)
    P = TrainerWrapper(kind, model=pm, optimizer=opt, criterion=get_criterion(conf), dataloader={"spkrs": spkr_dict(S)},
                       writer={"train": _W(), "dev": _W()}, expdir="/tmp/exp", conf=conf, feat_conf=conf["feature"],
                       scheduler=get_scheduler(conf, opt), scaler=None, resume=0, device="cuda", n_jobs=1)
    P.tqdm.close()
    return P


def _state(P):
    out = {}
    for k, m in P.model.items():
        for n, v in m.state_dict().items():
            out[f"{k}.{n}"] = v.detach().clone()
    for k, o in P.optimizer.items():
        for i, (p, st) in enumerate(o.state.items()):
            out[f"opt.{k}.{i}.exp_avg"] = st["exp_avg"].detach().clone()
            out[f"opt.{k}.{i}.exp_avg_sq"] = st["exp_avg_sq"].detach().clone()
    return out


@pytest.mark.parametrize("kind", ["lsgan", "vqvae"])
def test_graph_replay_is_bit_identical_to_eager(kind):
    from crank_b200.net.graph import GraphedTrainStep
    from crank_b200.synthetic import make_batch, to_device

    S, B, T, STEPS = 14, 4, 200, GraphedTrainStep.WARMUP + 1 + 5        # warm-up, capture call, 5 replays
    batches = [to_device(make_batch(B, T, S, seed=50 + i, ragged=True), "cuda") for i in range(STEPS)]
    clone = lambda b: {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in b.items()}  # noqa: E731

    Pe = _trainer(kind, S)
    Pg = _trainer(kind, S)
    for k in Pe.model:
        Pg.model[k].load_state_dict(Pe.model[k].state_dict())
    # the eager twin also keeps its Adam step count on the device so that both run the same optimizer kernel
    for o in Pe.optimizer.values():
        o.capturable = True
    stepper = GraphedTrainStep(Pg)
    for i in range(STEPS):
        ve = Pe.train(clone(batches[i]), "train")
        vg = stepper(clone(batches[i]))
        assert set(ve) == set(vg)
        for k in ve:
            assert ve[k] == vg[k], f"{kind} step {i} ({'replay' if i > GraphedTrainStep.WARMUP else 'eager/capture'}): " \
                                   f"loss {k}: eager {ve[k]!r} vs graphed {vg[k]!r}"
        se, sg = _state(Pe), _state(Pg)
        assert set(se) == set(sg)
        for k in se:
            assert torch.equal(se[k], sg[k]), f"{kind} step {i}: {k} differs between the eager and the graphed trainer"
    assert len(stepper._graphs) == 1
