"""Device batcher on cuda:0: bit-exact against the reference dataset's golden batch, and usable as the input of a
train step."""
import os
import random

import numpy as np
import pytest
import torch

from tests.test_cpu_data import BATCH_LEN, GOLD, IDX, SPKRS, TENSOR_KEYS, compare, oracle_batch, product_batch

pytestmark = pytest.mark.gpu


def test_device_batcher_on_gpu_matches_oracle_and_reference_golden():
    prod = product_batch("cuda:0")
    assert prod["in_feats"].is_cuda and prod["org_h"].is_cuda
    orc, _, _ = oracle_batch()
    compare(prod, orc)
    g = np.load(GOLD, allow_pickle=False)
    for k in TENSOR_KEYS:
        a = prod[k].cpu().numpy()
        assert np.array_equal(a.astype(g[k].dtype), g[k]), k


def test_train_step_consumes_a_device_batch():
    from crank_b200.conf import vcc2020_conf
    from crank_b200.data import DeviceBatcher, UtteranceStore
    from crank_b200.net.trainer import TrainerWrapper, get_criterion, get_model, get_optimizer, get_scheduler
    from oracle import dataset_port as dp

    utts, scaler = dp.make_corpus(12, SPKRS, seed=1)
    store = UtteranceStore(utts, SPKRS, scaler, device="cuda:0")
    conf = vcc2020_conf(trainer_type="lsgan", n_steps_gan_start=-1)
    torch.manual_seed(0)
    random.seed(0)
    models = get_model(conf, len(SPKRS), device="cuda:0")
    opt = get_optimizer(conf, models)

    class W:
        def add_scalar(self, *a, **k): pass
        def flush(self): pass
        def close(self): pass

    tr = TrainerWrapper("lsgan", model=models, optimizer=opt, criterion=get_criterion(conf),
                        dataloader={"spkrs": {s: i for i, s in enumerate(SPKRS)}}, writer={"train": W(), "dev": W()},
                        expdir="/tmp/crank_b200_data", conf=conf, feat_conf=conf["feature"],
                        scheduler=get_scheduler(conf, opt), scaler=None, resume=0, device="cuda:0", n_jobs=1)
    tr.tqdm.close()
    for batch in DeviceBatcher(store, BATCH_LEN).epoch(4):
        vals = tr.train(batch, "train")
        assert all(np.isfinite(v) for v in vals.values()), vals
        assert vals["G_l1"] > 0 and vals["D_real"] > 0
