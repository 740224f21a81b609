"""Generate the committed golden fixtures from the REAL reference (/root/reference).

Run in the build container only (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py

Writes
  ref_fixture_mlfb.npz  raw (int16; == float64*32768 exactly) and mlfb (float64, 1057x80) read from the
                        reference's own HDF5 fixture test/data/SF1/SF1_10001.feats.h5 (contiguous
                        uncompressed datasets; byte offsets found during the survey)
  ref_train_golden.npz  outputs of the reference's UNMODIFIED trainers / VQVAE2 (imported through
                        oracle/refshim.py, with the restated parallel_wavegan) on seeded synthetic
                        batches: per-step loss dicts, VQ indices, decoded-feature statistics and
                        parameter checksums for vqvae / lsgan / cyclegan / stargan.
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from crank_b200.conf import vcc2020_conf  # noqa: E402
from crank_b200.synthetic import clone_batch, make_batch, spkr_dict  # noqa: E402
from oracle import crank_port as cp  # noqa: E402
from oracle import refshim  # noqa: E402

GOLD_SPKRS = 14
GOLD_B, GOLD_T = 2, 96
GOLD_STEPS = 2
KINDS = ["vqvae", "lsgan", "cyclegan", "stargan"]


def golden_conf(kind):
    return vcc2020_conf(trainer_type=kind, n_steps_gan_start=-1, n_steps_cycle_start=-1,
                        discriminator_dropout=0.0)


def checksum(module):
    sd = module.state_dict()
    return np.array([sum(float(v.double().sum()) for v in sd.values()),
                     sum(float(v.double().abs().sum()) for v in sd.values())])


def fixture_mlfb():
    h5 = os.path.join(refshim.REFERENCE_ROOT, "test/data/SF1/SF1_10001.feats.h5")
    buf = open(h5, "rb").read()
    raw = np.frombuffer(buf[2048 : 2048 + 135294 * 8], dtype="<f8")
    mlfb = np.frombuffer(buf[1084400 : 1084400 + 1057 * 80 * 8], dtype="<f8").reshape(1057, 80)
    raw16 = np.round(raw * 32768).astype(np.int16)
    assert np.array_equal(raw16.astype(np.float64) / 32768, raw)
    np.savez_compressed(os.path.join(HERE, "ref_fixture_mlfb.npz"), raw_i16=raw16, mlfb=mlfb)


def train_golden():
    refshim.install()
    vq = refshim.ref("crank.net.module.vqvae2")
    spk = refshim.ref("crank.net.module.spkradv")
    tr = refshim.ref("crank.net.trainer")
    tu = refshim.ref("crank.net.trainer.utils")
    out = {}
    for kind in KINDS:
        conf = golden_conf(kind)
        random.seed(1234)
        np.random.seed(1234)
        torch.manual_seed(1234)      # crank/bin/train.py:49-51
        model = {"G": vq.VQVAE2(conf, spkr_size=GOLD_SPKRS),
                 "SPKRADV": spk.SpeakerAdversarialNetwork(conf, GOLD_SPKRS)}
        rest = cp.build_models(conf, GOLD_SPKRS)   # C and D are parallel_wavegan classes (restated)
        model["C"] = rest["C"]
        if "D" in rest:
            model["D"] = rest["D"]
        out[f"{kind}/init_checksum"] = np.stack([checksum(model[k]) for k in sorted(model)])
        opt = tu.get_optimizer(conf, model)
        crit = tu.get_criterion(conf, device="cpu")
        sch = tu.get_scheduler(conf, opt)
        W = refshim.NullWriter()
        T = tr.TrainerWrapper(kind, model=model, optimizer=opt, criterion=crit,
                              dataloader={"spkrs": spkr_dict(GOLD_SPKRS)}, writer={"train": W, "dev": W},
                              expdir="/tmp/exp", conf=conf, feat_conf=conf["feature"], scheduler=sch,
                              scaler=None, resume=0, device="cpu", n_jobs=1)
        T.tqdm.close()
        batch = make_batch(GOLD_B, GOLD_T, GOLD_SPKRS, seed=0, ragged=True)
        # G forward before any update (EMA fires: do it on a deep copy of the state afterwards restored)
        sd0 = {k: v.clone() for k, v in model["G"].state_dict().items()}
        with torch.no_grad():
            dec_h, spkrvec = T._get_dec_h(clone_batch(batch))
            o = model["G"].forward(batch["in_feats"], None, dec_h, spkrvec=spkrvec)
        model["G"].load_state_dict(sd0)
        out[f"{kind}/fwd_qidx0"] = o["qidx"][0].numpy()
        out[f"{kind}/fwd_qidx1"] = o["qidx"][1].numpy()
        out[f"{kind}/fwd_decoded"] = o["decoded"].numpy().astype(np.float32)
        for it in range(GOLD_STEPS):
            random.seed(100 + it)
            vals = T.train(clone_batch(batch), "train")
            keys = sorted(vals)
            out[f"{kind}/step{it}_keys"] = np.array(keys)
            out[f"{kind}/step{it}_vals"] = np.array([vals[k] for k in keys], dtype=np.float64)
        out[f"{kind}/final_checksum"] = np.stack([checksum(model[k]) for k in sorted(model)])
    np.savez_compressed(os.path.join(HERE, "ref_train_golden.npz"), **out)


if __name__ == "__main__":
    assert refshim.available(), "needs /root/reference"
    fixture_mlfb()
    train_golden()
    for f in ["ref_fixture_mlfb.npz", "ref_train_golden.npz"]:
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
