"""Trainer base: step loop, schedulers, checkpointing, conditioning vectors, loss bookkeeping.

API mirror of crank/net/trainer/basetrainer.py:26-309 (TrainerWrapper, BaseTrainer).  What is NOT
rebuilt: the HDF5 dumping and WORLD synthesis of dev/eval/reconstruction (basetrainer.py:322-435, out of scope per
SURVEY.md section 2 #7).  Griffin-Lim synthesis of mlfb outputs (`_generate_cvwav`) runs on the device.

Differences by design: every loss scalar of a step is fetched with ONE packed device->host copy
(`_parse_loss`), replacing the reference's ~15 `.item()` syncs (basetrainer.py:208-231).
"""

import logging
import random
from pathlib import Path

import torch

from ...synthetic import to_device
from .. import _dp


def TrainerWrapper(trainer_type, **ka):
    from . import CycleGANTrainer, LSGANTrainer, StarGANTrainer, VQVAETrainer

    table = {"vqvae": VQVAETrainer, "lsgan": LSGANTrainer, "cyclegan": CycleGANTrainer,
             "stargan": StarGANTrainer}
    if trainer_type not in table:
        raise NotImplementedError("conf['trainer_type']: {} is not supported.".format(trainer_type))
    return table[trainer_type](**ka)


class _NoBar:
    def update(self, n=1):
        pass

    def close(self):
        pass


class BaseTrainer(object):
    def __init__(self, model, optimizer, criterion, dataloader, writer, expdir, conf, feat_conf,
                 scheduler=None, scaler=None, resume=0, device="cuda", n_jobs=-1):
        self.model = model
        self.optimizer = optimizer
        self.criterion = criterion
        self.dataloader = dataloader
        self.writer = writer
        self.expdir = Path(expdir)
        self.conf = conf
        self.feat_conf = feat_conf
        self.scheduler = scheduler
        self.scaler = scaler
        self.device = device
        self.n_jobs = n_jobs

        self.spkrs = dataloader["spkrs"]
        self.n_spkrs = len(self.spkrs)
        self.n_cv_spkrs = 4 if self.n_spkrs > 4 else self.n_spkrs
        self.n_dev_samples = 5

        self.resume_steps = resume
        self.steps = resume
        self._sched_to(self.steps)
        self.finish_train = False
        try:
            from tqdm import tqdm

            self.tqdm = tqdm(initial=self.steps, total=self.conf["n_steps"], desc="train",
                             disable=not logging.getLogger().isEnabledFor(logging.INFO))
        except Exception:  # pragma: no cover
            self.tqdm = _NoBar()

    # ---- abstract surface ------------------------------------------------------------------
    def train(self, batch, phase="train"):
        raise NotImplementedError

    def dev(self, batch):
        raise NotImplementedError

    def eval(self, batch):
        raise NotImplementedError

    def reconstruction(self, batch, tdir="reconstruction"):
        raise NotImplementedError

    def check_custom_start(self):
        pass

    # ---- loop ------------------------------------------------------------------------------
    def run(self, flag="train", tdir=None):
        self.flag = flag
        if flag == "train":
            while not self.finish_train:
                self._tr_step()
            self.tqdm.close()
            self.writer["train"].close()
            self.writer["dev"].close()
            logging.info("Finish training")
        else:
            self._run_eval(flag, tdir)

    def save_model(self):
        _dp.flush()
        if _dp.active() and _dp.rank() != 0:      # one writer per checkpoint file (replicas are identical)
            return
        checkpoint = self.expdir / "checkpoint_{}steps.pkl".format(self.steps)
        state = {"steps": self.steps, "model": {"G": self.model["G"].state_dict()}}
        for m in ["SPKRADV", "D", "C"]:
            if m in self.model:
                state["model"][m] = self.model[m].state_dict()
        torch.save(state, checkpoint)

    def _run_eval(self, flag="eval", tdir=False):
        self.tqdm.close()
        if flag == "eval":
            for batch in self.dataloader["eval"]:
                self.eval(to_device(batch, self.device))
        if flag == "reconstruction":
            for dkey in ["train", "dev"]:
                for batch in self.dataloader[dkey]:
                    self.reconstruction(to_device(batch, self.device), tdir="reconstruction")

    def _tr_step(self):
        for batch in self.dataloader["train"]:
            batch = to_device(batch, self.device)
            loss_values = self.train(batch, phase="train")
            if self.steps % self.conf["n_steps_print_loss"] == 0:
                self._print_loss_values(loss_values, phase="train")
            self._dev_step()
            self._check_save_model()
            self._step_update()
            self._check_finish()
            self.check_custom_start()
            if self.finish_train:
                break

    def _dev_step(self):
        if (self.steps % self.conf["dev_steps"] == 0 and self.steps > self.conf["dev_steps"] - 1
                and self.steps != self.resume_steps):
            dev_loss_values = self._get_loss_dict()
            for dev_idx, batch in enumerate(self.dataloader["dev"]):
                dev_loss_values = self.dev(to_device(batch, self.device))
                if dev_idx > 0:
                    break
            self._print_loss_values(dev_loss_values, phase="dev")

    # ---- loss bookkeeping --------------------------------------------------------------------
    def _get_loss_dict(self):
        return {"objective": 0.0, "G": 0.0, "D": 0.0, "C": 0.0, "SPKRADV": 0.0}

    def _parse_loss(self, loss):
        """dict of python floats; all tensor entries fetched with a single D2H copy."""
        values = self._get_loss_dict()
        keys = [k for k, v in loss.items() if isinstance(v, torch.Tensor)]
        if keys:
            packed = _dp.average_loss_vector(torch.stack([loss[k].detach().reshape(()).float() for k in keys])).tolist()
            for k, v in zip(keys, packed):
                values[k] = values.get(k, 0.0) + v
        for k in loss:
            values.setdefault(k, 0.0)
        self._last_loss_values = values
        return values

    def _print_loss_values(self, loss_values, phase="train"):
        logging.info("{} iterations: {}".format(phase, self.steps))
        for k, v in sorted(loss_values.items()):
            if v != 0.0:
                logging.info("{}: {}".format(k, v))

    def _flush_writer(self, loss, phase):
        if _dp.active() and _dp.rank() != 0:      # rank 0 owns the event log (the values are rank-averaged)
            return
        if self.steps % self.conf["n_steps_print_loss"] == 0:
            values = getattr(self, "_last_loss_values", None) or self._parse_loss(loss)
            for k, v in loss.items():
                if isinstance(v, torch.Tensor):
                    self.writer[phase].add_scalar("loss/{}".format(k), values[k], self.steps)
            self.writer[phase].flush()

    def _check_save_model(self):
        if self.resume_steps != self.steps and self.steps % self.conf["n_steps_save_model"] == 0:
            self.save_model()

    def _sched_to(self, steps):
        if self.scheduler is None:
            return
        scheds = self.scheduler.values() if isinstance(self.scheduler, dict) else [self.scheduler]
        for s in scheds:
            # StepLR closed form (the reference calls the deprecated scheduler.step(epoch))
            s.last_epoch = int(steps)
            for group, base_lr in zip(s.optimizer.param_groups, s.base_lrs):
                group["lr"] = base_lr * s.gamma ** (int(steps) // s.step_size)
            s._last_lr = [g["lr"] for g in s.optimizer.param_groups]

    def _step_update(self):
        self.steps += 1
        self.tqdm.update(1)
        self._sched_to(self.steps)

    def _check_finish(self):
        if self.steps > self.conf["n_steps"]:
            self.finish_train = True

    # ---- optimisation --------------------------------------------------------------------------
    def step_model(self, loss, model="G"):
        """zero_grad -> backward -> (DP all-reduce) -> clip -> Adam   (trainer_vqvae.py:200-208)."""
        self.optimizer[model].zero_grad()
        loss[model].backward()
        params = [p for p in self.model[model].parameters()]
        clip = self.conf["optim"][model]["clip_grad_norm"]

        def reduce_and_update():
            _dp.average_gradients(params)
            if clip != 0:
                torch.nn.utils.clip_grad_norm_(params, clip)
            self.optimizer[model].step()

        # data parallel: on the communication stream (the next reader of these parameters waits for it)
        _dp.run_async(params, reduce_and_update)

    # ---- conditioning vectors (basetrainer.py:253-309) -----------------------------------------
    def _get_enc_h(self, batch, use_cvfeats=False, cv_spkr_name=None):
        if self.conf["encoder_f0"]:
            return self._get_f0_condition(batch, cv_spkr_name, use_cvfeats)
        return None

    def _get_dec_h(self, batch, use_cvfeats=False, cv_spkr_name=None):
        h, h_onehot = self._get_spkr_conditions(batch, cv_spkr_name, use_cvfeats)
        f0 = self._get_f0_condition(batch, cv_spkr_name, use_cvfeats) if self.conf["decoder_f0"] else None
        if not self.conf["use_spkr_embedding"]:
            return (torch.cat([f0, h_onehot], dim=-1) if f0 is not None else h_onehot), None
        return f0, h

    def _get_f0_condition(self, batch, cv_spkr_name, use_cvfeats=False):
        if cv_spkr_name is not None:
            lcf0 = self._get_cvf0(batch, cv_spkr_name)
        else:
            lcf0 = batch["cv_lcf0"] if use_cvfeats else batch["lcf0"]
        return torch.cat([lcf0, batch["uv"]], dim=-1)

    def _get_spkr_conditions(self, batch, cv_spkr_name, use_cvfeats=False):
        if cv_spkr_name is not None:
            B, T, _ = batch["in_feats"].size()
            num = self.spkrs[cv_spkr_name]
            h = torch.full((B, T), num, dtype=torch.long, device=batch["in_feats"].device)
            h_onehot = torch.nn.functional.one_hot(h, self.n_spkrs).float()
        else:
            key = "cv" if use_cvfeats else "org"
            h_onehot = batch[f"{key}_h_onehot"]
            h = batch[f"{key}_h"]
        # remove ignore_index (-100 on padded frames): every frame takes the utterance's label
        h = h[:, 0:1].expand_as(h).contiguous()
        return h, h_onehot

    def _get_cvf0(self, batch, spkr_name):
        """log-F0 mean/variance conversion org -> target speaker (dataset.py:290-293) on device."""
        sc = self.scaler
        out = []
        for n in range(batch["in_feats"].size(0)):
            org = sc[batch["org_spkr_name"][n]]["lcf0"]
            cv = sc[spkr_name]["lcf0"]
            glob = sc["lcf0"]
            lcf0 = batch["lcf0"][n] * float(glob.scale_[0]) + float(glob.mean_[0])
            conv = (float(cv.scale_[0]) / float(org.scale_[0])) * (lcf0 - float(org.mean_[0])) + float(cv.mean_[0])
            out.append((conv - float(glob.mean_[0])) / float(glob.scale_[0]))
        return torch.stack(out, dim=0).float()

    def _generate_cvwav(self, batch, outputs, cv_spkr_name=None, tdir="dev_wav", save_hdf5=True, save_decoded=True,
                        n_samples=1):
        """Decoded features -> waveforms (basetrainer.py:322-420), mlfb outputs only: the inverse scaler and Griffin-Lim
        run on the DEVICE (crank_b200.utils.griffin_lim; the reference hands each utterance to a joblib CPU worker
        running librosa, basetrainer.py:400-418), utterances of equal length are iterated as one batch, 16-bit WAVs are
        written with the stdlib `wave` module.  Not rebuilt (out of scope): HDF5 dumps
        (`save_hdf5`, needs h5py) and WORLD synthesis of mcep outputs -- both are skipped with a log line.
        Returns {wav path: float32 waveform tensor on the device} (the reference returns nothing)."""
        if not save_decoded:
            return {}
        if self.conf["output_feat_type"] != "mlfb":
            logging.info("crank_b200: WORLD synthesis of mcep outputs is not rebuilt; skipping %s", tdir)
            return {}
        import wave

        import numpy as np

        from ...utils import mlfb2wav

        tdir = self.expdir / tdir / str(self.steps)
        fc = self.feat_conf
        dec = outputs["decoded"]
        names = []
        for n in range(dec.size(0)):
            org = batch["org_spkr_name"][n]
            cv = org if cv_spkr_name is None else cv_spkr_name
            names.append(tdir / f"{batch['flbl'][n]}_org-{org}_cv-{cv}.wav")
        picked = list(range(dec.size(0)))
        if not (n_samples == -1 or n_samples > len(picked)):
            picked = random.sample(picked, n_samples)
        sc = None
        if self.scaler is not None and "mlfb" in self.scaler and "mlfb" not in self.conf.get("ignore_scaler", []):
            ms = self.scaler["mlfb"]
            sc = (torch.as_tensor(np.asarray(ms.mean_), dtype=torch.float32, device=dec.device),
                  torch.as_tensor(np.asarray(ms.scale_), dtype=torch.float32, device=dec.device))
        by_len = {}
        for n in picked:
            by_len.setdefault(int(batch["flen"][n]), []).append(n)
        wavs = {}
        for flen, ids in by_len.items():
            feats = torch.stack([dec[n, :flen] for n in ids]).float()
            if sc is not None:
                feats = feats * sc[1] + sc[0]                       # StandardScaler.inverse_transform
            y = mlfb2wav(feats, fs=fc["fs"], n_mels=fc["mlfb_dim"], fftl=fc["fftl"], win_length=fc["win_length"],
                         hop_size=fc["hop_size"], fmin=fc["fmin"], fmax=fc["fmax"])
            for j, n in enumerate(ids):
                wavs[names[n]] = y[j]
        for path, y in wavs.items():
            Path(path).parent.mkdir(parents=True, exist_ok=True)
            pcm = (y.clamp(-1.0, 1.0) * 32767.0).round().to(torch.int16).cpu().numpy()
            with wave.open(str(path), "wb") as f:
                f.setnchannels(1)
                f.setsampwidth(2)
                f.setframerate(int(fc["fs"]))
                f.writeframes(pcm.tobytes())
        return wavs


import contextlib


@contextlib.contextmanager
def frozen(*modules):
    """Run forwards whose PARAMETER gradients nobody reads (the discriminator / classifiers evaluated inside the
    generator update: the reference accumulates those gradients and discards them at the module's next
    zero_grad, crank/net/trainer/trainer_lsgan.py:146-160) without building them: the backward of such a
    forward only propagates input gradients, i.e. skips every weight-gradient kernel."""
    ps = [p for m in modules for p in m.parameters() if p.requires_grad]
    for p in ps:
        p.requires_grad_(False)
    try:
        yield
    finally:
        for p in ps:
            p.requires_grad_(True)


def pick_cv_speakers(spkrs, n):
    return random.sample(list(spkrs.keys()), n)
