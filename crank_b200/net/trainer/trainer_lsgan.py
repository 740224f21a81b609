"""VQ-VAE + LSGAN train step (BASELINE north-star workload) on the fused kernels.

API mirror of crank/net/trainer/trainer_lsgan.py:16-206: discriminator update (G forward, D on real
and detached fake), generator update (VQ-VAE losses, speaker-adversarial loss, second G forward for
the adversarial term), `train_first` ordering, `gan_flag` / `stop_generator` schedule, loss keys.
LSGAN targets are constants, so `MSE(D(x)[mask], 1)` is a masked sum/count reduction against a
scalar (no ones_like / masked_select tensors, trainer_lsgan.py:154-171).
"""

import os

import torch

from ... import ops
from .. import _dp
from .basetrainer import frozen
from .trainer_vqvae import VQVAETrainer


class LSGANTrainer(VQVAETrainer):
    def __init__(self, model, optimizer, criterion, dataloader, writer, expdir, conf, feat_conf,
                 scheduler=None, scaler=None, resume=0, device="cuda", n_jobs=-1):
        super().__init__(model, optimizer, criterion, dataloader, writer, expdir, conf, feat_conf,
                         scheduler=scheduler, scaler=scaler, resume=resume, device=device,
                         n_jobs=n_jobs)
        self.gan_flag = False
        self.cycle_flag = False
        self.stop_generator = False
        self.batch_D_passes = os.environ.get("CRANK_B200_BATCH_D", "1") != "0"      # update_D: one pass over [real | fake]
        self._check_cycle_start()
        self._check_gan_start()
        self.stop_generator = False  # the reference resets it after the checks (trainer_lsgan.py:53)

    def check_custom_start(self):
        self._check_cycle_start()
        self._check_gan_start()

    def _train_core(self, batch, phase="train"):
        _dp.begin_step(batch)
        loss = self._get_loss_dict()
        if self.gan_flag:
            loss = self.forward_lsgan(batch, loss, phase=phase)
        elif self.cycle_flag:
            loss = self.forward_cycle(batch, loss, phase=phase)
        else:
            loss = self.forward_vqvae(batch, loss, phase=phase)
        loss = self.forward_spkradv(batch, loss, phase=phase)
        loss = self.forward_spkrclassifier(batch, loss, phase=phase)
        return loss

    def forward_lsgan(self, batch, loss, phase="train"):
        order = ["G", "D"] if self.conf["train_first"] == "G" else ["D", "G"]
        for which in order:
            loss = (self.update_G if which == "G" else self.update_D)(batch, loss, phase=phase)
        loss["objective"] = loss["G"] + loss["D"]
        return loss

    def _discriminate(self, inputs):
        return self.model["D"].forward_cl(inputs)

    def update_G(self, batch, loss, phase="train"):
        enc_h = self._get_enc_h(batch)
        dec_h, spkrvec = self._get_dec_h(batch)
        feats = batch["in_feats"]
        outputs = self.model["G"].forward(feats, enc_h, dec_h, spkrvec)
        loss = self.calculate_vqvae_loss(batch, outputs, loss)
        if self.conf["use_spkradv_training"]:
            loss = self.calculate_spkradv_loss(batch, outputs, loss, phase=phase)
        if self.conf["cvadv_flag"]:
            dec_h, spkrvec = self._get_dec_h(batch, use_cvfeats=True)
            h = batch["cv_h"]
        else:
            h = batch["org_h"]
        adv_outputs = self.model["G"].forward(
            feats, enc_h, dec_h, spkrvec=spkrvec, use_ema=not self.conf["encoder_detach"],
            encoder_detach=self.conf["encoder_detach"])
        loss = self.calculate_adv_loss(batch, adv_outputs["decoded"], h, batch["decoder_mask"], loss)
        if phase == "train" and not self.stop_generator:
            self.step_model(loss, model="G")
        return loss

    def update_D(self, batch, loss, phase="train"):
        enc_h = self._get_enc_h(batch)
        mask = batch["decoder_mask"]
        if self.conf["cvadv_flag"]:
            dec_h, spkrvec = self._get_dec_h(batch, use_cvfeats=True)
            h = batch["cv_h"]
        else:
            dec_h, spkrvec = self._get_dec_h(batch)
            h = batch["org_h"]
        with torch.no_grad():       # only decoded.detach() is used below: no autograd graph, no saved activations
            outputs = self.model["G"].forward(batch["in_feats"], enc_h, dec_h, spkrvec)
        real_in = self.get_D_inputs(batch, batch["in_feats"], label="org")
        fake_in = self.get_D_inputs(batch, outputs["decoded"].detach(), label="cv")
        if self.batch_D_passes:
            # ONE discriminator pass over [real | fake] along the batch axis instead of the reference's two
            # (trainer_lsgan.py:168-178): the network has no cross-utterance coupling, so the outputs are the same
            # numbers and the parameter gradient is the same sum; half the launches of the D update (~70 of ~900 per step)
            # and one weight-gradient partial-sum epilogue instead of two.  (With dropout the two halves draw their masks
            # from one RNG call instead of two.)
            nb = real_in.shape[0]
            both = self._discriminate(torch.cat([real_in, fake_in], dim=0))
            real, fake = both[:nb], both[nb:]
        else:
            real = self._discriminate(real_in)
            fake = self._discriminate(fake_in)
        loss = self.calculate_discriminator_loss(real, batch["org_h"], mask, loss, label="real")
        loss = self.calculate_discriminator_loss(fake, h, mask, loss, label="fake")
        if phase == "train":
            self.step_model(loss, model="D")
        return loss

    def calculate_adv_loss(self, batch, decoded, h, mask, loss):
        with frozen(self.model["D"]):       # generator update: only the gradient w.r.t. the decoded features is used
            fake = self._discriminate(self.get_D_inputs(batch, decoded, label="cv"))
        if self.conf["acgan_flag"]:
            fake, spkr_cls = torch.split(fake, [1, self.n_spkrs], dim=2)
            loss = self.calculate_acgan_loss(spkr_cls, h, loss)
        loss["D_adv"] = ops.masked_l1_mse(fake, 1.0, mask)[1]
        loss["G"] += self.conf["alpha"]["adv"] * loss["D_adv"]
        return loss

    def calculate_discriminator_loss(self, sample, h, mask, loss, label="real", updates=None):
        if self.conf["acgan_flag"]:
            sample, spkr_cls = torch.split(sample, [1, self.n_spkrs], dim=2)
            loss = self.calculate_acgan_loss(spkr_cls, h, loss, label=label, model="D")
        target = 1.0 if label == "real" else 0.0
        loss[f"D_{label}"] = ops.masked_l1_mse(sample, target, mask)[1]
        if updates is None or label in updates:
            loss["D"] += self.conf["alpha"][label] * loss[f"D_{label}"]
        return loss

    def calculate_acgan_loss(self, spkr_cls, h, loss, label="adv", model="G"):
        loss[f"D_acgan_{label}"] = self.criterion["ce"](spkr_cls.reshape(-1, spkr_cls.size(2)), h.reshape(-1))
        if not (self.conf["use_real_only_acgan"] and label == "fake"):
            loss[model] += self.conf["alpha"]["acgan"] * loss[f"D_acgan_{label}"]
        return loss

    def _check_gan_start(self):
        if self.steps > self.conf["n_steps_gan_start"]:
            self.gan_flag = True
            if self.conf["n_steps_stop_generator"] > 0:
                self.stop_generator = True
        if self.steps > self.conf["n_steps_gan_start"] + self.conf["n_steps_stop_generator"]:
            self.stop_generator = False

    def get_D_inputs(self, batch, feats, label="org"):
        """[feats | uv | speaker code] -> (B, T, 80+1+32) (trainer_lsgan.py:194-206)."""
        parts = [feats]
        if self.conf["use_D_uv"]:
            parts.append(batch["uv"])
        if self.conf["use_D_spkrcode"]:
            if not self.conf["use_spkr_embedding"]:
                parts.append(batch[f"{label}_h_onehot"])
            else:
                h = batch[f"{label}_h"]
                h = h[:, 0:1].expand_as(h)  # drop the ignore_index padding
                _dp.wait_for((self.model["G"].spkr_embedding.weight,))
                parts.append(self.model["G"].spkr_embedding(h).detach())
        return torch.cat(parts, dim=-1).float()
