"""StarGAN-style train step.  API mirror of crank/net/trainer/trainer_stargan.py:17-118."""

import random

import torch

from .trainer_lsgan import LSGANTrainer


class StarGANTrainer(LSGANTrainer):
    def update_G(self, batch, loss, phase="train"):
        cycle_outputs = self._cycle(batch)
        if self.conf["use_vqvae_loss"]:
            loss = self.calculate_vqvae_loss(batch, cycle_outputs[0]["org"], loss)
        loss = self.calculate_cyclevqvae_loss(batch, cycle_outputs, loss)
        if self.conf["use_spkradv_training"]:
            for label in ["cv", "recon"]:
                loss = self.calculate_spkradv_loss(batch, cycle_outputs[0][label], loss, label=label,
                                                   phase=phase)
        loss = self.calculate_adv_loss(batch, cycle_outputs[0]["cv"]["decoded"], batch["cv_h"],
                                       batch["decoder_mask"], loss)
        if phase == "train" and not self.stop_generator:
            self.step_model(loss, model="G")
        return loss

    def update_D(self, batch, loss, phase="train"):
        enc_h_cv = self._get_enc_h(batch, use_cvfeats=True)
        dec_h_cv, spkrvec_cv = self._get_dec_h(batch, use_cvfeats=True)
        updates = random.choice(["real", "fake"]) if self.conf["switch_update"] else ["real", "fake"]
        real = self._discriminate(self.get_D_inputs(batch, batch["in_feats"], label="org"))
        loss = self.calculate_discriminator_loss(real, batch["org_h"], batch["decoder_mask"], loss,
                                                 label="real", updates=updates)
        with torch.no_grad():       # only decoded.detach() is used
            outputs = self.model["G"].forward(batch["in_feats"], enc_h_cv, dec_h_cv, spkrvec_cv)
        fake = self._discriminate(self.get_D_inputs(batch, outputs["decoded"].detach(), label="cv"))
        loss = self.calculate_discriminator_loss(fake, batch["cv_h"], batch["decoder_mask"], loss,
                                                 label="fake", updates=updates)
        if phase == "train":
            self.step_model(loss, model="D")
        return loss
