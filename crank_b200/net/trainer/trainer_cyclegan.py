"""Cyclic VQ-VAE + LSGAN train step.  API mirror of crank/net/trainer/trainer_cyclegan.py:18-179
(update_G :52-76, update_D :78-93, cycle adversarial / discriminator losses :95-179, including the
host-side `random.choice` of the fake branch :166)."""

import random

import torch

from ... import ops
from .basetrainer import frozen
from .trainer_lsgan import LSGANTrainer


class CycleGANTrainer(LSGANTrainer):
    def update_G(self, batch, loss, phase="train"):
        cycle_outputs = self._cycle(batch)
        loss = self.calculate_vqvae_loss(batch, cycle_outputs[0]["org"], loss)
        loss = self.calculate_cyclevqvae_loss(batch, cycle_outputs, loss)
        if self.conf["use_spkradv_training"]:
            loss = self.calculate_spkradv_loss(batch, cycle_outputs[0]["org"], loss, phase=phase)
        loss = self.calculate_cycleadv_loss(batch, cycle_outputs, loss)
        if phase == "train" and not self.stop_generator:
            self.step_model(loss, model="G")
        return loss

    def update_D(self, batch, loss, phase="train"):
        with torch.no_grad():       # every use below is .detach()-ed
            outputs = self._cycle(batch)
        loss = self.calculate_cycle_discriminator_loss(batch, outputs, loss)
        if phase == "train":
            self.step_model(loss, model="D")
        return loss

    def calculate_cycleadv_loss(self, batch, outputs, loss):
        mask = batch["decoder_mask"]
        for c in range(self.conf["n_cycles"]):
            for io in ["org", "cv"]:
                lbl = f"{c}cyc_{io}"
                with frozen(self.model["D"]):
                    D_out = self._discriminate(self.get_D_inputs(batch, outputs[c][io]["decoded"], label="cv"))
                if self.conf["acgan_flag"]:
                    D_out, spkr_cls = torch.split(D_out, [1, self.n_spkrs], dim=2)
                    loss[f"D_acgan_adv_{lbl}"] = self.criterion["ce"](
                        spkr_cls.reshape(-1, spkr_cls.size(2)), batch[f"{io}_h"].reshape(-1))
                    loss["G"] += self.conf["alpha"]["acgan"] * loss[f"D_acgan_adv_{lbl}"]
                    m = mask
                else:
                    m = None  # the reference only masks this term on the acgan branch (:108-118)
                loss[f"D_adv_{lbl}"] = ops.masked_l1_mse(D_out, 1.0, m)[1]
                loss["G"] += self.conf["alpha"]["adv"] * loss[f"D_adv_{lbl}"]
        return loss

    def calculate_cycle_discriminator_loss(self, batch, outputs, loss):
        a = self.conf["alpha"]
        for c in range(self.conf["n_cycles"]):
            lbl = f"{c}cyc"
            sample = {
                "real": self._discriminate(self.get_D_inputs(batch, batch["in_feats"], label="org")),
                "org_fake": self._discriminate(
                    self.get_D_inputs(batch, outputs[0]["org"]["decoded"].detach(), label="org")),
                "cv_fake": self._discriminate(
                    self.get_D_inputs(batch, outputs[0]["cv"]["decoded"].detach(), label="cv")),
            }
            if self.conf["acgan_flag"]:
                for k in list(sample.keys()):
                    h = batch["org_h"] if k in ["real", "org_fake"] else batch["cv_h"]
                    sample[k], spkr_cls = torch.split(sample[k], [1, self.n_spkrs], dim=2)
                    loss[f"D_ce_{k}_{lbl}"] = self.criterion["ce"](
                        spkr_cls.reshape(-1, spkr_cls.size(2)), h.reshape(-1))
                    if not (self.conf["use_real_only_acgan"] and k == "org_fake"):
                        loss["D"] += a["acgan"] * loss[f"D_ce_{k}_{lbl}"]
            loss[f"D_real_{lbl}"] = ops.masked_l1_mse(sample["real"], 1.0, batch["decoder_mask"])[1]
            fake_key = random.choice(["org_fake", "cv_fake"])
            mask = batch["cycle_decoder_mask"] if fake_key == "org_fake" else batch["decoder_mask"]
            loss[f"D_fake_{lbl}"] = ops.masked_l1_mse(sample[fake_key], 0.0, mask)[1]
            loss["D"] += a["fake"] * loss[f"D_fake_{lbl}"] + a["real"] * loss[f"D_real_{lbl}"]
        return loss
