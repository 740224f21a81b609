"""VQ-VAE train step on the fused kernels.

API mirror of crank/net/trainer/trainer_vqvae.py:20-369: same `train(batch, phase)` ordering
(generator step -> speaker-adversarial step -> speaker-classifier step), same loss-dict keys and
weights.  The masked MSE/L1 terms use sum/count reduction kernels instead of
`masked_select` + `MSELoss` (trainer_vqvae.py:229-237), which removes the dynamic shapes and
their host syncs while computing the same means.
"""

import torch

from ... import ops
from .. import _dp
from .basetrainer import BaseTrainer, frozen, pick_cv_speakers


class VQVAETrainer(BaseTrainer):
    def __init__(self, model, optimizer, criterion, dataloader, writer, expdir, conf, feat_conf,
                 scheduler=None, scaler=None, resume=0, device="cuda", n_jobs=-1):
        super().__init__(model, optimizer, criterion, dataloader, writer, expdir, conf, feat_conf,
                         scheduler=scheduler, scaler=scaler, resume=resume, device=device,
                         n_jobs=n_jobs)
        self.cycle_flag = False
        self._check_cycle_start()

    def check_custom_start(self):
        self._check_cycle_start()

    # ---- public steps --------------------------------------------------------------------------
    def train(self, batch, phase="train"):
        loss = self._train_core(batch, phase)
        loss_values = self._parse_loss(loss)
        self._flush_writer(loss, phase)
        return loss_values

    def _train_core(self, batch, phase="train"):
        """Everything of a step that runs on the device, no host synchronisation: returns the loss dict of 0-dim
        tensors (`_parse_loss` fetches them with one copy).  This is the unit a CUDA graph captures (net/graph.py)."""
        _dp.begin_step(batch)
        loss = self._get_loss_dict()
        if self.cycle_flag:
            loss = self.forward_cycle(batch, loss, phase=phase)
        else:
            loss = self.forward_vqvae(batch, loss, phase=phase)
        loss = self.forward_spkradv(batch, loss, phase=phase)
        loss = self.forward_spkrclassifier(batch, loss, phase=phase)
        return loss

    @torch.no_grad()
    def dev(self, batch):
        loss_values = self.train(batch, phase="dev")
        for cv_spkr_name in pick_cv_speakers(self.spkrs, self.n_cv_spkrs):
            outputs = self._convert(batch, cv_spkr_name)
            self._generate_cvwav(batch, outputs, cv_spkr_name, tdir="dev_wav", save_hdf5=False,
                                 n_samples=self.n_dev_samples)
        return loss_values

    @torch.no_grad()
    def reconstruction(self, batch, tdir="reconstruction"):
        outputs = self._convert(batch, None)
        self._generate_cvwav(batch, outputs, None, tdir=tdir, save_hdf5=True, save_decoded=False,
                             n_samples=-1)
        return outputs

    @torch.no_grad()
    def eval(self, batch):
        results = {}
        for cv_spkr_name in self.spkrs.keys():
            outputs = self._convert(batch, cv_spkr_name)
            self._generate_cvwav(batch, outputs, cv_spkr_name, tdir="eval_wav", save_hdf5=True,
                                 save_decoded=False, n_samples=-1)
            results[cv_spkr_name] = outputs
        return results

    def _feats(self, batch):
        return batch["in_feats"] if not self.conf["use_raw"] else batch["raw"]

    def _convert(self, batch, cv_spkr_name):
        enc_h = self._get_enc_h(batch, cv_spkr_name=cv_spkr_name)
        dec_h, spkrvec = self._get_dec_h(batch, cv_spkr_name=cv_spkr_name)
        return self.model["G"](self._feats(batch), enc_h, dec_h, spkrvec=spkrvec)

    # ---- sub-steps -----------------------------------------------------------------------------
    def forward_vqvae(self, batch, loss, phase="train"):
        enc_h = self._get_enc_h(batch)
        dec_h, spkrvec = self._get_dec_h(batch)
        outputs = self.model["G"].forward(self._feats(batch), enc_h, dec_h, spkrvec=spkrvec)
        loss = self.calculate_vqvae_loss(batch, outputs, loss)
        if self.conf["use_spkradv_training"]:
            loss = self.calculate_spkradv_loss(batch, outputs, loss, label="org", phase=phase)
        loss["objective"] += loss["G"]
        if phase == "train":
            self.step_model(loss, model="G")
        return loss

    def forward_cycle(self, batch, loss, phase="train"):
        cycle_outputs = self._cycle(batch)
        if self.conf["use_vqvae_loss"]:
            loss = self.calculate_vqvae_loss(batch, cycle_outputs[0]["org"], loss)
        loss = self.calculate_cyclevqvae_loss(batch, cycle_outputs, loss)
        if self.conf["use_spkradv_training"]:
            for label in ["cv", "recon"]:
                loss = self.calculate_spkradv_loss(batch, cycle_outputs[0][label], loss,
                                                   label=label, phase=phase)
        loss["objective"] += loss["G"]
        if phase == "train":
            self.step_model(loss, model="G")
        return loss

    def _cycle(self, batch):
        enc_h = self._get_enc_h(batch)
        enc_h_cv = self._get_enc_h(batch, use_cvfeats=True)
        dec_h, spkrvec = self._get_dec_h(batch)
        dec_h_cv, spkrvec_cv = self._get_dec_h(batch, use_cvfeats=True)
        return self.model["G"].cycle_forward(self._feats(batch), enc_h, dec_h, enc_h_cv, dec_h_cv,
                                             spkrvec, spkrvec_cv)

    def _encoder_outputs(self, outputs):
        """(list of encoder outputs with the causal warm-up region dropped, #dropped frames)."""
        if self.conf["causal"]:
            er = self.model["G"].encoder_receptive_size
            return [e[:, er:] for e in outputs["encoded_unmod"]], er
        return outputs["encoded_unmod"], 0

    def forward_spkradv(self, batch, loss, phase="train"):
        if self.conf["use_spkradv_training"]:
            enc_h = self._get_enc_h(batch)
            dec_h, spkrvec = self._get_dec_h(batch)
            # the reference re-runs G here (trainer_vqvae.py:168); kept: the EMA codebook update
            # fires on every G forward, so dropping the pass would change the training trajectory.
            # Only the (detached) encoder outputs are read: no autograd graph, and the bottom decoder
            # stack -- stateless, feeding no quantiser -- is not evaluated
            with torch.no_grad():
                outputs = self.model["G"].forward(self._feats(batch), enc_h, dec_h, spkrvec=spkrvec,
                                                  final_decoder=False)
            encoded, er = self._encoder_outputs(outputs)
            logits = self.model["SPKRADV"].forward(encoded, detach=True)
            ce = self.criterion["ce"](logits.reshape(-1, logits.size(2)), batch["org_h"][:, er:].reshape(-1))
            loss["SPKRADV"] = self.conf["alpha"]["ce"] * ce
            if phase == "train":
                self.step_model(loss, model="SPKRADV")
        return loss

    def _classify(self, feats):
        return self.model["C"].forward_cl(feats)

    def forward_spkrclassifier(self, batch, loss, phase="train"):
        if self.conf["use_spkr_classifier"]:
            real = self._classify(batch["in_feats"])
            loss["C_real"] = self.criterion["ce"](real.reshape(-1, real.size(2)), batch["org_h"].reshape(-1))
            loss["C"] += self.conf["alpha"]["ce"] * loss["C_real"]
            if phase == "train":
                self.step_model(loss, model="C")
        return loss

    # ---- losses --------------------------------------------------------------------------------
    def _shift(self, causal_size):
        return causal_size if self.conf["causal"] else 0

    def _vq_terms(self, loss, outputs, emask, suffix=""):
        for n in range(self.conf["n_vq_stacks"]):
            enc, emb = outputs["encoded"][n], outputs["emb_idx"][n]
            loss[f"G_commit{n}{suffix}"] = ops.masked_l1_mse(enc, emb.detach(), emask)[1]
            if not self.conf["ema_flag"]:
                loss[f"G_dict{n}{suffix}"] = ops.masked_l1_mse(emb, enc.detach(), emask)[1]
        return loss

    def calculate_vqvae_loss(self, batch, outputs, loss):
        cs = self.conf["causal_size"]
        decoded, target = outputs["decoded"], batch["out_feats"]
        loss["G_l1"], loss["G_mse"] = ops.masked_l1_mse(decoded, target, batch["decoder_mask"], self._shift(cs))
        loss["G_stft"] = self.criterion["fstft"](decoded, target, causal_size=cs)
        loss = self._vq_terms(loss, outputs, batch["encoder_mask"])
        return self._parse_vqvae_loss(loss)

    def calculate_cyclevqvae_loss(self, batch, outputs, loss):
        for c in range(self.conf["n_cycles"]):
            for io in ["cv", "recon"]:
                lbl = f"{c}cyc_{io}"
                o = outputs[c][io]
                if io == "cv":
                    emask = batch["encoder_mask"]
                    with frozen(self.model["C"]):       # part of the generator loss: classifier gradients are not used
                        fake = self._classify(o["decoded"])
                    loss[f"C_fake_{lbl}"] = self.criterion["ce"](
                        fake.reshape(-1, fake.size(2)), batch["cv_h"].reshape(-1))
                else:
                    emask = batch["cycle_encoder_mask"]
                    cs = self.conf["causal_size"] * 2 if self.conf["causal"] else 0
                    loss[f"G_l1_{lbl}"], loss[f"G_mse_{lbl}"] = ops.masked_l1_mse(
                        o["decoded"], batch["in_feats"], batch["cycle_decoder_mask"], self._shift(cs))
                    loss[f"G_stft_{lbl}"] = self.criterion["fstft"](o["decoded"], batch["in_feats"], causal_size=cs)
                loss = self._vq_terms(loss, o, emask, suffix=f"_{lbl}")
        return self._parse_cyclevqvae_loss(loss)

    def calculate_spkradv_loss(self, batch, outputs, loss, label="org", phase="train"):
        encoded, er = self._encoder_outputs(outputs)
        with frozen(self.model["SPKRADV"]):         # generator loss: only the (reversed) encoder gradient is used
            logits = self.model["SPKRADV"].forward(encoded)
        loss[f"G_spkradv_{label}"] = self.criterion["ce"](
            logits.reshape(-1, logits.size(2)), batch["org_h"][:, er:].reshape(-1))
        w = self.conf["alpha"]["ce"]
        if label == "recon":
            w = self.conf["alpha"]["cycle"] * w
        loss["G"] += w * loss[f"G_spkradv_{label}"]
        return loss

    def _parse_vqvae_loss(self, loss):
        a = self.conf["alpha"]
        for k in ["l1", "mse", "stft"]:
            loss["G"] += a[k] * loss[f"G_{k}"]
        for k in ["commit"] + ([] if self.conf["ema_flag"] else ["dict"]):
            for n in range(self.conf["n_vq_stacks"]):
                loss["G"] += a[k] * loss[f"G_{k}{n}"]
        return loss

    def _parse_cyclevqvae_loss(self, loss):
        a = self.conf["alpha"]
        for c in range(self.conf["n_cycles"]):
            for io in ["cv", "recon"]:
                lbl = f"{c}cyc_{io}"
                for n in range(self.conf["n_vq_stacks"]):
                    loss["G"] += a["cycle"] * a["commit"] * loss[f"G_commit{n}_{lbl}"]
                    if not self.conf["ema_flag"]:
                        loss["G"] += a["cycle"] * a["dict"] * loss[f"G_dict{n}_{lbl}"]
                if io == "recon":
                    for k in ["l1", "mse", "stft"]:
                        loss["G"] += a["cycle"] * a[k] * loss[f"G_{k}_{lbl}"]
                else:
                    loss["G"] += a["cycle"] * a["ce"] * loss[f"C_fake_{lbl}"]
        return loss

    def _check_cycle_start(self):
        if self.conf["use_cyclic_training"] and self.steps > self.conf["n_steps_cycle_start"]:
            self.cycle_flag = True
        if self.conf["use_cyclic_training"] and not self.conf["use_spkr_classifier"]:
            raise ValueError("use_cyclic_training requires use_spkr_classifier to be true")
