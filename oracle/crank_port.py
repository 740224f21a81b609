"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of crank's train-step math.

Plain PyTorch (CPU, fp32, autograd) restatement of the reference's hot path, written so that it
can run on the GPU box where /root/reference does not exist:

  Quantizer / VQVAE2           <- crank/net/module/vqvae2.py:38-347
  SpeakerAdversarialNetwork    <- crank/net/module/spkradv.py:20-81
  feature / STFT losses        <- crank/net/module/loss.py:18-114
  train steps (vqvae / lsgan / cyclegan / stargan)
                               <- crank/net/trainer/trainer_{vqvae,lsgan,cyclegan,stargan}.py
  conditioning vectors         <- crank/net/trainer/basetrainer.py:253-309

Pinning: tests/test_oracle_vs_reference.py runs THIS file against the reference's own unmodified
classes (through oracle/refshim.py) in the build container and requires identical losses /
parameters after a step (same torch ops => bit-identical or 1-ulp), and tests/golden/*.npz holds
outputs of the real reference that this file must reproduce anywhere.  The conv stacks themselves
come from oracle/pwg.py (restated third-party package; see its header for the parity status).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (crank_b200/) never does.
"""

import random

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pwg


# ------------------------------------------------------------------------------------------------
class Quantizer(nn.Module):
    """vqvae2.py:286-347 (bdt_flag=True: input (B,D,T))."""

    def __init__(self, emb_dim, emb_size, decay=0.99, eps=1e-5, ema_flag=False, bdt_flag=False):
        super().__init__()
        self.emb_dim, self.emb_size = emb_dim, emb_size
        self.ema_flag, self.bdt_flag = ema_flag, bdt_flag
        self.embedding = nn.Embedding(emb_size, emb_dim)
        self.embedding.weight.data.uniform_(-1.0 / emb_size, 1.0 / emb_size)
        if ema_flag:
            self.decay, self.eps = decay, eps
            self.register_buffer("ema_size", torch.zeros(emb_size))
            self.register_buffer("ema_w", torch.randn(emb_dim, emb_size))

    def vq(self, x):
        flat = x.reshape(-1, self.emb_dim)
        w = self.embedding.weight
        dist = (
            torch.sum(torch.pow(w, 2), dim=1)
            - 2 * torch.matmul(flat, w.T)
            + torch.sum(torch.pow(flat, 2), dim=1, keepdim=True)
        )
        idx = torch.argmin(dist, dim=1).view(x.size(0), x.size(1))
        return idx, F.one_hot(idx, self.emb_size)

    def forward(self, x, use_ema=True):
        if self.bdt_flag:
            x = x.transpose(1, 2)
        idx, onehot = self.vq(x)
        embed_idx = torch.matmul(onehot.float(), self.embedding.weight)
        if self.training and self.ema_flag and use_ema:
            self.ema_size = self.decay * self.ema_size + (1 - self.decay) * torch.sum(
                onehot.view(-1, self.emb_size), 0
            )
            embed_sum = torch.sum(torch.matmul(x.transpose(1, 2), onehot.float()), dim=0)
            self.ema_w.data = self.decay * self.ema_w.data + (1 - self.decay) * embed_sum
            n = torch.sum(self.ema_size)
            self.ema_size = (self.ema_size + self.eps) / (n + self.emb_size * self.eps) * n
            self.embedding.weight.data.copy_((self.ema_w / self.ema_size.unsqueeze(0)).transpose(0, 1))
        qx = x + (embed_idx - x).detach()
        if self.bdt_flag:
            qx = qx.transpose(1, 2)
        return embed_idx, qx, idx


class VQVAE2(nn.Module):
    """vqvae2.py:38-283 without the raw / sinc front ends (use_raw is exercised separately)."""

    def __init__(self, conf, spkr_size=0, scaler=None):
        super().__init__()
        self.conf, self.spkr_size = conf, spkr_size
        self.encoder_receptive_size = 0
        self.decoder_receptive_size = 0
        c = conf
        self.encoders, self.decoders, self.quantizers = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for n in range(c["n_vq_stacks"]):
            if n == 0:
                e_in, e_out, e_aux = c["input_size"], c["emb_dim"][0], (2 if c["encoder_f0"] else 0)
                d_in = sum(c["emb_dim"][i] for i in range(c["n_vq_stacks"]))
                d_out = c["output_size"]
                d_aux = (2 if c["decoder_f0"] else 0) + (
                    c["spkr_embedding_size"] if c["use_spkr_embedding"] else spkr_size
                )
            else:
                e_in, e_out, e_aux = c["emb_dim"][n - 1], c["emb_dim"][n], 0
                d_in, d_out, d_aux = c["emb_dim"][n], c["emb_dim"][n - 1], 0
            kw = dict(
                kernel_size=c["kernel_size"][n],
                layers=c["n_layers"][n] * c["n_layers_stacks"][n],
                stacks=c["n_layers_stacks"][n],
                residual_channels=64, gate_channels=128, skip_channels=64, aux_context_window=0,
                dropout=0.0, bias=True, use_weight_norm=True, use_causal_conv=c["causal"],
                upsample_conditional_features=False,
            )
            self.encoders.append(pwg.ParallelWaveGANGenerator(in_channels=e_in, out_channels=e_out, aux_channels=e_aux, **kw))
            self.decoders.append(pwg.ParallelWaveGANGenerator(in_channels=d_in, out_channels=d_out, aux_channels=d_aux, **kw))
            self.encoder_receptive_size += self.encoders[-1].receptive_field_size
            self.decoder_receptive_size += self.decoders[-1].receptive_field_size
            self.quantizers.append(Quantizer(c["emb_dim"][n], c["emb_size"][n], ema_flag=c["ema_flag"], bdt_flag=True))
        if c["use_spkr_embedding"]:
            self.spkr_embedding = nn.Embedding(spkr_size, c["spkr_embedding_size"])

    @staticmethod
    def _t(h):
        return h.transpose(1, 2) if h is not None else None

    def _get_dec_h(self, dec_h, spkrvec):
        if spkrvec is not None:
            emb = self.spkr_embedding(spkrvec)
            dec_h = emb if dec_h is None else torch.cat([dec_h, emb], axis=-1)
        return dec_h

    def encode(self, x, enc_h=None):
        out = []
        for n in range(self.conf["n_vq_stacks"]):
            enc = self.encoders[n](x, c=enc_h) if n == 0 else self.encoders[n](enc, c=None)
            out.append(enc)
        return out

    def decode(self, enc, dec_h, use_ema=True, detach=False):
        dec = 0
        emb_idxs, qxs, qidxs = [], [], []
        for n in reversed(range(self.conf["n_vq_stacks"])):
            enc[n] = enc[n] + dec            # mutates the caller's list (vqvae2.py:177)
            emb_idx, qx, qidx = self.quantizers[n](enc[n], use_ema=use_ema)
            if detach:
                qx = qx.detach()
            emb_idxs.append(emb_idx)
            qxs.append(qx)
            qidxs.append(qidx)
            if n != 0:
                dec = self.decoders[n](qx, c=None)
            else:
                dec = self.decoders[n](torch.cat(qxs, dim=1), c=dec_h)
        return enc, dec, emb_idxs, qxs, qidxs

    @staticmethod
    def make_dict(enc, dec, emb_idxs, qidxs, enc_unmod):
        return {
            "encoded": [e.transpose(1, 2) for e in enc],
            "encoded_unmod": [e.transpose(1, 2) for e in enc_unmod] if enc_unmod is not None else None,
            "decoded": dec.transpose(1, 2),
            "emb_idx": emb_idxs[::-1],
            "qidx": qidxs[::-1],
        }

    def forward(self, x, enc_h, dec_h, spkrvec=None, use_ema=True, encoder_detach=False):
        x = x.transpose(1, 2)
        dec_h = self._t(self._get_dec_h(dec_h, spkrvec))
        enc = self.encode(x, enc_h=self._t(enc_h))
        enc_unmod = [e.clone() for e in enc]
        enc, dec, emb_idxs, _, qidxs = self.decode(enc, dec_h, use_ema=use_ema, detach=encoder_detach)
        return self.make_dict(enc, dec, emb_idxs, qidxs, enc_unmod)

    def cycle_forward(self, x, org_enc_h, org_dec_h, cv_enc_h, cv_dec_h, org_spkrvec, cv_spkrvec):
        x = x.transpose(1, 2)
        org_dec_h = self._t(self._get_dec_h(org_dec_h, org_spkrvec))
        cv_dec_h = self._t(self._get_dec_h(cv_dec_h, cv_spkrvec))
        org_enc_h, cv_enc_h = self._t(org_enc_h), self._t(cv_enc_h)
        outputs = []
        for _ in range(self.conf["n_cycles"]):
            enc = self.encode(x, enc_h=org_enc_h)
            org_unmod = [e.clone() for e in enc]
            cv_unmod = [e.clone() for e in enc]
            org_enc, org_dec, org_emb, _, org_q = self.decode(enc, org_dec_h)
            cv_enc, cv_dec, cv_emb, _, cv_q = self.decode(enc, cv_dec_h)
            enc = self.encode(cv_dec, enc_h=cv_enc_h)
            rec_unmod = [e.clone() for e in enc]
            rec_enc, rec_dec, rec_emb, _, rec_q = self.decode(enc, org_dec_h)
            outputs.append({
                "org": self.make_dict(org_enc, org_dec, org_emb, org_q, org_unmod),
                "cv": self.make_dict(cv_enc, cv_dec, cv_emb, cv_q, cv_unmod),
                "recon": self.make_dict(rec_enc, rec_dec, rec_emb, rec_q, rec_unmod),
            })
            x = rec_dec.clone().detach()
        return outputs


# ------------------------------------------------------------------------------------------------
class _GRL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        ctx.save_for_backward(scale)
        return x

    @staticmethod
    def backward(ctx, g):
        (scale,) = ctx.saved_tensors
        return scale * -g, None


class _GRLLayer(nn.Module):  # keeps the attribute path `grl.scale` (a plain tensor, not a buffer)
    def __init__(self, scale):
        super().__init__()
        self.scale = torch.tensor(scale)

    def forward(self, x):
        return _GRL.apply(x, self.scale)


class SpeakerAdversarialNetwork(nn.Module):
    """spkradv.py:20-60."""

    def __init__(self, conf, spkr_size=0):
        super().__init__()
        self.conf, self.spkr_size = conf, spkr_size
        self.grl = _GRLLayer(conf["spkradv_lambda"])
        self.classifier = pwg.ParallelWaveGANDiscriminator(
            in_channels=sum(conf["emb_dim"][: conf["n_vq_stacks"]]), out_channels=spkr_size,
            kernel_size=conf["spkradv_kernel_size"], layers=conf["n_spkradv_layers"],
            conv_channels=64, dilation_factor=1, nonlinear_activation="LeakyReLU",
            nonlinear_activation_params={"negative_slope": 0.2}, bias=True, use_weight_norm=True,
        )

    def forward(self, x, detach=False):
        x = torch.cat(x, axis=-1)
        if detach:
            x = x.detach()
        x = self.grl(x).transpose(1, 2)
        return self.classifier(x).transpose(1, 2)


# ------------------------------------------------------------------------------------------------
def feature_loss(kind, x, y, mask=None, causal=False, causal_size=0):
    """CustomFeatureLoss.forward for l1 / mse (loss.py:30-47)."""
    if causal:
        if causal_size > 0:
            x, y = x[:, causal_size:], y[:, :-causal_size]
            mask = mask[:, causal_size:] if mask is not None else None
        elif causal_size < 0:
            cs = -causal_size
            y, x = y[:, cs:], x[:, :-cs]
            mask = mask[:, :-cs] if mask is not None else None
    if mask is not None:
        x, y = x.masked_select(mask), y.masked_select(mask)
    return F.l1_loss(x, y) if kind == "l1" else F.mse_loss(x, y)


def stft_mag(x, n_fft, hop_length, win_length):
    """loss.py:50-60 with torch.stft's real parameter meaning spelled out."""
    x = x.transpose(1, 2).reshape(-1, x.size(1))
    window = torch.hann_window(win_length, device=x.device)      # (device-agnostic: bench.py --eager-gpu-baseline runs this port on the GPU)
    z = torch.stft(x, n_fft, hop_length, win_length, window, return_complex=True)
    y = torch.clamp(z.real ** 2 + z.imag ** 2, min=1e-7).transpose(2, 1)
    return torch.sqrt(y)


def multi_stft_loss(x, y, stft_params, causal=False, causal_size=0):
    """CustomFeatureLoss('stft') -> MultiSizeSTFTLoss -> STFTLoss (loss.py:63-114), including the
    positional (fft, hop, win) -> (fft, win, hop) swap of loss.py:99-102: the yaml's hop_sizes end up
    as torch.stft's win_length and the yaml's win_sizes as its hop_length."""
    if causal:
        if causal_size > 0:
            x, y = x[:, causal_size:], y[:, :-causal_size]
        elif causal_size < 0:
            x, y = x[:, :causal_size], y[:, -causal_size:]
    logratio = stft_params.get("logratio", 0.0)
    losses = []
    for fft, win_y, hop_y in zip(stft_params["fft_sizes"], stft_params["win_sizes"], stft_params["hop_sizes"]):
        xm = stft_mag(x, fft, hop_length=win_y, win_length=hop_y)
        ym = stft_mag(y, fft, hop_length=win_y, win_length=hop_y)
        mag = F.l1_loss(xm, ym)
        lmag = F.l1_loss(xm.log(), ym.log())
        losses.append((1 - logratio) * mag + logratio * lmag)
    return sum(losses) / len(losses)


# ------------------------------------------------------------------------------------------------
def build_models(conf, spkr_size):
    """crank/bin/train.py:56-131."""
    m = {"G": VQVAE2(conf, spkr_size=spkr_size)}
    if conf["use_spkradv_training"]:
        m["SPKRADV"] = SpeakerAdversarialNetwork(conf, spkr_size)
    if conf["use_spkr_classifier"]:
        m["C"] = pwg.ParallelWaveGANDiscriminator(
            in_channels=conf["input_size"], out_channels=spkr_size,
            kernel_size=conf["spkr_classifier_kernel_size"], layers=conf["n_spkr_classifier_layers"],
            conv_channels=64, dilation_factor=1, nonlinear_activation="LeakyReLU",
            nonlinear_activation_params={"negative_slope": 0.2}, bias=True, use_weight_norm=True)
    if conf["trainer_type"] in ["lsgan", "cyclegan", "stargan"]:
        cin = conf["input_size"] + (1 if conf["use_D_uv"] else 0)
        if conf["use_D_spkrcode"]:
            cin += conf["spkr_embedding_size"] if conf["use_spkr_embedding"] else spkr_size
        cout = 1 + (spkr_size if conf["acgan_flag"] else 0)
        m["D"] = pwg.ResidualParallelWaveGANDiscriminator(
            in_channels=cin, out_channels=cout, kernel_size=conf["discriminator_kernel_size"],
            layers=conf["n_discriminator_layers"] * conf["n_discriminator_stacks"],
            stacks=conf["n_discriminator_stacks"], dropout=conf["discriminator_dropout"])
    return m


def build_optimizers(conf, model):
    """crank/net/trainer/utils.py:40-58: adam -> torch.optim.Adam; radam / lamb -> the restated third-party optimizers of
    oracle/optim_port.py (torch_optimizer / pytorch_lamb are absent here)."""
    from . import optim_port

    def one(kind, params, lr):
        if kind == "adam":
            return torch.optim.Adam(params, lr=lr)
        if kind == "radam":
            return optim_port.RAdam(params, lr=lr)
        if kind == "lamb":
            return optim_port.Lamb(params, lr=lr)
        raise ValueError("Invalid optimizer type")

    return {k: one(conf["optim"][k].get("type", "adam"), model[k].parameters(), conf["optim"][k]["lr"])
            for k in ["G", "D", "C", "SPKRADV"] if k in model}


class OracleTrainer:
    """One `train(batch, phase)` of the four reference trainers, as straight-line code."""

    def __init__(self, trainer_type, model, optimizer, conf, steps=0):
        assert trainer_type in ("vqvae", "lsgan", "cyclegan", "stargan")
        self.kind, self.model, self.optimizer, self.conf = trainer_type, model, optimizer, conf
        self.steps = steps
        self.ce = nn.CrossEntropyLoss(ignore_index=-100)
        self.cycle_flag = bool(conf["use_cyclic_training"] and steps > conf["n_steps_cycle_start"])
        self.gan_flag = trainer_type != "vqvae" and steps > conf["n_steps_gan_start"]

    # -- conditioning (basetrainer.py:253-309) --
    def _enc_h(self, b, cv=False):
        if self.conf["encoder_f0"]:
            return torch.cat([b["cv_lcf0"] if cv else b["lcf0"], b["uv"]], axis=-1)
        return None

    def _dec_h(self, b, cv=False):
        h = (b["cv_h"] if cv else b["org_h"]).clone()
        onehot = b["cv_h_onehot"] if cv else b["org_h_onehot"]
        h[:, :] = h[:, 0:1]
        f0 = torch.cat([b["cv_lcf0"] if cv else b["lcf0"], b["uv"]], axis=-1) if self.conf["decoder_f0"] else None
        if not self.conf["use_spkr_embedding"]:
            return (torch.cat([f0, onehot], dim=-1) if f0 is not None else onehot), None
        return f0, h

    def _step(self, loss, key):
        self.optimizer[key].zero_grad()
        loss[key].backward()
        clip = self.conf["optim"][key]["clip_grad_norm"]
        if clip != 0:
            torch.nn.utils.clip_grad_norm_(self.model[key].parameters(), clip)
        self.optimizer[key].step()

    # -- losses (trainer_vqvae.py:210-328) --
    def _vq_terms(self, loss, o, emask, suffix=""):
        for n in range(self.conf["n_vq_stacks"]):
            enc, emb = o["encoded"][n], o["emb_idx"][n]
            loss[f"G_commit{n}{suffix}"] = F.mse_loss(enc.masked_select(emask), emb.masked_select(emask).detach())
            if not self.conf["ema_flag"]:
                loss[f"G_dict{n}{suffix}"] = F.mse_loss(emb.masked_select(emask), enc.masked_select(emask).detach())

    def _vqvae_loss(self, b, o, loss):
        c, a = self.conf, self.conf["alpha"]
        cs = c["causal_size"]
        loss["G_l1"] = feature_loss("l1", o["decoded"], b["out_feats"], b["decoder_mask"], c["causal"], cs)
        loss["G_mse"] = feature_loss("mse", o["decoded"], b["out_feats"], b["decoder_mask"], c["causal"], cs)
        loss["G_stft"] = multi_stft_loss(o["decoded"], b["out_feats"], c["stft_params"], c["causal"], cs)
        self._vq_terms(loss, o, b["encoder_mask"])
        for k in ["l1", "mse", "stft"]:
            loss["G"] += a[k] * loss[f"G_{k}"]
        for n in range(c["n_vq_stacks"]):
            loss["G"] += a["commit"] * loss[f"G_commit{n}"]
        if not c["ema_flag"]:
            for n in range(c["n_vq_stacks"]):
                loss["G"] += a["dict"] * loss[f"G_dict{n}"]

    def _cyclevqvae_loss(self, b, outs, loss):
        c, a = self.conf, self.conf["alpha"]
        for cyc in range(c["n_cycles"]):
            for io in ["cv", "recon"]:
                lbl = f"{cyc}cyc_{io}"
                o = outs[cyc][io]
                if io == "cv":
                    emask = b["encoder_mask"]
                    fake = self.model["C"](o["decoded"].transpose(1, 2)).transpose(1, 2)
                    loss[f"C_fake_{lbl}"] = self.ce(fake.reshape(-1, fake.size(2)), b["cv_h"].reshape(-1))
                else:
                    emask = b["cycle_encoder_mask"]
                    cs = c["causal_size"] * 2 if c["causal"] else 0
                    loss[f"G_l1_{lbl}"] = feature_loss("l1", o["decoded"], b["in_feats"], b["cycle_decoder_mask"], c["causal"], cs)
                    loss[f"G_mse_{lbl}"] = feature_loss("mse", o["decoded"], b["in_feats"], b["cycle_decoder_mask"], c["causal"], cs)
                    loss[f"G_stft_{lbl}"] = multi_stft_loss(o["decoded"], b["in_feats"], c["stft_params"], c["causal"], cs)
                self._vq_terms(loss, o, emask, suffix=f"_{lbl}")
        for cyc in range(c["n_cycles"]):
            for io in ["cv", "recon"]:
                lbl = f"{cyc}cyc_{io}"
                for n in range(c["n_vq_stacks"]):
                    loss["G"] += a["cycle"] * a["commit"] * loss[f"G_commit{n}_{lbl}"]
                    if not c["ema_flag"]:
                        loss["G"] += a["cycle"] * a["dict"] * loss[f"G_dict{n}_{lbl}"]
                if io == "recon":
                    for k in ["l1", "mse", "stft"]:
                        loss["G"] += a["cycle"] * a[k] * loss[f"G_{k}_{lbl}"]
                else:
                    loss["G"] += a["cycle"] * a["ce"] * loss[f"C_fake_{lbl}"]

    def _spkradv_loss(self, b, o, loss, label="org"):
        er = self.model["G"].encoder_receptive_size if self.conf["causal"] else 0
        enc = [e[:, er:] for e in o["encoded_unmod"]] if er else o["encoded_unmod"]
        cls = self.model["SPKRADV"].forward(enc)
        loss[f"G_spkradv_{label}"] = self.ce(cls.reshape(-1, cls.size(2)), b["org_h"][:, er:].reshape(-1))
        w = self.conf["alpha"]["ce"] * (self.conf["alpha"]["cycle"] if label == "recon" else 1)
        loss["G"] += w * loss[f"G_spkradv_{label}"]

    # -- D plumbing (trainer_lsgan.py:146-206) --
    def _D_inputs(self, b, feats, label="org"):
        parts = [feats]
        if self.conf["use_D_uv"]:
            parts.append(b["uv"])
        if self.conf["use_D_spkrcode"]:
            if not self.conf["use_spkr_embedding"]:
                parts.append(b[f"{label}_h_onehot"])
            else:
                h = b[f"{label}_h"].clone()
                h[:, :] = h[:, 0:1]
                parts.append(self.model["G"].spkr_embedding(h).detach())
        return torch.cat(parts, axis=-1).float()

    def _D(self, x):
        return self.model["D"](x.transpose(1, 2)).transpose(1, 2)

    def _acgan(self, cls, h, loss, label="adv", model="G"):
        loss[f"D_acgan_{label}"] = self.ce(cls.reshape(-1, cls.size(2)), h.reshape(-1))
        if not (self.conf["use_real_only_acgan"] and label == "fake"):
            loss[model] += self.conf["alpha"]["acgan"] * loss[f"D_acgan_{label}"]

    def _disc_loss(self, sample, h, mask, loss, label, updates=None):
        if self.conf["acgan_flag"]:
            sample, cls = torch.split(sample, [1, len_spk(sample) - 1], dim=2)
            self._acgan(cls, h, loss, label=label, model="D")
        s = sample.masked_select(mask)
        tgt = torch.ones_like(s) if label == "real" else torch.zeros_like(s)
        loss[f"D_{label}"] = F.mse_loss(s, tgt)
        if updates is None or label in updates:
            loss["D"] += self.conf["alpha"][label] * loss[f"D_{label}"]

    def _adv_loss(self, b, decoded, h, mask, loss):
        fake = self._D(self._D_inputs(b, decoded, label="cv"))
        if self.conf["acgan_flag"]:
            fake, cls = torch.split(fake, [1, len_spk(fake) - 1], dim=2)
            self._acgan(cls, h, loss)
        fake = fake.masked_select(mask)
        loss["D_adv"] = F.mse_loss(fake, torch.ones_like(fake))
        loss["G"] += self.conf["alpha"]["adv"] * loss["D_adv"]

    def _cycle(self, b):
        dec_h, spk = self._dec_h(b)
        dec_h_cv, spk_cv = self._dec_h(b, cv=True)
        return self.model["G"].cycle_forward(b["in_feats"], self._enc_h(b), dec_h, self._enc_h(b, cv=True), dec_h_cv, spk, spk_cv)

    # -- generator-side sub-steps --
    def _forward_vqvae(self, b, loss, train):
        dec_h, spk = self._dec_h(b)
        o = self.model["G"].forward(b["in_feats"], self._enc_h(b), dec_h, spkrvec=spk)
        self._vqvae_loss(b, o, loss)
        if self.conf["use_spkradv_training"]:
            self._spkradv_loss(b, o, loss, "org")
        loss["objective"] += loss["G"]
        if train:
            self._step(loss, "G")

    def _forward_cycle(self, b, loss, train):
        outs = self._cycle(b)
        if self.conf["use_vqvae_loss"]:
            self._vqvae_loss(b, outs[0]["org"], loss)
        self._cyclevqvae_loss(b, outs, loss)
        if self.conf["use_spkradv_training"]:
            for label in ["cv", "recon"]:
                self._spkradv_loss(b, outs[0][label], loss, label)
        loss["objective"] += loss["G"]
        if train:
            self._step(loss, "G")

    def _update_G(self, b, loss, train):
        c = self.conf
        if self.kind == "lsgan":
            dec_h, spk = self._dec_h(b)
            o = self.model["G"].forward(b["in_feats"], self._enc_h(b), dec_h, spk)
            self._vqvae_loss(b, o, loss)
            if c["use_spkradv_training"]:
                self._spkradv_loss(b, o, loss)
            h = b["org_h"]
            if c["cvadv_flag"]:
                dec_h, spk = self._dec_h(b, cv=True)
                h = b["cv_h"]
            adv = self.model["G"].forward(b["in_feats"], self._enc_h(b), dec_h, spkrvec=spk,
                                          use_ema=not c["encoder_detach"], encoder_detach=c["encoder_detach"])
            self._adv_loss(b, adv["decoded"], h, b["decoder_mask"], loss)
        elif self.kind == "cyclegan":
            outs = self._cycle(b)
            self._vqvae_loss(b, outs[0]["org"], loss)
            self._cyclevqvae_loss(b, outs, loss)
            if c["use_spkradv_training"]:
                self._spkradv_loss(b, outs[0]["org"], loss)
            for cyc in range(c["n_cycles"]):
                for io in ["org", "cv"]:
                    lbl = f"{cyc}cyc_{io}"
                    D_out = self._D(self._D_inputs(b, outs[cyc][io]["decoded"], label="cv"))
                    if c["acgan_flag"]:      # trainer_cyclegan.py:107-118: only this branch masks the adversarial term
                        D_out, cls = torch.split(D_out, [1, len_spk(D_out) - 1], dim=2)
                        D_out = D_out.masked_select(b["decoder_mask"])
                        loss[f"D_acgan_adv_{lbl}"] = self.ce(cls.reshape(-1, cls.size(2)), b[f"{io}_h"].reshape(-1))
                        loss["G"] += c["alpha"]["acgan"] * loss[f"D_acgan_adv_{lbl}"]
                    loss[f"D_adv_{lbl}"] = F.mse_loss(D_out, torch.ones_like(D_out))
                    loss["G"] += c["alpha"]["adv"] * loss[f"D_adv_{lbl}"]
        else:  # stargan
            outs = self._cycle(b)
            if c["use_vqvae_loss"]:
                self._vqvae_loss(b, outs[0]["org"], loss)
            self._cyclevqvae_loss(b, outs, loss)
            if c["use_spkradv_training"]:
                for label in ["cv", "recon"]:
                    self._spkradv_loss(b, outs[0][label], loss, label)
            self._adv_loss(b, outs[0]["cv"]["decoded"], b["cv_h"], b["decoder_mask"], loss)
        if train:
            self._step(loss, "G")

    def _update_D(self, b, loss, train):
        c = self.conf
        if self.kind == "lsgan":
            h = b["org_h"]
            if c["cvadv_flag"]:
                dec_h, spk = self._dec_h(b, cv=True)
                h = b["cv_h"]
            else:
                dec_h, spk = self._dec_h(b)
            o = self.model["G"].forward(b["in_feats"], self._enc_h(b), dec_h, spk)
            real = self._D(self._D_inputs(b, b["in_feats"], "org"))
            self._disc_loss(real, b["org_h"], b["decoder_mask"], loss, "real")
            fake = self._D(self._D_inputs(b, o["decoded"].detach(), "cv"))
            self._disc_loss(fake, h, b["decoder_mask"], loss, "fake")
        elif self.kind == "cyclegan":
            outs = self._cycle(b)
            for cyc in range(c["n_cycles"]):
                lbl = f"{cyc}cyc"
                sample = {
                    "real": self._D(self._D_inputs(b, b["in_feats"], "org")),
                    "org_fake": self._D(self._D_inputs(b, outs[0]["org"]["decoded"].detach(), "org")),
                    "cv_fake": self._D(self._D_inputs(b, outs[0]["cv"]["decoded"].detach(), "cv")),
                }
                if c["acgan_flag"]:          # trainer_cyclegan.py:143-159
                    for k in list(sample.keys()):
                        h = b["org_h"] if k in ["real", "org_fake"] else b["cv_h"]
                        sample[k], cls = torch.split(sample[k], [1, len_spk(sample[k]) - 1], dim=2)
                        loss[f"D_ce_{k}_{lbl}"] = self.ce(cls.reshape(-1, cls.size(2)), h.reshape(-1))
                        if not (c["use_real_only_acgan"] and k == "org_fake"):
                            loss["D"] += c["alpha"]["acgan"] * loss[f"D_ce_{k}_{lbl}"]
                rs = sample["real"].masked_select(b["decoder_mask"])
                loss[f"D_real_{lbl}"] = F.mse_loss(rs, torch.ones_like(rs))
                fake_key = random.choice(["org_fake", "cv_fake"])
                mask = b["cycle_decoder_mask"] if fake_key == "org_fake" else b["decoder_mask"]
                fs = sample[fake_key].masked_select(mask)
                loss[f"D_fake_{lbl}"] = F.mse_loss(fs, torch.zeros_like(fs))
                loss["D"] += c["alpha"]["fake"] * loss[f"D_fake_{lbl}"] + c["alpha"]["real"] * loss[f"D_real_{lbl}"]
        else:  # stargan
            dec_h_cv, spk_cv = self._dec_h(b, cv=True)
            updates = random.choice(["real", "fake"]) if c["switch_update"] else ["real", "fake"]
            real = self._D(self._D_inputs(b, b["in_feats"], "org"))
            self._disc_loss(real, b["org_h"], b["decoder_mask"], loss, "real", updates)
            o = self.model["G"].forward(b["in_feats"], self._enc_h(b, cv=True), dec_h_cv, spk_cv)
            fake = self._D(self._D_inputs(b, o["decoded"].detach(), "cv"))
            self._disc_loss(fake, b["cv_h"], b["decoder_mask"], loss, "fake", updates)
        if train:
            self._step(loss, "D")

    # -- the step --
    def train(self, b, phase="train"):
        c = self.conf
        train = phase == "train"
        loss = {"objective": 0.0, "G": 0.0, "D": 0.0, "C": 0.0, "SPKRADV": 0.0}
        if self.kind != "vqvae" and self.gan_flag:
            if c["train_first"] == "G":
                self._update_G(b, loss, train)
                self._update_D(b, loss, train)
            else:
                self._update_D(b, loss, train)
                self._update_G(b, loss, train)
            loss["objective"] = loss["G"] + loss["D"]
        elif self.cycle_flag:
            self._forward_cycle(b, loss, train)
        else:
            self._forward_vqvae(b, loss, train)
        if c["use_spkradv_training"]:            # forward_spkradv (trainer_vqvae.py:163-184)
            dec_h, spk = self._dec_h(b)
            o = self.model["G"].forward(b["in_feats"], self._enc_h(b), dec_h, spkrvec=spk)
            er = self.model["G"].encoder_receptive_size if c["causal"] else 0
            enc = [e[:, er:] for e in o["encoded_unmod"]] if er else o["encoded_unmod"]
            cls = self.model["SPKRADV"].forward(enc, detach=True)
            loss["SPKRADV"] = c["alpha"]["ce"] * self.ce(cls.reshape(-1, cls.size(2)), b["org_h"][:, er:].reshape(-1))
            if train:
                self._step(loss, "SPKRADV")
        if c["use_spkr_classifier"]:             # forward_spkrclassifier (trainer_vqvae.py:186-198)
            real = self.model["C"](b["in_feats"].transpose(1, 2)).transpose(1, 2)
            loss["C_real"] = self.ce(real.reshape(-1, real.size(2)), b["org_h"].reshape(-1))
            loss["C"] += c["alpha"]["ce"] * loss["C_real"]
            if train:
                self._step(loss, "C")
        self.steps += 1
        return {k: (v.item() if isinstance(v, torch.Tensor) else float(v)) for k, v in loss.items()}


def len_spk(t):
    return t.size(2)
