"""VQVAE2 generator + Quantizer on the sm_100a kernels.

API mirror of crank/net/module/vqvae2.py (VQVAE2 :38-283, Quantizer :286-347): same constructor,
`forward` / `cycle_forward` / `encode` / `decode` / `make_dict` signatures, output dict schema,
attributes (`encoders`, `decoders`, `quantizers`, `spkr_embedding`, `*_receptive_size`, `conf`)
and state-dict keys.  Internally everything stays channels-last (B, T, C): the reference's
transposes to (B, C, T) exist only to feed `F.conv1d` and are gone.

Reference behaviours reproduced on purpose:
  * `decode()` mutates the list it is given (vqvae2.py:177); `cycle_forward` therefore feeds its
    second decode with already-incremented encoder outputs and returns ONE shared `encoded` list
    for "org" and "cv" (vqvae2.py:118-119).
  * the decoder input concatenates the quantised stacks top-first ([qx1, qx0], vqvae2.py:181-190).
  * the EMA codebook update fires on every forward while `self.training` (nobody calls .eval()),
    using the pre-update codebook for the current call (vqvae2.py:315-330).
"""

import torch
import torch.nn as nn

from .. import _dp
from ... import lib as L
from ... import ops
from ...parallel_wavegan.models import ParallelWaveGANGenerator


class VQVAE2(nn.Module):
    def __init__(self, conf, spkr_size=0, scaler=None):
        super().__init__()
        self.conf = conf
        self.spkr_size = spkr_size
        self.encoder_receptive_size = 0
        self.decoder_receptive_size = 0
        self._construct_net()
        if conf["use_spkr_embedding"]:
            self.spkr_embedding = nn.Embedding(spkr_size, conf["spkr_embedding_size"])
        if conf["use_raw"]:
            from .mlfb import LogMelFilterBankLayer

            feat = conf["feature"]
            mlfb_scaler = scaler["mlfb"] if conf["use_preprocessed_scaler"] else None
            self.preprocess_layer = LogMelFilterBankLayer(
                fs=feat["fs"], hop_size=feat["hop_size"], fft_size=feat["fftl"],
                win_length=feat["win_length"], window=conf["raw_window_type"], center=False,
                n_mels=feat["mlfb_dim"], fmin=feat["fmin"], fmax=feat["fmax"], scaler=mlfb_scaler,
            )
        elif conf["use_sinc_conv"]:
            raise NotImplementedError(
                "use_sinc_conv is out of scope (off in every recipe; the reference wiring is broken, "
                "crank/net/module/vqvae2.py:76-82)"
            )

    def _pre(self, x):
        if self.conf["use_raw"]:
            return self.preprocess_layer(x)
        return x

    def forward(self, x, enc_h, dec_h, spkrvec=None, use_ema=True, encoder_detach=False, final_decoder=True):
        """`final_decoder=False` (not in the reference signature; default = reference behaviour) skips the bottom
        decoder stack, whose output no quantiser consumes: used by the speaker-adversarial update, which only reads
        the encoder outputs of this pass but must keep its EMA codebook updates (trainer_vqvae.py:165-180).
        "decoded" is then None."""
        self._dp_wait()
        x = self._pre(x)
        dec_h = self._get_dec_h(dec_h, spkrvec)
        enc = self.encode(x, enc_h=enc_h)
        enc_unmod = list(enc)
        enc, dec, emb_idxs, _, qidxs = self.decode(enc, dec_h, use_ema=use_ema, detach=encoder_detach,
                                                   final_decoder=final_decoder)
        return self.make_dict(enc, dec, emb_idxs, qidxs, enc_unmod)

    def _dp_wait(self):
        # data parallel: a side-stream gradient all-reduce + Adam step (or EMA codebook update) of this generator may
        # still be in flight: the compute stream waits for it before the parameters / codebooks are read
        if _dp.active():
            _dp.wait_for(list(self.parameters()) + list(self.buffers()))

    def cycle_forward(self, x, org_enc_h, org_dec_h, cv_enc_h, cv_dec_h, org_spkrvec, cv_spkrvec):
        self._dp_wait()
        x = self._pre(x)
        org_dec_h = self._get_dec_h(org_dec_h, org_spkrvec)
        cv_dec_h = self._get_dec_h(cv_dec_h, cv_spkrvec)
        outputs = []
        for _ in range(self.conf["n_cycles"]):
            enc = self.encode(x, enc_h=org_enc_h)
            org_unmod, cv_unmod = list(enc), list(enc)
            # both decodes share (and mutate) `enc`
            org_enc, org_dec, org_emb, _, org_q = self.decode(enc, org_dec_h)
            cv_enc, cv_dec, cv_emb, _, cv_q = self.decode(enc, cv_dec_h)
            enc = self.encode(cv_dec, enc_h=cv_enc_h)
            rec_unmod = list(enc)
            rec_enc, rec_dec, rec_emb, _, rec_q = self.decode(enc, org_dec_h)
            outputs.append({
                "org": self.make_dict(org_enc, org_dec, org_emb, org_q, org_unmod),
                "cv": self.make_dict(cv_enc, cv_dec, cv_emb, cv_q, cv_unmod),
                "recon": self.make_dict(rec_enc, rec_dec, rec_emb, rec_q, rec_unmod),
            })
            x = rec_dec.detach()
        return outputs

    def _get_dec_h(self, dec_h, spkrvec):
        if spkrvec is not None:
            emb = self.spkr_embedding(spkrvec)
            dec_h = emb if dec_h is None else torch.cat([dec_h, emb], dim=-1)
        return dec_h

    def encode(self, x, enc_h=None):
        encoded = []
        for n in range(self.conf["n_vq_stacks"]):
            enc = self.encoders[n].forward_cl(x if n == 0 else enc, enc_h if n == 0 else None)
            encoded.append(enc)
        return encoded

    def decode(self, enc, dec_h, use_ema=True, detach=False, final_decoder=True):
        dec = None
        emb_idxs, emb_idx_qxs, qidxs = [], [], []
        for n in reversed(range(self.conf["n_vq_stacks"])):
            if dec is not None:
                enc[n] = enc[n] + dec
            emb_idx, qx, qidx = self.quantizers[n].forward_cl(enc[n], use_ema=use_ema)
            if detach:
                qx = qx.detach()
            emb_idxs.append(emb_idx)
            emb_idx_qxs.append(qx)
            qidxs.append(qidx)
            if n != 0:
                dec = self.decoders[n].forward_cl(qx, None)
            elif final_decoder:
                dec = self.decoders[n].forward_cl(torch.cat(emb_idx_qxs, dim=-1), dec_h)
            else:
                dec = None
        return enc, dec, emb_idxs, emb_idx_qxs, qidxs

    def remove_weight_norm(self):
        for n in range(self.conf["n_vq_stacks"]):
            self.encoders[n].remove_weight_norm()
            self.decoders[n].remove_weight_norm()

    def make_dict(self, enc, dec, emb_idxs, qidxs, enc_unmod):
        # index 0 = bottom stack; everything (B, T, D)
        return {
            "encoded": list(enc),
            "encoded_unmod": list(enc_unmod) if enc_unmod is not None else None,
            "decoded": dec,
            "emb_idx": emb_idxs[::-1],
            "qidx": qidxs[::-1],
        }

    def _construct_net(self):
        c = self.conf
        self.encoders = nn.ModuleList()
        self.decoders = nn.ModuleList()
        self.quantizers = nn.ModuleList()
        for n in range(c["n_vq_stacks"]):
            if n == 0:
                enc_io = (c["input_size"], c["emb_dim"][0], 2 if c["encoder_f0"] else 0)
                dec_aux = 2 if c["decoder_f0"] else 0
                dec_aux += c["spkr_embedding_size"] if c["use_spkr_embedding"] else self.spkr_size
                dec_io = (sum(c["emb_dim"][i] for i in range(c["n_vq_stacks"])), c["output_size"], dec_aux)
            else:
                enc_io = (c["emb_dim"][n - 1], c["emb_dim"][n], 0)
                dec_io = (c["emb_dim"][n], c["emb_dim"][n - 1], 0)
            common = dict(
                kernel_size=c["kernel_size"][n],
                layers=c["n_layers"][n] * c["n_layers_stacks"][n],
                stacks=c["n_layers_stacks"][n],
                residual_channels=64, gate_channels=128, skip_channels=64,
                aux_context_window=0, dropout=0.0, bias=True, use_weight_norm=True,
                use_causal_conv=c["causal"], upsample_conditional_features=False,
            )
            self.encoders.append(ParallelWaveGANGenerator(
                in_channels=enc_io[0], out_channels=enc_io[1], aux_channels=enc_io[2], **common))
            self.decoders.append(ParallelWaveGANGenerator(
                in_channels=dec_io[0], out_channels=dec_io[1], aux_channels=dec_io[2], **common))
            self.encoder_receptive_size += self.encoders[-1].receptive_field_size
            self.decoder_receptive_size += self.decoders[-1].receptive_field_size
            self.quantizers.append(
                Quantizer(c["emb_dim"][n], c["emb_size"][n], ema_flag=c["ema_flag"], bdt_flag=True))


class Quantizer(nn.Module):
    def __init__(self, emb_dim, emb_size, decay=0.99, eps=1e-5, ema_flag=False, bdt_flag=False):
        super().__init__()
        self.emb_dim = emb_dim
        self.emb_size = emb_size
        self.ema_flag = ema_flag
        self.bdt_flag = bdt_flag
        self.embedding = nn.Embedding(emb_size, emb_dim)
        self.embedding.weight.data.uniform_(-1.0 / emb_size, 1.0 / emb_size)
        if ema_flag:
            self.decay = decay
            self.eps = eps
            self.register_buffer("ema_size", torch.zeros(emb_size))
            self.register_buffer("ema_w", torch.randn(emb_dim, emb_size))
        self._op_blob, self._op_key = None, None

    def forward(self, x, use_ema=True):
        """Reference layout: x (B,D,T) when bdt_flag else (B,T,D) ->
        (embed_idx (B,T,D), embed_idx_qx same layout as x, idx (B,T))."""
        if self.bdt_flag:
            x = x.transpose(1, 2)
        e, qx, idx = self.forward_cl(x, use_ema=use_ema)
        if self.bdt_flag:
            qx = qx.transpose(1, 2)
        return e, qx, idx

    def _operand_blob(self):
        """Cached tensor-core operand of the codebook for crk_vq_argmin_fast.  The fused EMA kernel rewrites it together
        with the codebook; anything else that touches the weight (optimizer step, load_state_dict, init) bumps the
        tensor version and triggers a re-pack here."""
        W = self.embedding.weight
        if not (W.is_cuda and ops.vq_fast_ok(self.emb_size, self.emb_dim)):
            return None
        key = (W.data_ptr(), W._version)
        if self._op_key != key or self._op_blob is None or self._op_blob.device != W.device:
            reuse = self._op_blob if (self._op_blob is not None and self._op_blob.device == W.device) else None
            self._op_blob = ops.vq_pack_operand(W, reuse)
            self._op_key = key
        return self._op_blob

    def forward_cl(self, x, use_ema=True):
        W = self.embedding.weight
        if _dp.active():
            _dp.wait_for([W] + ([self.ema_size, self.ema_w] if self.ema_flag else []))
        # a gradient-trained codebook changes through raw-pointer optimizer kernels: pack per call there
        blob = self._operand_blob() if self.ema_flag else None
        e, qx, idx = ops.VQFn.apply(x, W if not self.ema_flag else W.detach(), blob)
        if self.training and self.ema_flag and use_ema:
            with torch.no_grad():
                # data parallel: the [counts | sums] all-reduce and the EMA kernel that follows it go to the
                # communication stream; this quantiser's NEXT call (the only reader of the new codebook) waits for them
                keys = (W, self.ema_size, self.ema_w)
                ops.vq_ema_update(x.detach(), idx, self.ema_size, self.ema_w, W.data, self.decay,
                                  self.eps, reduce_fn=_dp.stats_reducer(),
                                  runner=(lambda fn, stats: _dp.run_async(keys, fn, tensors=(stats,))) if _dp.active() else None,
                                  opblob=blob)
        return e, qx, idx

    def vq(self, x):
        """(idx (B,T), one-hot (B,T,K)) like the reference helper (vqvae2.py:338-347)."""
        _, _, idx = ops.VQFn.apply(x, self.embedding.weight.detach())
        return idx, torch.nn.functional.one_hot(idx, self.emb_size)
