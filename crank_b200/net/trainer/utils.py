"""Factories with the reference's names and return schemas (crank/net/trainer/utils.py:22-74,
crank/bin/train.py:56-131): criterion dict keys, one optimizer + StepLR per sub-model, and the
{"G","SPKRADV","C","D"} model dict."""

from torch.optim.lr_scheduler import StepLR

from ..module.loss import CrossEntropyLoss, CustomFeatureLoss, MaskedMSELoss
from ..module.spkradv import SpeakerAdversarialNetwork
from ..module.vqvae2 import VQVAE2
from ...parallel_wavegan.models import (
    ParallelWaveGANDiscriminator,
    ResidualParallelWaveGANDiscriminator,
)
from .optim import FusedAdam, FusedLamb, FusedRAdam


def get_criterion(conf, device="cuda"):
    return {
        "mse": MaskedMSELoss("mse"),
        "l1": MaskedMSELoss("l1"),
        "ce": CrossEntropyLoss(ignore_index=-100),
        "fmse": CustomFeatureLoss(loss_type="mse", causal=conf["causal"]),
        "fl1": CustomFeatureLoss(loss_type="l1", causal=conf["causal"]),
        "fstft": CustomFeatureLoss(loss_type="stft", stft_params=conf["stft_params"],
                                   causal=conf["causal"], device=device),
    }


def get_optimizer(conf, model):
    optimizer = {}
    for m in ["G", "D", "C", "SPKRADV"]:
        if m in model:
            kind = conf["optim"][m]["type"]
            table = {"adam": FusedAdam, "radam": FusedRAdam, "lamb": FusedLamb}     # crank/net/trainer/utils.py:41-49
            if kind not in table:
                raise ValueError("Invalid optimizer type")
            optimizer[m] = table[kind](model[m].parameters(), lr=conf["optim"][m]["lr"])
    return optimizer


def get_scheduler(conf, optimizer):
    return {
        m: StepLR(optimizer[m], step_size=conf["optim"][m]["decay_step_size"],
                  gamma=conf["optim"][m]["decay_size"])
        for m in ["G", "D", "C", "SPKRADV"] if m in optimizer
    }


def get_model(conf, spkr_size=0, device="cuda", scaler=None):
    """crank/bin/train.py:56-131."""
    models = {"G": VQVAE2(conf, spkr_size=spkr_size, scaler=scaler).to(device)}
    if conf["use_spkradv_training"]:
        models["SPKRADV"] = SpeakerAdversarialNetwork(conf, spkr_size).to(device)
    if conf["use_spkr_classifier"]:
        models["C"] = ParallelWaveGANDiscriminator(
            in_channels=conf["input_size"], out_channels=spkr_size,
            kernel_size=conf["spkr_classifier_kernel_size"],
            layers=conf["n_spkr_classifier_layers"], conv_channels=64, dilation_factor=1,
            nonlinear_activation="LeakyReLU", nonlinear_activation_params={"negative_slope": 0.2},
            bias=True, use_weight_norm=True,
        ).to(device)
    if conf["trainer_type"] in ["lsgan", "cyclegan", "stargan"]:
        in_ch = conf["input_size"]
        if conf["use_D_uv"]:
            in_ch += 1
        if conf["use_D_spkrcode"]:
            in_ch += conf["spkr_embedding_size"] if conf["use_spkr_embedding"] else spkr_size
        out_ch = 1 + (spkr_size if conf["acgan_flag"] else 0)
        if not conf["use_residual_network"]:
            # the reference's non-residual branch cannot be constructed (train.py:121 multiplies an
            # int by a list -> TypeError); keep failing loudly rather than inventing a variant
            raise TypeError("use_residual_network: false is broken in the reference (crank/bin/train.py:121)")
        models["D"] = ResidualParallelWaveGANDiscriminator(
            in_channels=in_ch, out_channels=out_ch, kernel_size=conf["discriminator_kernel_size"],
            layers=conf["n_discriminator_layers"] * conf["n_discriminator_stacks"],
            stacks=conf["n_discriminator_stacks"], dropout=conf["discriminator_dropout"],
        ).to(device)
    return models
