"""Hyper-parameter defaults of the crank recipes, as a Python dict.

Key names and values follow the reference's `egs/vaevc/template/conf/default.yml:1-132`
(the values define every tensor shape on the hot path); recipe overlays
(`vcc2018v1/conf/mlfb_vqvae.yml`, `vcc2020v1/conf/mlfb_vqvae.yml`) only change
`feature.fs`/`shiftms` and `ignore_scaler`.  A user yaml is overlaid with `load_yaml`
exactly like `crank/utils/utils.py:67-84` does with `$CRANK_DEFAULT_YAML`.
"""

import copy
import os


def _optim(lr):
    return {
        "type": "adam",
        "lr": lr,
        "decay_size": 0.5,
        "decay_step_size": 200000,
        "clip_grad_norm": 0.0,
    }


_DEFAULT = {
    "feature": {
        "label": "mlfb", "fs": 22050, "fftl": 1024, "win_length": 1024, "hop_size": 128,
        "window_types": ["hann"], "fmin": 80, "fmax": 7600, "mlfb_dim": 80,
        "n_iteration": 100, "framems": 20, "shiftms": 5.80499, "mcep_dim": 34,
        "mcep_alpha": 0.466,
    },
    # general
    "trainer_type": "vqvae", "input_feat_type": "mlfb", "output_feat_type": "mlfb",
    "use_raw": False, "use_preprocessed_scaler": False, "use_sinc_conv": False,
    "raw_window_type": "hann", "input_size": 80, "output_size": 80,
    "n_steps": 200000, "dev_steps": 2000, "n_steps_save_model": 5000,
    "n_steps_print_loss": 50, "batch_size": 50, "batch_len": 500,
    "cache_dataset": True, "spec_augment": False, "n_spec_augment": 0,
    "use_mcep_0th": False, "ignore_scaler": ["raw", "mcep"],
    "sinc_conv_kernel_sizes": 65, "sinc_conv_channels": 32,
    "sinc_conv_down_sample_kernel_sizes": [4, 4, 4, 2],
    # loss weights
    "alpha": {
        "l1": 2, "mse": 0, "stft": 1, "commit": 0.25, "dict": 0.5, "cycle": 0.1,
        "ce": 1, "adv": 1, "real": 0.5, "fake": 0.5, "acgan": 1,
    },
    "stft_params": {
        "fft_sizes": [64, 128], "win_sizes": [64, 128], "hop_sizes": [16, 32], "logratio": 0,
    },
    "optim": {
        "G": _optim(0.0002), "D": _optim(0.00005), "C": _optim(0.0001),
        "SPKRADV": _optim(0.0001),
    },
    # generator
    "encoder_f0": False, "decoder_f0": True, "encoder_energy": False,
    "decoder_energy": False, "causal": False, "causal_size": 0,
    "use_spkr_embedding": True, "spkr_embedding_size": 32, "ema_flag": True,
    "n_vq_stacks": 2, "n_layers_stacks": [4, 3, 2], "n_layers": [2, 2, 2],
    "kernel_size": [5, 3, 3], "emb_dim": [64, 64, 64], "emb_size": [512, 512, 512],
    "use_spkradv_training": True, "n_spkradv_layers": 3, "spkradv_kernel_size": 3,
    "spkradv_lambda": 0.1, "use_spkr_classifier": True, "n_spkr_classifier_layers": 8,
    "spkr_classifier_kernel_size": 5, "use_cyclic_training": False,
    "n_steps_cycle_start": 50000, "n_cycles": 1,
    # discriminator
    "n_steps_gan_start": 100000, "gan_type": "lsgan", "use_residual_network": True,
    "n_discriminator_layers": 2, "n_discriminator_stacks": 4,
    "discriminator_kernel_size": 5, "discriminator_dropout": 0.25, "train_first": "D",
    "switch_update": False, "cvadv_flag": False, "acgan_flag": False,
    "encoder_detach": False, "use_real_only_acgan": False, "use_D_uv": True,
    "use_D_spkrcode": True, "use_vqvae_loss": True, "n_steps_stop_generator": 0,
}


def default_conf(**overrides):
    """A fresh deep copy of the template defaults, with top-level overrides applied."""
    conf = copy.deepcopy(_DEFAULT)
    _overlay(conf, overrides)
    return conf


def _overlay(base, new):
    for k, v in new.items():
        if isinstance(v, dict) and k in base and isinstance(base[k], dict):
            _overlay(base[k], v)
        else:
            base[k] = v


def load_yaml(ymlf):
    """Overlay `ymlf` onto `$CRANK_DEFAULT_YAML` when set (crank/utils/utils.py:67-84),
    else onto the built-in template defaults."""
    import yaml

    with open(ymlf) as fp:
        yml = yaml.load(fp, Loader=yaml.SafeLoader)
    default_ymlf = os.environ.get("CRANK_DEFAULT_YAML")
    if default_ymlf is None:
        base = default_conf()
    else:
        with open(default_ymlf) as fp:
            base = yaml.load(fp, Loader=yaml.SafeLoader)
    _overlay(base, yml or {})
    return base


def vcc2018_conf(**overrides):
    """egs/vaevc/vcc2018v1/conf/mlfb_vqvae.yml over default.yml (12 speakers, fs 22050)."""
    c = default_conf(ignore_scaler=[])
    _overlay(c, overrides)
    return c


def vcc2020_conf(**overrides):
    """egs/vaevc/vcc2020v1/conf/mlfb_vqvae.yml over default.yml (14 speakers, fs 24000)."""
    c = default_conf(ignore_scaler=[])
    c["feature"]["fs"] = 24000
    c["feature"]["shiftms"] = 5.333333
    _overlay(c, overrides)
    return c
