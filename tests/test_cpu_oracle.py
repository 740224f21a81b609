"""CPU tests (run everywhere, no GPU): the oracle against the reference's golden vectors, the oracle
port against the real reference when it is present, host logic, and the C-ABI's symbol table."""
import ctypes
import os
import random
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD_DIR = os.path.join(ROOT, "tests", "golden")


def test_mel_oracle_reproduces_reference_h5_fixture():
    from oracle import mel

    fx = np.load(os.path.join(GOLD_DIR, "ref_fixture_mlfb.npz"))
    raw = fx["raw_i16"].astype(np.float64) / 32768.0
    got = mel.logmelfilterbank(raw, 22050, fft_size=1024, hop_size=128, win_length=1024,
                               window=mel.hann(1024, periodic=False), num_mels=80, fmin=80, fmax=7600)
    assert got.shape == fx["mlfb"].shape == (1057, 80)
    assert np.abs(got - fx["mlfb"]).max() < 1e-6      # measured 4.2e-8
    # the periodic window is NOT what the offline extractor used (crank/feature/feature.py:174)
    per = mel.logmelfilterbank(raw, 22050, fft_size=1024, hop_size=128, win_length=1024, window="hann",
                               num_mels=80, fmin=80, fmax=7600)
    assert np.abs(per - fx["mlfb"]).max() > 1e-3


def test_product_mel_basis_equals_oracle_mel_basis():
    from crank_b200.net.module.mlfb import mel_basis
    from oracle import mel

    for sr in (22050, 24000, 8000):
        a = mel_basis(sr, 1024, 80, 80, min(7600, sr / 2))
        b = mel.mel_basis(sr, 1024, 80, 80, min(7600, sr / 2))
        assert a.shape == (80, 513) and np.abs(a - b).max() < 1e-9


def test_restated_parallel_wavegan_structure():
    """parameter counts / receptive field / key names the survey measured on the reference's graph."""
    from crank_b200.conf import vcc2020_conf
    from oracle import crank_port as cp

    conf = vcc2020_conf(trainer_type="lsgan")
    m = cp.build_models(conf, 14)
    n = {k: sum(p.numel() for p in v.parameters()) for k, v in m.items()}
    assert n == {"G": 1352672, "SPKRADV": 39836, "C": 153884, "D": 408962}
    assert m["G"].encoder_receptive_size == 68 and m["G"].decoder_receptive_size == 68
    keys = list(m["G"].state_dict().keys())
    assert "encoders.0.conv_layers.3.conv.weight_g" in keys and "quantizers.0.ema_w" in keys
    assert "spkr_embedding.weight" in keys


@pytest.mark.parametrize("kind", ["vqvae", "lsgan", "cyclegan", "stargan"])
def test_oracle_port_reproduces_reference_golden(kind):
    """oracle/crank_port.py must reproduce what the REAL reference trainers produced (tests/golden)."""
    from crank_b200.synthetic import clone_batch, make_batch
    from oracle import crank_port as cp
    from tests.golden.make_golden import GOLD_B, GOLD_SPKRS, GOLD_T, checksum, golden_conf

    gold = np.load(os.path.join(GOLD_DIR, "ref_train_golden.npz"))
    conf = golden_conf(kind)
    random.seed(1234)
    np.random.seed(1234)
    torch.manual_seed(1234)
    model = {"G": cp.VQVAE2(conf, spkr_size=GOLD_SPKRS), "SPKRADV": cp.SpeakerAdversarialNetwork(conf, GOLD_SPKRS)}
    rest = cp.build_models(conf, GOLD_SPKRS)
    model["C"] = rest["C"]
    if "D" in rest:
        model["D"] = rest["D"]
    init = np.stack([checksum(model[k]) for k in sorted(model)])
    if not np.allclose(init, gold[f"{kind}/init_checksum"], rtol=1e-12):
        pytest.skip("torch CPU RNG stream differs from the container that produced the golden")
    O = cp.OracleTrainer(kind, model, cp.build_optimizers(conf, model), conf)
    batch = make_batch(GOLD_B, GOLD_T, GOLD_SPKRS, seed=0, ragged=True)
    for it in range(2):
        random.seed(100 + it)
        vals = O.train(clone_batch(batch), "train")
        keys = list(gold[f"{kind}/step{it}_keys"])
        assert sorted(vals) == keys
        got = np.array([vals[k] for k in keys])
        np.testing.assert_allclose(got, gold[f"{kind}/step{it}_vals"], rtol=1e-6, atol=1e-9)
    fin = np.stack([checksum(model[k]) for k in sorted(model)])
    np.testing.assert_allclose(fin, gold[f"{kind}/final_checksum"], rtol=1e-9)


def _oracle_vs_live_reference(kind, overrides, T=80):
    from oracle import refshim

    if not refshim.available():
        pytest.skip("/root/reference is not present on this box")
    from crank_b200.conf import vcc2020_conf
    from crank_b200.synthetic import clone_batch, make_batch, spkr_dict
    from oracle import crank_port as cp

    refshim.install()
    vq = refshim.ref("crank.net.module.vqvae2")
    spk = refshim.ref("crank.net.module.spkradv")
    tr = refshim.ref("crank.net.trainer")
    tu = refshim.ref("crank.net.trainer.utils")
    S = 5
    conf = vcc2020_conf(trainer_type=kind, n_steps_gan_start=-1, discriminator_dropout=0.0, **overrides)
    torch.manual_seed(7)
    ref_m = {"G": vq.VQVAE2(conf, spkr_size=S), "SPKRADV": spk.SpeakerAdversarialNetwork(conf, S)}
    extra = cp.build_models(conf, S)
    ref_m["C"] = extra["C"]
    if "D" in extra:
        ref_m["D"] = extra["D"]
    om = cp.build_models(conf, S)
    for k in om:
        om[k].load_state_dict(ref_m[k].state_dict())
    opt = tu.get_optimizer(conf, ref_m)
    W = refshim.NullWriter()
    T_ = tr.TrainerWrapper(kind, model=ref_m, optimizer=opt, criterion=tu.get_criterion(conf, device="cpu"),
                           dataloader={"spkrs": spkr_dict(S)}, writer={"train": W, "dev": W}, expdir="/tmp/exp",
                           conf=conf, feat_conf=conf["feature"], scheduler=tu.get_scheduler(conf, opt), scaler=None,
                           resume=0, device="cpu", n_jobs=1)
    T_.tqdm.close()
    O = cp.OracleTrainer(kind, om, cp.build_optimizers(conf, om), conf)
    b = make_batch(2, T, S, seed=3, ragged=True)
    random.seed(1)                      # cyclegan / stargan draw host-side choices (trainer_cyclegan.py:166)
    r = T_.train(clone_batch(b), "train")
    random.seed(1)
    o = O.train(clone_batch(b), "train")
    assert set(r) == set(o)
    for k in r:
        assert r[k] == o[k], (k, r[k], o[k])
    for k in om:
        for (n1, p1), (n2, p2) in zip(ref_m[k].state_dict().items(), om[k].state_dict().items()):
            assert n1 == n2 and torch.equal(p1, p2), (k, n1)


@pytest.mark.parametrize("kind", ["vqvae", "lsgan", "cyclegan", "stargan"])
def test_oracle_port_is_bit_identical_to_live_reference(kind):
    _oracle_vs_live_reference(kind, {})


@pytest.mark.parametrize("kind,overrides", [
    ("lsgan", dict(causal=True, causal_size=4)),       # left padding, shifted losses, receptive-field crop (Appendix A)
    ("cyclegan", dict(n_vq_stacks=3)),
    ("vqvae", dict(n_vq_stacks=1)),
    ("lsgan", dict(use_spkr_embedding=False)),         # one-hot speaker code instead of the embedding
    ("cyclegan", dict(encoder_f0=True)),
    ("lsgan", dict(ema_flag=False)),                   # dictionary loss + codebook gradient instead of EMA
    ("cyclegan", dict(acgan_flag=True)),
    ("lsgan", dict(acgan_flag=True)),
    ("stargan", dict(cvadv_flag=True)),
    ("lsgan", dict(encoder_detach=True)),
])
def test_oracle_port_matches_live_reference_on_config_variants(kind, overrides):
    """Every recipe switch that changes the train step's graph: loss dicts and all parameters after one step are
    bit-identical to the unmodified reference trainers (causal needs frames beyond the receptive field)."""
    _oracle_vs_live_reference(kind, overrides, T=160 if overrides.get("causal") else 80)


def test_c_abi_library_exports_every_declared_symbol():
    """include/crank_b200.h <-> libcrank_b200.so <-> crank_b200/lib.py stay in sync (no compute calls)."""
    from crank_b200 import build, lib

    build.build()
    header = open(os.path.join(ROOT, "include", "crank_b200.h")).read()
    declared = set(re.findall(r"\b(crk_[a-z0-9_]+)\s*\(", header))
    declared -= {"crk_conv_desc", "crk_wavenet_cfg", "crk_convstack_cfg"}
    assert len(declared) >= 30
    handle = ctypes.CDLL(lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in crank_b200.h but not exported"
        assert name in lib.SIGNATURES, f"{name} has no ctypes signature in crank_b200/lib.py"
    assert set(lib.SIGNATURES) == declared
    L = lib.lib()
    assert L.crk_version() >= 100
    assert L.crk_strerror(0) == b"ok" and L.crk_strerror(-1) == b"invalid argument"


def test_layout_queries_and_argument_errors_without_gpu():
    from crank_b200 import lib as L

    cfg = L.WavenetCfg(in_ch=80, out_ch=64, aux_ch=0, layers=8, stacks=4, kernel_size=5, causal=0,
                       first_act=0, head_act=1, slope=0.0)
    descs, th, we = L.describe_wavenet(cfg)
    assert len(descs) == 1 + 8 * 3 + 2 and th == 411008     # encoder-0 parameter count
    assert L.lib().crk_wavenet_act_floats(ctypes.byref(cfg), 2, 100) > 0
    bad = L.WavenetCfg(in_ch=80, out_ch=64, aux_ch=0, layers=7, stacks=4, kernel_size=5, causal=0,
                       first_act=0, head_act=1, slope=0.0)
    with pytest.raises(L.CrkError):
        L.describe_wavenet(bad)
    cs = L.ConvstackCfg(in_ch=80, out_ch=14, layers=8, kernel_size=5, conv_ch=64, dilation_factor=1, slope=0.2)
    descs, th, we = L.describe_convstack(cs)
    assert th == 153884 and [d.cin for d in descs] == [80] + [64] * 7
    # NULL pointers are rejected before any CUDA call
    assert L.lib().crk_vq_prepare(None, None, None, 512, 64, None) == -1


def test_product_refuses_cpu_tensors():
    from crank_b200 import lib as L
    from crank_b200.parallel_wavegan.models import ParallelWaveGANGenerator

    net = ParallelWaveGANGenerator(in_channels=80, out_channels=64, kernel_size=3, layers=2, stacks=1,
                                   aux_channels=0, upsample_conditional_features=False)
    with pytest.raises(L.CrkError):
        net(torch.randn(1, 80, 16), None)


def test_state_dict_round_trip_with_reference_key_names():
    from crank_b200.conf import vcc2020_conf
    from crank_b200.net.trainer import get_model
    from oracle import crank_port as cp

    conf = vcc2020_conf(trainer_type="lsgan")
    torch.manual_seed(0)
    om = cp.build_models(conf, 14)
    pm = get_model(conf, 14, device="cpu")
    for k in om:
        res = pm[k].load_state_dict(om[k].state_dict())
        assert not res.missing_keys and not res.unexpected_keys
        back = pm[k].state_dict()
        assert set(back) == set(om[k].state_dict())
        for name, v in om[k].state_dict().items():
            assert torch.equal(back[name], v), (k, name)
    # checkpoint schema of basetrainer.save_model
    with pytest.raises(RuntimeError):
        pm["C"].load_state_dict({"conv_layers.0.weight_g": torch.zeros(3)})


def test_conf_defaults_and_yaml_overlay(tmp_path):
    from crank_b200.conf import default_conf, load_yaml, vcc2020_conf

    c = default_conf()
    assert c["stft_params"]["hop_sizes"] == [16, 32] and c["optim"]["D"]["lr"] == 5e-5
    assert vcc2020_conf()["feature"]["fs"] == 24000
    y = tmp_path / "x.yml"
    y.write_text("trainer_type: lsgan\nalpha:\n  adv: 2\n")
    m = load_yaml(str(y))
    assert m["trainer_type"] == "lsgan" and m["alpha"]["adv"] == 2 and m["alpha"]["l1"] == 2


def test_synthetic_batch_schema():
    from crank_b200.synthetic import make_batch

    b = make_batch(3, 50, 14, ragged=True)
    assert b["in_feats"].shape == (3, 50, 80) and b["org_h"].dtype == torch.int64
    assert b["encoder_mask"].dtype == torch.bool and b["encoder_mask"].shape == (3, 50, 1)
    for i in range(3):
        n = int(b["flen"][i])
        assert (b["org_h"][i, n:] == -100).all() and (b["org_h"][i, :n] >= 0).all()
        assert not b["decoder_mask"][i, n:].any() and b["in_feats"][i, n:].abs().sum() == 0
    assert ((b["cv_h"][:, 0] - b["org_h"][:, 0]) % 14 == 1).all()


def test_conf_defaults_equal_the_reference_default_yaml():
    """crank_b200.conf mirrors egs/vaevc/template/conf/default.yml key by key (every tensor shape on the hot path
    derives from it); checked against the reference checkout when it is present."""
    import yaml

    from crank_b200.conf import default_conf
    from oracle import refshim

    path = os.path.join(refshim.REFERENCE_ROOT, "egs", "vaevc", "template", "conf", "default.yml")
    if not os.path.exists(path):
        pytest.skip("reference checkout not present on this box")

    def flat(d, pre=""):
        out = {}
        for k, v in d.items():
            if isinstance(v, dict):
                out.update(flat(v, pre + k + "."))
            else:
                out[pre + k] = v
        return out

    ref, mine = flat(yaml.safe_load(open(path))), flat(default_conf())
    assert set(ref) == set(mine), (sorted(set(ref) - set(mine)), sorted(set(mine) - set(ref)))
    diff = {k: (ref[k], mine[k]) for k in ref if ref[k] != mine[k]}
    assert not diff, diff


def test_frozen_context_restores_requires_grad_even_on_error():
    from crank_b200.net.trainer.basetrainer import frozen

    m = torch.nn.Linear(3, 2)
    m.bias.requires_grad_(False)                       # already frozen parameters stay frozen afterwards
    with frozen(m):
        assert not m.weight.requires_grad
        y = m(torch.ones(1, 3, requires_grad=True))
    assert m.weight.requires_grad and not m.bias.requires_grad
    x = torch.ones(1, 3, requires_grad=True)
    with frozen(m):
        out = m(x).sum()
    out.backward()
    assert m.weight.grad is None and x.grad is not None
    with pytest.raises(RuntimeError):
        with frozen(m):
            raise RuntimeError("boom")
    assert m.weight.requires_grad


def test_missing_extension_fails_loudly(monkeypatch):
    """No silent fallback: without the built library every op entry raises (and says how to build it)."""
    from crank_b200 import lib

    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libcrank_b200.so")
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        lib.lib()
    with pytest.raises(RuntimeError):
        lib.call("crk_adam_step")
