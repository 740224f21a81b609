"""Host logic of the product (trainers, VQVAE2 / Quantizer orchestration, flat-parameter state dicts, FusedAdam
plumbing, frozen / no-grad passes) for every recipe switch, WITHOUT a GPU: the C-ABI ops are swapped for torch-CPU
stand-ins (tests/cpu_emulation.py, test infrastructure) and the product's train steps are compared with the oracle
port, which is itself bit-identical to the unmodified reference trainers (tests/test_cpu_oracle.py).
The kernels behind the ops are verified separately on the GPU (`-m gpu`)."""
import random

import numpy as np
import pytest
import torch

from tests.cpu_emulation import emulated_ops

S = 5


class _W:
    def add_scalar(self, *a, **k): pass
    def flush(self): pass
    def close(self): pass


def _run(kind, overrides, T, steps=2):
    from crank_b200.conf import vcc2020_conf
    from crank_b200.net.trainer import TrainerWrapper, get_criterion, get_model, get_optimizer, get_scheduler
    from crank_b200.synthetic import clone_batch, make_batch, spkr_dict
    from oracle import crank_port as cp

    conf = vcc2020_conf(trainer_type=kind, n_steps_gan_start=-1, discriminator_dropout=0.0, **overrides)
    torch.manual_seed(11)
    om = cp.build_models(conf, S)
    O = cp.OracleTrainer(kind, om, cp.build_optimizers(conf, om), conf)
    with emulated_ops():
        pm = get_model(conf, S, device="cpu")
        for k in om:
            pm[k].load_state_dict(om[k].state_dict())
        opt = get_optimizer(conf, pm)
        P = TrainerWrapper(kind, model=pm, optimizer=opt, criterion=get_criterion(conf),
                           dataloader={"spkrs": spkr_dict(S)}, writer={"train": _W(), "dev": _W()},
                           expdir="/tmp/crank_b200_hostlogic", conf=conf, feat_conf=conf["feature"],
                           scheduler=get_scheduler(conf, opt), scaler=None, resume=0, device="cpu", n_jobs=1)
        P.tqdm.close()
        worst = 0.0
        for step in range(steps):
            b = make_batch(2, T, S, seed=20 + step, ragged=True)
            random.seed(step)
            ov = O.train(clone_batch(b), "train")
            random.seed(step)
            pv = P.train(clone_batch(b), "train")
            assert set(ov) <= set(pv), sorted(set(ov) - set(pv))
            for k, ref in ov.items():
                err = abs(pv[k] - ref) / max(abs(ref), 1e-3)
                worst = max(worst, err)
                # the stand-ins are not bit-identical to the oracle's modules, and Adam turns rounding-level
                # gradient components into +-lr updates, so quantities evaluated after an update inside the step
                # agree to ~1e-4; a wrong mask / weight / missing term shows up at 1e-2 .. 1
                assert err <= 3e-4, f"{kind} {overrides} step {step}: {k}: product {pv[k]} vs oracle {ref}"
        # every parameter / buffer after the steps, through the reference-keyed state dict
        for k in om:
            sd_o, sd_p = om[k].state_dict(), pm[k].state_dict()
            assert set(sd_o) == set(sd_p), (k, sorted(set(sd_o) ^ set(sd_p))[:6])
            for name, t in sd_o.items():
                d = (sd_p[name].double() - t.double()).abs().max().item()
                scale = t.double().abs().max().item()
                # Adam normalises the update: a gradient component that is ~0 moves by +-lr whatever its rounding, so
                # single elements may differ by up to 2 * lr * steps = 8e-4 in absolute terms (weights are ~0.3)
                assert d <= max(5e-4 * scale, 1e-3), f"{kind} {overrides}: {k}.{name} differs by {d:.2e} (max |ref| {scale:.2e})"
    return worst


@pytest.mark.parametrize("kind", ["vqvae", "lsgan", "cyclegan", "stargan"])
def test_product_train_steps_match_oracle_with_emulated_ops(kind):
    _run(kind, {}, T=96)


@pytest.mark.parametrize("kind,overrides", [
    ("lsgan", dict(causal=True, causal_size=4)),
    ("cyclegan", dict(causal=True, causal_size=4)),
    ("vqvae", dict(n_vq_stacks=1)),
    ("cyclegan", dict(n_vq_stacks=3)),
    ("lsgan", dict(use_spkr_embedding=False)),
    ("cyclegan", dict(encoder_f0=True)),
    ("lsgan", dict(ema_flag=False)),
    ("cyclegan", dict(acgan_flag=True)),
    ("lsgan", dict(acgan_flag=True)),
    ("stargan", dict(cvadv_flag=True)),
    ("lsgan", dict(encoder_detach=True)),
    ("lsgan", dict(train_first="G")),
    ("vqvae", dict(use_cyclic_training=True, n_steps_cycle_start=-1)),
    ("lsgan", dict(optim={m: {"type": "radam"} for m in ("G", "D", "C", "SPKRADV")})),
    ("vqvae", dict(optim={m: {"type": "lamb"} for m in ("G", "D", "C", "SPKRADV")})),
])
def test_product_host_logic_on_config_variants(kind, overrides):
    _run(kind, overrides, T=176 if overrides.get("causal") else 96, steps=1 if overrides.get("causal") else 2)


def test_too_wide_stacks_fail_loudly():
    """The kernels take at most 256 input channels (n_vq_stacks = 3 needs 192: built, see the variant above and
    tests/test_gpu_trainstep.py) and 128 output / aux channels; anything wider says so at construction."""
    from crank_b200.parallel_wavegan.models import ParallelWaveGANGenerator

    with emulated_ops():
        with pytest.raises(NotImplementedError, match="at most 256 input"):
            ParallelWaveGANGenerator(in_channels=320, out_channels=80, kernel_size=5, layers=8, stacks=4, aux_channels=0,
                                     upsample_conditional_features=False)
        with pytest.raises(NotImplementedError, match="at most 256 input"):
            ParallelWaveGANGenerator(in_channels=64, out_channels=80, kernel_size=5, layers=8, stacks=4, aux_channels=192,
                                     upsample_conditional_features=False)


def test_emulation_is_test_only_and_restored():
    from crank_b200 import lib, ops
    from crank_b200.parallel_wavegan import models

    before = (models.WavenetFn, ops.VQFn, ops.MaskedLossFn, lib.require_cuda)
    with emulated_ops():
        assert models.WavenetFn is not before[0]
    assert (models.WavenetFn, ops.VQFn, ops.MaskedLossFn, lib.require_cuda) == before
    with pytest.raises(lib.CrkError):
        ops.masked_l1_mse(torch.zeros(1, 2, 3), 0.0)          # the product still refuses CPU tensors


def test_conversion_conditioning_matches_live_reference():
    """eval / dev / reconstruction: the conditioning tensors built for a TARGET speaker (`_get_enc_h`, `_get_dec_h`,
    log-F0 conversion through the scalers, basetrainer.py:253-320) equal the reference trainer's, and the no-grad
    conversion pass (`_convert`) equals the oracle generator on those conditions."""
    from oracle import refshim

    if not refshim.available():
        pytest.skip("/root/reference is not present on this box")
    from sklearn.preprocessing import StandardScaler

    from crank_b200.conf import vcc2020_conf
    from crank_b200.net.trainer import TrainerWrapper, get_criterion, get_model, get_optimizer, get_scheduler
    from crank_b200.synthetic import clone_batch, make_batch, spkr_dict
    from oracle import crank_port as cp

    refshim.install()
    tr = refshim.ref("crank.net.trainer")
    tu = refshim.ref("crank.net.trainer.utils")
    conf = vcc2020_conf(trainer_type="vqvae", encoder_f0=True)
    rs = np.random.RandomState(0)
    names = list(spkr_dict(S).keys())
    scaler = {"lcf0": StandardScaler().fit(5.0 + 0.3 * rs.randn(500, 1))}
    for i, n in enumerate(names):
        scaler[n] = {"lcf0": StandardScaler().fit(4.6 + 0.2 * i + (0.15 + 0.02 * i) * rs.randn(300, 1))}
    torch.manual_seed(2)
    om = cp.build_models(conf, S)
    ropt = tu.get_optimizer(conf, om)
    R = tr.TrainerWrapper("vqvae", model=om, optimizer=ropt, criterion=tu.get_criterion(conf, device="cpu"),
                          dataloader={"spkrs": spkr_dict(S)}, writer={"train": _W(), "dev": _W()}, expdir="/tmp/exp",
                          conf=conf, feat_conf=conf["feature"], scheduler=tu.get_scheduler(conf, ropt), scaler=scaler,
                          resume=0, device="cpu", n_jobs=1)
    R.tqdm.close()
    b = make_batch(3, 64, S, seed=9, ragged=True)
    target = names[3]
    with emulated_ops():
        pm = get_model(conf, S, device="cpu")
        for k in om:
            pm[k].load_state_dict(om[k].state_dict())
        popt = get_optimizer(conf, pm)
        P = TrainerWrapper("vqvae", model=pm, optimizer=popt, criterion=get_criterion(conf),
                           dataloader={"spkrs": spkr_dict(S)}, writer={"train": _W(), "dev": _W()}, expdir="/tmp/exp",
                           conf=conf, feat_conf=conf["feature"], scheduler=get_scheduler(conf, popt), scaler=scaler,
                           resume=0, device="cpu", n_jobs=1)
        P.tqdm.close()
        for kw in (dict(cv_spkr_name=target), dict(use_cvfeats=True), dict()):
            re_h, pe_h = R._get_enc_h(clone_batch(b), **kw), P._get_enc_h(clone_batch(b), **kw)
            assert torch.allclose(re_h, pe_h, atol=2e-5), kw
            (rd, rh), (pd, ph) = R._get_dec_h(clone_batch(b), **kw), P._get_dec_h(clone_batch(b), **kw)
            assert torch.allclose(rd, pd, atol=2e-5) and torch.equal(rh, ph), kw
        for p_ in pm.values():
            p_.eval()
        for o_ in om.values():
            o_.eval()
        out = P.eval(clone_batch(b))[target]            # @torch.no_grad entry point: one conversion per target speaker
        enc_h = R._get_enc_h(clone_batch(b), cv_spkr_name=target)
        dec_h, spkrvec = R._get_dec_h(clone_batch(b), cv_spkr_name=target)
        with torch.no_grad():
            ref = om["G"].forward(b["in_feats"], enc_h, dec_h, spkrvec=spkrvec)
        assert not out["decoded"].requires_grad
        assert torch.allclose(out["decoded"], ref["decoded"], atol=1e-4, rtol=1e-4)
        for n in range(conf["n_vq_stacks"]):
            assert torch.equal(out["qidx"][n], ref["qidx"][n])


def test_capturable_fused_adam_equals_default():
    """FusedAdam(capturable=True) keeps the step count in a device tensor (crk_adam_step_dev, graph-capturable);
    the update must be the one of the default host-counter path, step after step, incl. a late switch-over."""
    from crank_b200.net.trainer.optim import FusedAdam

    torch.manual_seed(0)
    w0 = torch.randn(257)
    grads = [torch.randn(257) for _ in range(4)]
    with emulated_ops():
        pa, pb = torch.nn.Parameter(w0.clone()), torch.nn.Parameter(w0.clone())
        oa, ob = FusedAdam([pa], lr=2e-4), FusedAdam([pb], lr=2e-4, capturable=True)
        for i, g in enumerate(grads):
            pa.grad, pb.grad = g.clone(), g.clone()
            oa.step()
            ob.step()
            assert torch.equal(pa, pb), i
        assert int(ob.state[pb]["step_dev"].item()) == 4
        # an optimizer that ran eagerly first and becomes capturable later continues from its host step count
        oa.capturable = True
        pa.grad, pb.grad = grads[0].clone(), grads[0].clone()
        oa.step()
        ob.step()
        assert torch.equal(pa, pb) and int(oa.state[pa]["step_dev"].item()) == 5


def test_graphed_step_control_flow_with_a_fake_capture(monkeypatch):
    """net/graph.py (experimental, not yet run on hardware): warm-up count, capture-once-per-signature, static input
    buffers, loss-vector assembly, with torch.cuda's graph objects replaced by a fake: "capturing" executes nothing
    (the step body hands back the loss dict of the eager step that preceded it, like a real capture records without
    running) and replay() re-runs the step on the static batch.  Values must equal plain trainer.train."""
    import contextlib

    from crank_b200.conf import vcc2020_conf
    from crank_b200.net import graph as G
    from crank_b200.net.trainer import TrainerWrapper, get_criterion, get_model, get_optimizer, get_scheduler
    from crank_b200.synthetic import clone_batch, make_batch, spkr_dict

    flag = {"capturing": False}

    class FakeGraph:
        closure = None

        def replay(self):
            self.closure()

    class FakeStream:
        def wait_stream(self, other): pass

    @contextlib.contextmanager
    def fake_capture(g):
        flag["capturing"] = True
        yield
        flag["capturing"] = False

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: FakeStream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "CUDAGraph", FakeGraph)
    monkeypatch.setattr(torch.cuda, "graph", fake_capture)

    conf = vcc2020_conf(trainer_type="lsgan", n_steps_gan_start=-1, discriminator_dropout=0.0)

    def build():
        torch.manual_seed(4)
        m = get_model(conf, S, device="cpu")
        opt = get_optimizer(conf, m)
        t = TrainerWrapper("lsgan", model=m, optimizer=opt, criterion=get_criterion(conf),
                           dataloader={"spkrs": spkr_dict(S)}, writer={"train": _W(), "dev": _W()},
                           expdir="/tmp/crank_b200_graph", conf=conf, feat_conf=conf["feature"],
                           scheduler=get_scheduler(conf, opt), scaler=None, resume=0, device="cpu", n_jobs=1)
        t.tqdm.close()
        return t

    batches = [make_batch(2, 96, S, seed=40 + i, ragged=True) for i in range(6)]
    with emulated_ops():
        ref = build()
        ref_vals = [ref.train(clone_batch(b), "train") for b in batches]
        t = build()
        step = G.GraphedTrainStep(t)
        assert all(o.capturable for o in t.optimizer.values())
        orig_core = t._train_core
        last = {}

        def core(batch, phase):
            if not flag["capturing"]:
                last["loss"] = orig_core(batch, phase)
            return last["loss"]

        t._train_core = core
        vals = []
        for i, b in enumerate(batches):
            vals.append(step(clone_batch(b)))
            if i == G.GraphedTrainStep.WARMUP:            # the capture call just happened: wire the fake replay
                (graph, static, keys, packed, const), = step._graphs.values()

                def replay_closure(static=static, keys=keys, packed=packed):
                    loss = orig_core(static, "train")
                    packed.copy_(torch.stack([loss[k].detach().reshape(()).float() for k in keys]))

                graph.closure = replay_closure
        assert len(step._graphs) == 1
    for i, (a, b) in enumerate(zip(vals, ref_vals)):
        for k in b:
            assert abs(a[k] - b[k]) <= 3e-4 * max(abs(b[k]), 1e-3), (i, k, a[k], b[k])


def test_offline_mlfb_extraction_wrapper_against_the_reference_fixture():
    """crank_b200.feature.extract_mlfb (symmetric hann, centred reflect padding, Slaney basis) reproduces the mlfb the
    reference stored for test/data/SF1_10001.wav (committed copy) -- host logic only: the log-mel kernel behind
    ops.logmel is emulated here and verified on the GPU in tests/test_gpu_kernels.py."""
    import os

    from crank_b200.feature import extract_mlfb, symmetric_hann
    from oracle import mel as omel

    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_fixture_mlfb.npz"))
    raw, ref = fx["raw_i16"].astype(np.float64) / 32768.0, fx["mlfb"]
    assert np.abs(symmetric_hann(1024) - omel.hann(1024, periodic=False)).max() < 1e-15
    with emulated_ops():
        got = extract_mlfb(raw, fs=22050, device="cpu").numpy()
        both = extract_mlfb(np.stack([raw, raw[::-1].copy()]), fs=22050, device="cpu")
    assert got.shape == ref.shape == (1057, 80)
    # fp32 against the float64 fixture; quiet bins sit near the eps floor where log10 amplifies rounding
    assert np.abs(got - ref).max() < 2e-3 and np.abs(got - ref).mean() < 2e-5, (np.abs(got - ref).max(), np.abs(got - ref).mean())
    assert both.shape == (2, 1057, 80) and torch.allclose(both[0], torch.from_numpy(got))


def test_mel_band_tables_reproduce_the_dense_basis():
    """ops._mel_bands (the banded form of the mel basis the fused log-mel kernels read, and its transpose for the backward):
    expanding the runs gives back the dense (bins, n_mels) matrix exactly."""
    import numpy as np
    import torch

    from crank_b200 import ops
    from crank_b200.net.module.mlfb import mel_basis

    for fs, fmin, fmax in ((24000, 80, 7600), (22050, 80, 7600), (16000, 0, None)):
        w = mel_basis(fs, 1024, 80, fmin, fmax).T.copy()            # (513, 80)
        st, ln, of, bw, nnz, tst, tln, tof, tbw = ops._mel_bands(torch.from_numpy(w))
        assert nnz <= ops.MEL_FUSED_MAXNNZ
        dense = np.zeros_like(w)
        for m in range(80):
            dense[int(st[m]):int(st[m]) + int(ln[m]), m] = bw[int(of[m]):int(of[m]) + int(ln[m])].numpy()
        assert np.array_equal(dense, w)
        dense_t = np.zeros_like(w)
        for k in range(513):
            dense_t[k, int(tst[k]):int(tst[k]) + int(tln[k])] = tbw[int(tof[k]):int(tof[k]) + int(tln[k])].numpy()
        assert np.array_equal(dense_t, w)


def test_stft_and_mlfb_layers_stand_alone_match_the_fused_definition():
    """STFTLayer.forward / MLFBLayer.forward (mlfb.py:36-110) compose to the same log-mel the oracle computes."""
    import numpy as np
    import torch

    from crank_b200.net.module.mlfb import MLFBLayer, STFTLayer
    from oracle import mel as omel

    g = torch.Generator().manual_seed(0)
    x = 0.1 * torch.randn(2, 4000, generator=g)
    st = STFTLayer(fs=24000, hop_size=128, fft_size=1024, center=False)
    ml = MLFBLayer(fs=24000, fft_size=1024, n_mels=80, fmin=80, fmax=7600)
    spec = st(x)
    got = ml(torch.sqrt(spec[..., 0] ** 2 + spec[..., 1] ** 2)).numpy()
    basis = omel.mel_basis(24000, 1024, 80, 80, 7600).T.astype(np.float64)
    ref = np.stack([np.log10(np.maximum(1e-10, omel.stft_mag(x[b].double().numpy(), 1024, 128, omel.hann(1024, True), center=False) @ basis))
                    for b in range(2)])
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-4


def test_dev_wav_generation_host_logic(tmp_path):
    """BaseTrainer._generate_cvwav (basetrainer.py:322-420 of the reference): inverse scaler, grouping of equal-length
    utterances into one Griffin-Lim batch, sample selection, file naming and 16-bit WAV writing -- on CPU with emulated
    ops (the Griffin-Lim itself is plain torch; checked against the oracle in tests/test_cpu_griffinlim.py)."""
    import wave

    from crank_b200.conf import vcc2020_conf
    from crank_b200.net.trainer import TrainerWrapper, get_criterion, get_model, get_optimizer, get_scheduler
    from crank_b200.synthetic import make_batch, spkr_dict

    class _Scaler:      # sklearn StandardScaler surface
        mean_ = np.linspace(-1.0, 1.0, 80)
        scale_ = np.linspace(0.5, 1.5, 80)

    conf = vcc2020_conf(trainer_type="vqvae")
    with emulated_ops():
        pm = get_model(conf, S, device="cpu")
        opt = get_optimizer(conf, pm)
        P = TrainerWrapper("vqvae", model=pm, optimizer=opt, criterion=get_criterion(conf),
                           dataloader={"spkrs": spkr_dict(S)}, writer={"train": _W(), "dev": _W()},
                           expdir=str(tmp_path), conf=conf, feat_conf=conf["feature"],
                           scheduler=get_scheduler(conf, opt), scaler={"mlfb": _Scaler()}, resume=0, device="cpu", n_jobs=1)
        P.tqdm.close()
        batch = make_batch(4, 64, S, seed=3, ragged=False)
        batch["flen"] = torch.tensor([64, 40, 64, 40])
        with torch.no_grad():
            out = P._convert(batch, None)
            # keep the log-mel in a sane range so that 10**x does not overflow in the pinv projection
            out["decoded"] = out["decoded"].clamp(-2.0, 2.0)
            none = P._generate_cvwav(batch, out, None, tdir="eval_wav", save_decoded=False)
            wavs = P._generate_cvwav(batch, out, "spk1", tdir="dev_wav", save_hdf5=False, n_samples=-1)
    assert none == {}
    assert len(wavs) == 4
    hop = conf["feature"]["hop_size"]
    for n, (path, y) in enumerate(sorted(wavs.items(), key=lambda kv: str(kv[0]))):
        assert str(path).endswith("_cv-spk1.wav") and "dev_wav" in str(path)
        assert torch.isfinite(y).all()
        with wave.open(str(path), "rb") as f:
            assert f.getframerate() == conf["feature"]["fs"] and f.getsampwidth() == 2 and f.getnframes() == y.numel()
    lengths = sorted(int(y.numel()) for y in wavs.values())
    assert lengths == sorted([hop * 39, hop * 39, hop * 63, hop * 63])


@pytest.mark.parametrize("window", ["hann", "param", "conv"])
def test_raw_front_end_window_types_host_logic(window):
    """LogMelFilterBankLayer (mlfb.py:72-171) for the three window types with emulated ops: `hann` runs without autograd,
    `param` exposes a learnable window parameter (state-dict key stft_layer.window) that receives a gradient, `conv` a learnable
    pre-filter (stft_layer.window_conv.0.*) that does; centre padding and the scaler epilogue as in the reference layer."""
    from crank_b200.net.module.mlfb import LogMelFilterBankLayer

    class _Scaler:
        mean_ = np.zeros(80)
        var_ = np.ones(80) * 4.0

    torch.manual_seed(2)
    with emulated_ops():
        layer = LogMelFilterBankLayer(fs=24000, hop_size=128, fft_size=1024, win_length=1024, window=window, center=True,
                                      n_mels=80, fmin=80, fmax=7600, scaler=_Scaler())
        x = 0.1 * torch.randn(2, 4096)
        out = layer(x)
        assert out.shape == (2, 1 + 4096 // 128, 80)
        keys = set(layer.state_dict())
        if window == "hann":
            assert not out.requires_grad and not any(k.startswith("stft_layer.window") for k in keys)
        else:
            out.square().mean().backward()
            if window == "param":
                assert "stft_layer.window" in keys
                g = layer.stft_layer.window.grad
            else:
                assert {"stft_layer.window_conv.0.weight", "stft_layer.window_conv.0.bias"} <= keys
                g = layer.stft_layer.window_conv[0].weight.grad
            assert g is not None and torch.isfinite(g).all() and g.abs().max() > 0
        # the scaler epilogue: (x - 0) / 2
        layer.scaler_layer = None
        assert torch.allclose(layer(x).detach() / 2.0, out.detach(), atol=1e-5)
