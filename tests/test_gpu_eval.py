"""Eval path on the device: Griffin-Lim of decoded mlfb (crank/utils/utils.py:94-107,
210-269; basetrainer.py:322-420) vs the numpy oracle, and the trainer's dev() hook writing WAV files."""
import os
import wave

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_griffin_lim_on_device_matches_oracle():
    from crank_b200.utils import griffin_lim, logmelspc_to_linearspc
    from oracle import griffinlim as ogl
    from oracle import mel as omel

    rs = np.random.RandomState(3)
    t = np.arange(9000) / 24000.0
    x = 0.4 * np.sin(2 * np.pi * 440 * t) + 0.15 * np.sin(2 * np.pi * 2100 * t) + 0.01 * rs.randn(len(t))
    win = omel.hann(1024, periodic=True)
    S = np.abs(ogl.stft(x, 1024, 128, win)).T
    ang = np.exp(2j * np.pi * rs.rand(S.shape[1], S.shape[0]))
    ref = ogl.griffin_lim(S, 1024, 128, 1024, ang, n_iters=10)
    got = griffin_lim(torch.from_numpy(S).cuda(), 1024, 128, 1024, n_iters=10,
                      init_angles=torch.from_numpy(ang).cuda())
    assert got.is_cuda
    err = np.abs(got.cpu().numpy() - ref).max() / np.abs(ref).max()
    print(f"device Griffin-Lim vs float64 oracle (10 iterations, same initial phases): rel err {err:.2e}")
    assert err < 2e-3, err
    lm = rs.randn(40, 80) * 0.5 - 2.0
    a = logmelspc_to_linearspc(torch.from_numpy(lm).cuda(), 24000, 80, 1024, 80, 7600).cpu().numpy()
    b = ogl.logmelspc_to_linearspc(lm, 24000, 80, 1024, 80, 7600)
    assert np.abs(a - b).max() / np.abs(b).max() < 1e-4


def test_dev_step_writes_griffin_lim_wavs(tmp_path):
    from crank_b200.synthetic import make_batch, to_device
    from tests.test_gpu_trainstep import _build_pair

    conf, om, pm, O, P = _build_pair("vqvae")
    P.expdir = tmp_path
    P.n_dev_samples = 2
    P.n_cv_spkrs = 1
    batch = to_device(make_batch(3, 96, len(P.spkrs), seed=5, ragged=True), "cuda")
    if not hasattr(P, "scaler") or P.scaler is None:
        P.scaler = None
    # dev() needs the F0 conversion statistics only when a speaker is converted; exercise the hook directly on a
    # reconstruction forward (cv_spkr_name=None), which is what `reconstruction` / crank/bin/griffin_lim.py feed it
    with torch.no_grad():
        out = P._convert(batch, None)
        wavs = P._generate_cvwav(batch, out, None, tdir="dev_wav", save_hdf5=False, n_samples=2)
    assert len(wavs) == 2
    for path, y in wavs.items():
        assert os.path.exists(path), path
        assert y.is_cuda and torch.isfinite(y).all()
        with wave.open(str(path), "rb") as f:
            assert f.getframerate() == conf["feature"]["fs"] and f.getsampwidth() == 2
            assert f.getnframes() == y.numel() > 0
