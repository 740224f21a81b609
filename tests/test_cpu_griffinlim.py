"""Griffin-Lim (eval path, crank/utils/utils.py:210-269): the product's torch implementation against the numpy oracle with
identical initial phases, and the oracle's own structural pins."""
import numpy as np
import torch

from oracle import griffinlim as ogl
from oracle import mel as omel


def _signal(n=6000, fs=24000):
    t = np.arange(n) / fs
    return 0.4 * np.sin(2 * np.pi * 330 * t) + 0.2 * np.sin(2 * np.pi * 1234 * t + 0.3) + 0.01 * np.random.RandomState(0).randn(n)


def test_oracle_stft_istft_round_trip():
    x = _signal()
    win = omel.hann(1024, periodic=True)
    D = ogl.stft(x, 1024, 128, win)
    y = ogl.istft(D, 1024, 128, win)
    assert np.abs(y - x[:len(y)]).max() < 1e-10


def test_product_griffin_lim_matches_oracle_with_same_initial_phases():
    from crank_b200.utils import griffin_lim

    x = _signal()
    win = omel.hann(1024, periodic=True)
    S = np.abs(ogl.stft(x, 1024, 128, win)).T                      # (T, bins)
    rs = np.random.RandomState(1)
    ang = np.exp(2j * np.pi * rs.rand(S.shape[1], S.shape[0]))
    ref = ogl.griffin_lim(S, 1024, 128, 1024, ang, n_iters=8)
    got = griffin_lim(torch.from_numpy(S), 1024, 128, 1024, n_iters=8, init_angles=torch.from_numpy(ang)).numpy()
    assert got.shape == ref.shape
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 2e-3, err                                          # complex64 iteration vs float64 oracle
    # batched call == per-utterance calls
    both = griffin_lim(torch.from_numpy(np.stack([S, S])), 1024, 128, 1024, n_iters=8,
                       init_angles=torch.from_numpy(np.stack([ang, ang]))).numpy()
    assert np.abs(both[0] - got).max() < 1e-5 and np.abs(both[1] - got).max() < 1e-5


def test_griffin_lim_improves_spectral_convergence():
    from crank_b200.utils import griffin_lim

    x = _signal()
    win = omel.hann(1024, periodic=True)
    S = np.abs(ogl.stft(x, 1024, 128, win)).T

    def sc(y):
        R = np.abs(ogl.stft(y, 1024, 128, win)).T
        n = min(len(R), len(S))
        return np.linalg.norm(R[:n] - S[:n]) / np.linalg.norm(S[:n])

    g = torch.Generator().manual_seed(0)
    y1 = griffin_lim(torch.from_numpy(S), 1024, 128, 1024, n_iters=1, generator=g).numpy()
    g = torch.Generator().manual_seed(0)
    y30 = griffin_lim(torch.from_numpy(S), 1024, 128, 1024, n_iters=30, generator=g).numpy()
    assert sc(y30) < 0.5 * sc(y1) and sc(y30) < 0.15, (sc(y1), sc(y30))


def test_logmel_to_linear_matches_oracle():
    from crank_b200.utils import logmelspc_to_linearspc

    rs = np.random.RandomState(2)
    lm = rs.randn(50, 80) * 0.5 - 2.0
    ref = ogl.logmelspc_to_linearspc(lm, 24000, 80, 1024, 80, 7600)
    got = logmelspc_to_linearspc(torch.from_numpy(lm), 24000, 80, 1024, 80, 7600).numpy()
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-4
