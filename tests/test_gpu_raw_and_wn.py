"""GPU tests of two boundary behaviours the round-1 suite never executed (verdict parity items 1c, 1d):

  * `VQVAE2(conf, use_raw=True).forward(raw, ...)`: the on-the-fly log-mel front end inside the generator
    (crank/net/module/vqvae2.py:23-35,56-67), shaped like the reference's own test/test_vqvae.py:34-69
    (BASELINE config 1: the 135 294-sample SF1_10001 wav -> T = 1049 frames, 10 speakers).  The reference test
    feeds a batch-1 waveform with batch-3 conditioning and relies on broadcasting inside the aux add; here the
    waveform is repeated to the conditioning's batch so that every operand has the same batch.
  * `remove_weight_norm()` (crank/net/module/vqvae2.py:192-195 -> parallel_wavegan's remove_weight_norm, restated at
    oracle/pwg.py): the function computed by the stack is unchanged and the state dict switches to plain `weight`.
"""
import os

import numpy as np
import pytest
import torch

from tests.util import assert_close, rel_err

pytestmark = pytest.mark.gpu
FIX = os.path.join(os.path.dirname(__file__), "golden", "ref_fixture_mlfb.npz")


def test_vqvae2_use_raw_forward_matches_oracle_front_end_plus_generator():
    from crank_b200.conf import vcc2020_conf
    from crank_b200.net.module.vqvae2 import VQVAE2
    from oracle import crank_port as cp
    from oracle import mel as omel

    S, B = 10, 3
    conf = vcc2020_conf(trainer_type="vqvae", use_raw=True, input_feat_type="mlfb", ignore_scaler=["raw"],
                        use_preprocessed_scaler=False)
    conf["feature"] = dict(conf["feature"], fs=22050)
    raw = np.load(FIX)["raw_i16"].astype(np.float64) / 32768.0
    feat = conf["feature"]
    # the online layer: periodic hann, center=False (vqvae2.py:56-67, mlfb.py:100-101)
    spc = omel.stft_mag(raw, feat["fftl"], feat["hop_size"], omel.hann(feat["win_length"], periodic=True), center=False)
    basis = omel.mel_basis(feat["fs"], feat["fftl"], feat["mlfb_dim"], feat["fmin"], feat["fmax"])
    mlfb = np.log10(np.maximum(1e-10, spc @ basis.T.astype(np.float64)))
    T = mlfb.shape[0]
    assert T == 1049                                               # test_vqvae.py:52 (1057 - 8)

    torch.manual_seed(1234)
    oconf = dict(conf, use_raw=False)
    og = cp.VQVAE2(oconf, spkr_size=S)
    pg = VQVAE2(conf, spkr_size=S).cuda()
    missing = pg.load_state_dict(og.state_dict(), strict=False)
    assert not [k for k in missing.missing_keys if "preprocess_layer" not in k], missing
    g = torch.Generator().manual_seed(3)
    dec_h = torch.randn(B, T, 2, generator=g)
    spkrvec = (torch.ones(B, T) * 5).long()
    x_mlfb = torch.from_numpy(mlfb).float()[None].expand(B, T, 80).contiguous()
    x_raw = torch.from_numpy(raw).float()[None].expand(B, -1).contiguous()
    with torch.no_grad():
        oo = og.forward(x_mlfb, None, dec_h, spkrvec=spkrvec)
        po = pg.forward(x_raw.cuda(), None, dec_h.cuda(), spkrvec=spkrvec.cuda())
    assert po["decoded"].shape == (B, T, 80)
    for n in range(conf["n_vq_stacks"]):
        diff = (po["qidx"][n].cpu() != oo["qidx"][n]).float().mean().item()
        print(f"use_raw: stack {n}: fraction of VQ indices differing from the oracle {diff:.2e}")
        # the front ends differ by ~1e-5 (fp32 device STFT vs float64 oracle): only near-tie frames may flip
        assert diff <= 5e-3
        assert_close(po["encoded_unmod"][n], oo["encoded_unmod"][n], 1e-4, f"encoded_unmod{n}")
    e = rel_err(po["decoded"], oo["decoded"])
    print(f"use_raw: decoded rel err {e:.2e}")
    assert e <= 2e-3          # a flipped near-tie code changes a few frames' decoder input; encoder outputs are pinned at 1e-4


@pytest.mark.parametrize("kind", ["generator", "residual_d", "conv_d"])
def test_remove_weight_norm_keeps_the_function_and_switches_the_state_dict(kind):
    from crank_b200.parallel_wavegan import models as pm
    from oracle import pwg

    torch.manual_seed(4)
    if kind == "generator":
        kw = dict(in_channels=80, out_channels=64, kernel_size=5, layers=4, stacks=2, aux_channels=2,
                  aux_context_window=0, upsample_conditional_features=False)
        o, p = pwg.ParallelWaveGANGenerator(**kw), pm.ParallelWaveGANGenerator(**kw).cuda()
        args = (torch.randn(2, 80, 150), torch.randn(2, 2, 150))
    elif kind == "residual_d":
        kw = dict(in_channels=113, out_channels=1, kernel_size=5, layers=4, stacks=2, dropout=0.0)
        o, p = pwg.ResidualParallelWaveGANDiscriminator(**kw), pm.ResidualParallelWaveGANDiscriminator(**kw).cuda()
        args = (torch.randn(2, 113, 150),)
    else:
        kw = dict(in_channels=80, out_channels=14, kernel_size=5, layers=4, conv_channels=64)
        o, p = pwg.ParallelWaveGANDiscriminator(**kw), pm.ParallelWaveGANDiscriminator(**kw).cuda()
        args = (torch.randn(2, 80, 150),)
    # make g != ||v|| so that folding actually changes the stored tensors
    with torch.no_grad():
        for name, prm in o.named_parameters():
            if name.endswith("weight_g"):
                prm.mul_(1.0 + 0.3 * torch.rand_like(prm))
    p.load_state_dict(o.state_dict())
    cargs = tuple(a.cuda() for a in args)
    with torch.no_grad():
        y_o, y_p = o(*args), p(*cargs)
    assert_close(y_p, y_o, 1e-4, "before removal")
    o.remove_weight_norm()
    p.remove_weight_norm()
    with torch.no_grad():
        y_o2, y_p2 = o(*args), p(*cargs)
    assert_close(y_p2, y_o2, 1e-4, "after removal vs oracle")
    assert_close(y_p2, y_p, 1e-5, "function unchanged by removal")
    so, sp = o.state_dict(), p.state_dict()
    assert set(so) == set(sp), set(so) ^ set(sp)
    assert not any(k.endswith("weight_g") or k.endswith("weight_v") for k in sp)
    for k, v in so.items():
        assert_close(sp[k], v, 1e-5, k)
