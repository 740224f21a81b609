"""GPU parity: each C-ABI kernel family against the CPU oracle on the same seeded inputs.

Tolerances (written here as the contract): fp32 tensors <= 1e-4 relative (north star:
"<=1e-4 rel on reconstruction / loss tensors"); in practice the fp32 CUDA-core path lands ~1e-6.
VQ indices: bit-exact on a spread (EMA-warmed) codebook.
"""
import numpy as np
import pytest
import torch

from tests.util import assert_close, compare_conv_grads, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_cuda_core_path():
    """this file pins the fp32 CUDA-core kernels; tests/test_gpu_tc.py repeats it on the tensor cores"""
    from crank_b200 import lib as L

    L.set_precision("fp32")
    yield
    L.set_precision("tf32x3")

TOL = 1e-4


def _dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _pair_generator(in_ch, out_ch, aux, k, layers, stacks, causal=False, seed=0):
    from crank_b200.parallel_wavegan import models as pm
    from oracle import pwg

    torch.manual_seed(seed)
    kw = dict(in_channels=in_ch, out_channels=out_ch, kernel_size=k, layers=layers, stacks=stacks,
              aux_channels=aux, aux_context_window=0, dropout=0.0, use_causal_conv=causal,
              upsample_conditional_features=False)
    o = pwg.ParallelWaveGANGenerator(**kw)
    # perturb g and bias so that weight-norm and bias paths are non-trivial
    with torch.no_grad():
        for n, p in o.named_parameters():
            if n.endswith("weight_g"):
                p.mul_(1.0 + 0.2 * torch.randn_like(p))
            if n.endswith("bias"):
                p.add_(0.1 * torch.randn_like(p))
    p = pm.ParallelWaveGANGenerator(**kw)
    p.load_state_dict(o.state_dict())
    return o, p.to(_dev())


@pytest.mark.parametrize(
    "in_ch,out_ch,aux,k,layers,stacks,causal,B,T",
    [
        (80, 64, 0, 5, 8, 4, False, 3, 100),    # encoder 0
        (64, 64, 0, 3, 6, 3, False, 2, 130),    # encoder 1 / decoder 1
        (128, 80, 34, 5, 8, 4, False, 2, 75),   # decoder 0 (aux = f0 + speaker embedding)
        (80, 64, 2, 5, 4, 2, True, 2, 70),      # causal + encoder_f0
        (80, 64, 0, 3, 2, 1, False, 1, 17),     # tiny T (< one tile)
    ],
)
def test_wavenet_stack_fwd_bwd(in_ch, out_ch, aux, k, layers, stacks, causal, B, T):
    o, p = _pair_generator(in_ch, out_ch, aux, k, layers, stacks, causal)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, in_ch, T, generator=g)
    c = torch.randn(B, aux, T, generator=g) if aux > 0 else None
    xo = x.clone().requires_grad_(True)
    co = c.clone().requires_grad_(True) if c is not None else None
    yo = o(xo, co)
    xp = x.to(_dev()).requires_grad_(True)
    cp = c.to(_dev()).requires_grad_(True) if c is not None else None
    yp = p(xp, cp)
    assert yp.shape == yo.shape
    assert_close(yp, yo, TOL, "stack forward")
    dy = torch.randn(yo.shape, generator=g)
    yo.backward(dy)
    yp.backward(dy.to(_dev()))
    assert_close(xp.grad, xo.grad, TOL, "dx")
    if c is not None:
        assert_close(cp.grad, co.grad, TOL, "dc")
    compare_conv_grads(p, o, TOL, "stack")


def test_residual_discriminator_fwd_bwd_with_injected_dropout():
    from crank_b200.parallel_wavegan import models as pm
    from oracle import pwg

    torch.manual_seed(3)
    kw = dict(in_channels=113, out_channels=1, kernel_size=5, layers=8, stacks=4)
    o = pwg.ResidualParallelWaveGANDiscriminator(dropout=0.0, **kw)
    p = pm.ResidualParallelWaveGANDiscriminator(dropout=0.25, **kw)
    p.load_state_dict(o.state_dict())
    p = p.to(_dev())
    B, T = 2, 90
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, 113, T, generator=g)
    # identical dropout masks on both sides: oracle gets them through a forward-pre hook on each block
    keep = 0.75
    masks = (torch.rand(8, B * T, 64, generator=g) < keep).float() / keep
    hooks = []
    for l, blk in enumerate(o.conv_layers):
        def pre(mod, args, l=l):
            xx, cc = args
            m = masks[l].view(B, T, 64).transpose(1, 2)
            mod._residual = xx
            return (xx * m, cc)

        def post(mod, args, out):
            # residual connection must use the un-dropped input: x_out = (out1x1 + residual)*sqrt(.5)
            xx_dropped = args[0]
            y, s = out
            return (y + (mod._residual - xx_dropped) * (0.5 ** 0.5), s)

        hooks.append(blk.register_forward_pre_hook(pre))
        hooks.append(blk.register_forward_hook(post))
    xo = x.clone().requires_grad_(True)
    yo = o(xo)
    xp = x.to(_dev()).requires_grad_(True)
    yp = p.forward_cl(xp.transpose(1, 2), dropmul=masks.to(_dev())).transpose(1, 2)
    assert_close(yp, yo, TOL, "D forward")
    dy = torch.randn(yo.shape, generator=g)
    yo.backward(dy)
    yp.backward(dy.to(_dev()))
    assert_close(xp.grad, xo.grad, TOL, "D dx")
    compare_conv_grads(p, o, TOL, "D")
    for h in hooks:
        h.remove()


@pytest.mark.parametrize("in_ch,out_ch,k,layers,B,T", [(80, 14, 5, 8, 2, 150), (128, 14, 3, 3, 3, 70), (80, 12, 5, 8, 1, 64)])
def test_convstack_fwd_bwd(in_ch, out_ch, k, layers, B, T):
    from crank_b200.parallel_wavegan import models as pm
    from oracle import pwg

    torch.manual_seed(5)
    kw = dict(in_channels=in_ch, out_channels=out_ch, kernel_size=k, layers=layers, conv_channels=64,
              dilation_factor=1)
    o = pwg.ParallelWaveGANDiscriminator(**kw)
    with torch.no_grad():
        for n, q in o.named_parameters():
            if n.endswith("bias"):
                q.add_(0.1 * torch.randn_like(q))
    p = pm.ParallelWaveGANDiscriminator(**kw)
    p.load_state_dict(o.state_dict())
    p = p.to(_dev())
    g = torch.Generator().manual_seed(6)
    x = torch.randn(B, in_ch, T, generator=g)
    xo = x.clone().requires_grad_(True)
    yo = o(xo)
    xp = x.to(_dev()).requires_grad_(True)
    yp = p.forward_cl(xp.transpose(1, 2), grad_scale=-0.1).transpose(1, 2)
    assert_close(yp, yo, TOL, "convstack forward")
    dy = torch.randn(yo.shape, generator=g)
    yo.backward(dy)
    yp.backward(dy.to(_dev()))
    assert_close(xp.grad, -0.1 * xo.grad, TOL, "convstack dx (gradient reversal folded in)")
    compare_conv_grads(p, o, TOL, "convstack")


# ---- vector quantiser -------------------------------------------------------------------------
def _warm_quantizers(seed=0, K=512, D=64, calls=3, F=4000):
    from crank_b200.net.module.vqvae2 import Quantizer as PQ
    from oracle.crank_port import Quantizer as OQ

    torch.manual_seed(seed)
    oq = OQ(D, K, ema_flag=True, bdt_flag=True)
    pq = PQ(D, K, ema_flag=True, bdt_flag=True)
    pq.load_state_dict(oq.state_dict())
    pq = pq.to(_dev())
    g = torch.Generator().manual_seed(seed + 1)
    return oq, pq, g


def test_vq_ema_buffers_and_exact_indices_on_warm_codebook():
    oq, pq, g = _warm_quantizers()
    B, T, D = 4, 1000, 64
    mism_total = 0
    for call in range(4):
        x = torch.randn(B, D, T, generator=g)
        eo, qo, io = oq(x)
        ep, qp, ip = pq(x.to(_dev()))
        mism = (ip.cpu() != io).sum().item()
        # call 0 runs on the fresh-init codebook U(-1/K, 1/K), where ties are decided by BLAS rounding
        # (covered by test_vq_near_ties_only_on_fresh_codebook); later calls must be exact
        if call > 0:
            mism_total += mism
        # all three outputs on agreeing frames
        agree = (ip.cpu() == io)
        assert agree.float().mean().item() > 0.999
        assert_close(ep.cpu()[agree], eo[agree], 1e-5, "embed_idx")
        assert_close(qp.cpu().transpose(1, 2)[agree], qo.transpose(1, 2)[agree], 1e-5, "embed_idx_qx")
        if mism:
            # the EMA statistics of mismatching frames differ; re-sync the product state so later calls
            # test the kernel, not the divergence
            pq.load_state_dict(oq.state_dict())
            pq.to(_dev())
        else:
            assert_close(pq.ema_size, oq.ema_size, 1e-5, "ema_size")
            assert_close(pq.ema_w, oq.ema_w, 1e-5, "ema_w")
            assert_close(pq.embedding.weight, oq.embedding.weight, 1e-5, "codebook")
    assert mism_total == 0, f"{mism_total} index mismatches on EMA-warmed codebooks"


def test_vq_near_ties_only_on_fresh_codebook():
    from crank_b200.net.module.vqvae2 import Quantizer as PQ
    from oracle.crank_port import Quantizer as OQ

    torch.manual_seed(11)
    K, D = 512, 64
    oq = OQ(D, K, ema_flag=False, bdt_flag=False)
    pq = PQ(D, K, ema_flag=False, bdt_flag=False)
    pq.load_state_dict(oq.state_dict())
    pq = pq.to(_dev())
    x = torch.randn(8, 1000, D, generator=torch.Generator().manual_seed(12))
    eo, qo, io = oq(x)
    ep, qp, ip = pq(x.to(_dev()))
    ip = ip.cpu()
    bad = (ip != io).reshape(-1)
    flat = x.reshape(-1, D)
    w = oq.embedding.weight.detach()
    dist = (w.pow(2).sum(1) - 2 * flat @ w.T + flat.pow(2).sum(1, keepdim=True))
    d_ref = dist.gather(1, io.reshape(-1, 1)).squeeze(1)
    d_ours = dist.gather(1, ip.reshape(-1, 1)).squeeze(1)
    ulp = torch.finfo(torch.float32).eps * d_ref.abs()
    assert ((d_ours - d_ref).abs()[bad] <= 8 * ulp[bad]).all(), "a mismatch that is not a near-tie"
    print(f"fresh-codebook index disagreements (all near-ties): {int(bad.sum())} / {bad.numel()}")
    assert bad.float().mean().item() < 2e-3


def test_vq_codebook_gradient_without_ema():
    from crank_b200.net.module.vqvae2 import Quantizer as PQ
    from oracle.crank_port import Quantizer as OQ

    torch.manual_seed(13)
    K, D = 128, 64
    oq = OQ(D, K, ema_flag=False, bdt_flag=True)
    with torch.no_grad():
        oq.embedding.weight.normal_()
    pq = PQ(D, K, ema_flag=False, bdt_flag=True)
    pq.load_state_dict(oq.state_dict())
    pq = pq.to(_dev())
    x = torch.randn(2, D, 300, generator=torch.Generator().manual_seed(14))
    xo = x.clone().requires_grad_(True)
    xp = x.to(_dev()).requires_grad_(True)
    eo, qo, io = oq(xo)
    ep, qp, ip = pq(xp)
    assert (ip.cpu() == io).all()
    (eo.pow(2).sum() + (qo * 0.5).sum()).backward()
    (ep.pow(2).sum() + (qp * 0.5).sum()).backward()
    assert_close(xp.grad, xo.grad, 1e-6, "straight-through dx")
    assert_close(pq.embedding.weight.grad, oq.embedding.weight.grad, 1e-5, "codebook grad")


# ---- losses -----------------------------------------------------------------------------------
@pytest.mark.parametrize("shift", [0, 3, -2])
@pytest.mark.parametrize("with_mask", [True, False])
def test_masked_l1_mse(shift, with_mask):
    from crank_b200 import ops
    from oracle.crank_port import feature_loss

    g = torch.Generator().manual_seed(20)
    B, T, D = 3, 77, 80
    x = torch.randn(B, T, D, generator=g)
    y = torch.randn(B, T, D, generator=g)
    mask = (torch.rand(B, T, 1, generator=g) < 0.8) if with_mask else None
    xo = x.clone().requires_grad_(True)
    lo1 = feature_loss("l1", xo, y, mask, causal=True, causal_size=shift)
    lo2 = feature_loss("mse", xo, y, mask, causal=True, causal_size=shift)
    (2.0 * lo1 + 0.7 * lo2).backward()
    xp = x.to(_dev()).requires_grad_(True)
    l1, l2 = ops.masked_l1_mse(xp, y.to(_dev()), mask.to(_dev()) if with_mask else None, shift)
    (2.0 * l1 + 0.7 * l2).backward()
    assert_close(l1, lo1, 1e-5, "l1")
    assert_close(l2, lo2, 1e-5, "mse")
    assert_close(xp.grad, xo.grad, 1e-5, "d loss / dx")


def test_masked_mse_against_constant():
    from crank_b200 import ops

    g = torch.Generator().manual_seed(21)
    s = torch.randn(4, 50, 1, generator=g)
    mask = torch.rand(4, 50, 1, generator=g) < 0.7
    so = s.clone().requires_grad_(True)
    sel = so.masked_select(mask)
    lo = torch.nn.functional.mse_loss(sel, torch.ones_like(sel))
    lo.backward()
    sp = s.to(_dev()).requires_grad_(True)
    lp = ops.masked_l1_mse(sp, 1.0, mask.to(_dev()))[1]
    lp.backward()
    assert_close(lp, lo, 1e-6, "lsgan mse")
    assert_close(sp.grad, so.grad, 1e-6, "lsgan mse grad")


@pytest.mark.parametrize("T", [500, 128, 70])
def test_stft_trajectory_loss_recipe_parameters(T):
    from crank_b200.conf import default_conf
    from crank_b200.net.module.loss import CustomFeatureLoss
    from oracle.crank_port import multi_stft_loss

    conf = default_conf()
    g = torch.Generator().manual_seed(22)
    x = torch.randn(2, T, 80, generator=g)
    y = torch.randn(2, T, 80, generator=g)
    xo = x.clone().requires_grad_(True)
    lo = multi_stft_loss(xo, y, conf["stft_params"])
    lo.backward()
    crit = CustomFeatureLoss(loss_type="stft", stft_params=conf["stft_params"], causal=False)
    xp = x.to(_dev()).requires_grad_(True)
    lp = crit(xp, y.to(_dev()))
    lp.backward()
    assert_close(lp, lo, 1e-5, "stft loss")
    assert_close(xp.grad, xo.grad, 1e-4, "stft loss grad")


def test_stft_loss_overlapping_frames_kwargs_path():
    """STFTLoss used directly (test/test_loss.py:28-34 of the reference): fft 32, win 20, hop 10."""
    from crank_b200.net.module.loss import STFTLoss
    from oracle.crank_port import stft_mag

    g = torch.Generator().manual_seed(23)
    x = torch.randn(3, 200, 10, generator=g)
    y = torch.randn(3, 200, 10, generator=g)
    xo = x.clone().requires_grad_(True)
    lo = torch.nn.functional.l1_loss(stft_mag(xo, 32, 10, 20), stft_mag(y, 32, 10, 20))
    lo.backward()
    xp = x.to(_dev()).requires_grad_(True)
    lp = STFTLoss(fft_size=32, win_size=20, hop_size=10)(xp, y.to(_dev()))
    lp.backward()
    assert_close(lp, lo, 1e-5, "stft loss (overlap)")
    assert_close(xp.grad, xo.grad, 1e-4, "stft loss grad (overlap)")


@pytest.mark.parametrize("logratio", [0.3, 1.0])
def test_stft_loss_logratio_forward_and_gradient(logratio):
    """(1 - r) * L1(mag) + r * L1(log mag)  (crank/net/module/loss.py:80-84), value and gradient."""
    from crank_b200.net.module.loss import STFTLoss
    from oracle.crank_port import stft_mag

    g = torch.Generator().manual_seed(29)
    x = torch.randn(2, 150, 12, generator=g)
    y = torch.randn(2, 150, 12, generator=g)
    xo = x.clone().requires_grad_(True)
    xm, ym = stft_mag(xo, 32, 10, 20), stft_mag(y, 32, 10, 20)
    lo = (1 - logratio) * torch.nn.functional.l1_loss(xm, ym) + logratio * torch.nn.functional.l1_loss(xm.log(), ym.log())
    lo.backward()
    xp = x.to(_dev()).requires_grad_(True)
    lp = STFTLoss(fft_size=32, win_size=20, hop_size=10, logratio=logratio)(xp, y.to(_dev()))
    lp.backward()
    assert_close(lp, lo, 1e-5, "stft loss (logratio)")
    # d|log m_x - log m_y|/dx ~ 1/m_x^2: bins with a near-zero magnitude amplify fp32 rounding of the DFT sums
    # (measured: the fp32 oracle itself is ~1e-4 away from a float64 evaluation), so the gradient is judged
    # against float64 with the fp32 oracle's own distance as the yardstick
    xd = x.double().requires_grad_(True)

    def mag64(t):
        z = torch.stft(t.transpose(1, 2).reshape(-1, t.size(1)), 32, 10, 20, torch.hann_window(20, dtype=torch.float64),
                       return_complex=True)
        return torch.sqrt(torch.clamp(z.real ** 2 + z.imag ** 2, min=1e-7).transpose(2, 1))

    xm64, ym64 = mag64(xd), mag64(y.double())
    l64 = (1 - logratio) * torch.nn.functional.l1_loss(xm64, ym64) + logratio * torch.nn.functional.l1_loss(xm64.log(), ym64.log())
    l64.backward()
    e_oracle = rel_err(xo.grad, xd.grad)
    e_ours = rel_err(xp.grad, xd.grad)
    assert e_ours <= 1e-4 + 4 * e_oracle, f"stft loss grad (logratio): {e_ours:.2e} vs float64 (fp32 oracle: {e_oracle:.2e})"


def test_cross_entropy_ignore_index():
    from crank_b200 import ops

    g = torch.Generator().manual_seed(24)
    Fn, S = 999, 14
    logits = torch.randn(Fn, S, generator=g) * 3
    labels = torch.randint(0, S, (Fn,), generator=g)
    labels[torch.rand(Fn, generator=g) < 0.2] = -100
    lo_in = logits.clone().requires_grad_(True)
    lo = torch.nn.functional.cross_entropy(lo_in, labels, ignore_index=-100)
    (lo * 1.5).backward()
    lp_in = logits.to(_dev()).requires_grad_(True)
    lp = ops.cross_entropy(lp_in, labels.to(_dev()))
    (lp * 1.5).backward()
    assert_close(lp, lo, 1e-6, "ce")
    assert_close(lp_in.grad, lo_in.grad, 1e-5, "ce grad")


def test_fused_adam_matches_torch_adam():
    from crank_b200.net.trainer.optim import FusedAdam

    g = torch.Generator().manual_seed(25)
    p0 = torch.randn(10007, generator=g)
    po = torch.nn.Parameter(p0.clone())
    pp = torch.nn.Parameter(p0.clone().to(_dev()))
    oo = torch.optim.Adam([po], lr=2e-4)
    op = FusedAdam([pp], lr=2e-4)
    for _ in range(5):
        gr = torch.randn(10007, generator=g)
        po.grad = gr.clone()
        pp.grad = gr.to(_dev())
        oo.step()
        op.step()
    assert_close(pp, po, 1e-6, "adam params")
    assert (pp.detach().cpu() - po.detach()).abs().max().item() < 1e-6


def test_logmel_matches_oracle_on_reference_wav_fixture():
    """The reference's own KAT: test/data/SF1_10001.wav -> mlfb stored in its feats.h5 (committed copy)."""
    import os

    from crank_b200.net.module.mlfb import LogMelFilterBankLayer
    from oracle import mel as omel

    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_fixture_mlfb.npz"))
    raw, mlfb_ref = fx["raw_i16"].astype(np.float64) / 32768.0, fx["mlfb"]
    # offline convention of crank/feature: symmetric hann, centred reflect padding
    o = omel.logmelfilterbank(raw, 22050, fft_size=1024, hop_size=128, win_length=1024,
                              window=omel.hann(1024, periodic=False), num_mels=80, fmin=80, fmax=7600)
    assert np.abs(o - mlfb_ref).max() < 1e-6
    # online layer (periodic hann, center=False inside VQVAE2; test_feature_pytorch.py uses 1e-3)
    layer = LogMelFilterBankLayer(fs=22050, hop_size=128, fft_size=1024, win_length=1024, window="hann",
                                  center=True, n_mels=80, fmin=80, fmax=7600).to(_dev())
    got = layer(torch.from_numpy(raw).float()[None].to(_dev()))[0].cpu().numpy()
    o_per = omel.logmelfilterbank(raw, 22050, fft_size=1024, hop_size=128, win_length=1024, window="hann",
                                  num_mels=80, fmin=80, fmax=7600)
    assert got.shape == o_per.shape
    # north star: <= 1e-4 relative on mlfb tensors (max |diff| / max |ref|, the convention of tests/util.rel_err).
    # An fp32 STFT pipeline reaches ~1e-5 here (torch CPU fp32 vs the float64 oracle: 1.2e-5; 6.5e-5 max abs).
    err_abs = np.abs(got - o_per).max()
    err_rel = err_abs / np.abs(o_per).max()
    lin_rel = np.abs(10.0 ** got.astype(np.float64) - 10.0 ** o_per).max() / (10.0 ** o_per).max()
    print(f"online log-mel vs float64 oracle: max abs {err_abs:.2e}, rel-to-max {err_rel:.2e}, linear-mel rel {lin_rel:.2e}")
    assert err_rel <= 1e-4, err_rel
    assert err_abs <= 3e-4, err_abs
    assert lin_rel <= 1e-5, lin_rel


def test_offline_mlfb_extraction_matches_reference_fixture():
    """crank_b200.feature.extract_mlfb on the device (symmetric hann, centred) against the mlfb the reference stored
    for its own test wav (float64 offline pipeline; test/test_feature_pytorch.py allows 1e-3 at this boundary)."""
    import os

    from crank_b200.feature import extract_mlfb

    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_fixture_mlfb.npz"))
    raw, ref = fx["raw_i16"].astype(np.float64) / 32768.0, fx["mlfb"]
    got = extract_mlfb(raw, fs=22050, device=_dev()).cpu().numpy()
    assert got.shape == ref.shape
    # fp32 on the device vs the float64 fixture; the wrong (periodic) window would be off by 2.5e-2
    err_abs, err_mean = np.abs(got - ref).max(), np.abs(got - ref).mean()
    err_rel = err_abs / np.abs(ref).max()
    print(f"offline mlfb vs the reference's feats.h5: max abs {err_abs:.2e}, mean {err_mean:.2e}, rel-to-max {err_rel:.2e}")
    assert err_rel <= 1e-4, err_rel          # north star tolerance on mlfb tensors
    assert err_abs <= 3e-4, err_abs
    assert err_mean <= 5e-6, err_mean


@pytest.mark.parametrize("hop,n_frames,B", [(128, 37, 3), (128, 16, 1), (256, 1, 2), (120, 50, 2)])
def test_fused_logmel_matches_cufft_path_and_oracle(hop, n_frames, B):
    """crk_logmel_fused_fwd (one kernel: framing + window + radix-4 FFT-1024 + banded mel + log10 + scaler) against the
    float64 oracle (oracle/mel.py restates mlfb.py:134-171) and against the cuFFT path, on ragged frame counts (odd,
    < 16, not a multiple of the 16 frames a CTA owns) and with the scaler epilogue."""
    from crank_b200 import ops
    from crank_b200.net.module.mlfb import mel_basis
    from oracle import mel as omel

    g = torch.Generator().manual_seed(hop + n_frames)
    n = 1024 + (n_frames - 1) * hop + 7          # 7 trailing samples that belong to no frame
    t = torch.arange(n)[None] / 24000.0
    wav = 0.3 * torch.sin(2 * np.pi * 220.0 * t * (1 + torch.arange(B)[:, None])) + 0.05 * torch.randn(B, n, generator=g)
    basis = torch.from_numpy(mel_basis(24000, 1024, 80, 80, 7600).T.copy()).to(_dev())
    win = torch.hann_window(1024).to(_dev())
    mean, std = torch.randn(80, generator=g).to(_dev()), (0.5 + torch.rand(80, generator=g)).to(_dev())
    for mu, sd in ((None, None), (mean, std)):
        fused = ops.logmel(wav.to(_dev()), win, basis, 1024, hop, mean=mu, std=sd, fused=True)
        plain = ops.logmel(wav.to(_dev()), win, basis, 1024, hop, mean=mu, std=sd, fused=False)
        assert fused.shape == plain.shape == (B, n_frames, 80)
        ob = omel.mel_basis(24000, 1024, 80, 80, 7600).T.astype(np.float64)
        ref = np.stack([np.log10(np.maximum(1e-10, omel.stft_mag(wav[b].double().numpy(), 1024, hop, omel.hann(1024, True),
                                                                 center=False) @ ob)) for b in range(B)])
        if mu is not None:
            ref = (ref - mean.cpu().double().numpy()) / std.cpu().double().numpy()
        e_f = np.abs(fused.cpu().numpy() - ref).max() / np.abs(ref).max()
        e_p = np.abs(plain.cpu().numpy() - ref).max() / np.abs(ref).max()
        print(f"fused log-mel hop {hop} M {n_frames}: rel-to-max vs float64 oracle {e_f:.2e} (cuFFT path {e_p:.2e})")
        assert e_f <= 1e-4, e_f


def _torch_logmel_ref(x, win, basis, hop, mean=None, std=None, eps=1e-10):
    """float64 autograd restatement of mlfb.py:134-171 with center=False: frames * window -> rfft -> |.| -> mel -> log10"""
    fr = x.unfold(-1, 1024, hop) * win
    mag = torch.fft.rfft(fr, dim=-1).abs()
    out = torch.clamp(mag @ basis, min=eps).log10()
    if mean is not None:
        out = (out - mean) / std
    return out


@pytest.mark.parametrize("hop,n_frames,B,scaler", [(128, 37, 2, True), (128, 16, 1, False), (240, 5, 2, False)])
def test_fused_logmel_backward_matches_float64_autograd(hop, n_frames, B, scaler):
    """crk_logmel_fused_bwd: d loss / d window and d loss / d wav of the fused front end against torch float64 autograd
    through the same computation (what the reference's learnable windows get from autograd, mlfb.py:72-110)."""
    from crank_b200 import ops
    from crank_b200.net.module.mlfb import mel_basis

    g = torch.Generator().manual_seed(7 * hop + n_frames)
    n = 1024 + (n_frames - 1) * hop + 3
    t = torch.arange(n)[None] / 24000.0
    wav = 0.3 * torch.sin(2 * np.pi * 310.0 * t * (1 + torch.arange(B)[:, None])) + 0.05 * torch.randn(B, n, generator=g)
    basis = torch.from_numpy(mel_basis(24000, 1024, 80, 80, 7600).T.copy())
    win = torch.hann_window(1024) * (1.0 + 0.1 * torch.randn(1024, generator=g))
    mean = torch.randn(80, generator=g) if scaler else None
    std = (0.5 + torch.rand(80, generator=g)) if scaler else None
    R = torch.randn(B, n_frames, 80, generator=g)
    xr = wav.double().requires_grad_(True)
    wr = win.double().requires_grad_(True)
    ref = _torch_logmel_ref(xr, wr, basis.double(), hop, None if mean is None else mean.double(),
                            None if std is None else std.double())
    (ref * R.double()).sum().backward()
    xp = wav.to(_dev()).requires_grad_(True)
    wp = win.to(_dev()).requires_grad_(True)
    out = ops.logmel_learnable(xp, wp, basis.to(_dev()), 1024, hop, mean=None if mean is None else mean.to(_dev()),
                               std=None if std is None else std.to(_dev()))
    (out * R.to(_dev())).sum().backward()
    e_o = (out.detach().cpu().double() - ref.detach()).abs().max() / ref.detach().abs().max()
    e_w = (wp.grad.cpu().double() - wr.grad).abs().max() / wr.grad.abs().max()
    e_x = (xp.grad.cpu().double() - xr.grad).abs().max() / xr.grad.abs().max()
    print(f"fused log-mel backward hop {hop} M {n_frames}: out {e_o:.2e}  d window {e_w:.2e}  d wav {e_x:.2e}")
    assert e_o <= 1e-4 and e_w <= 1e-4 and e_x <= 1e-4, (e_o, e_w, e_x)


@pytest.mark.parametrize("window", ["param", "conv"])
def test_learnable_stft_windows_train(window):
    """LogMelFilterBankLayer(window="param" / "conv") (mlfb.py:72-96): forward and the gradients of the learnable parameters
    against a float64 torch restatement with the same parameters."""
    from crank_b200.net.module.mlfb import LogMelFilterBankLayer

    torch.manual_seed(3)
    layer = LogMelFilterBankLayer(fs=24000, hop_size=128, fft_size=1024, win_length=1024, window=window, center=False,
                                  n_mels=80, fmin=80, fmax=7600).to(_dev())
    g = torch.Generator().manual_seed(11)
    n = 1024 + 20 * 128
    wav = 0.2 * torch.randn(2, n, generator=g)
    R = torch.randn(2, 21, 80, generator=g)
    tf32_was = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False      # the "conv" pre-filter is a stock cuDNN conv: keep it in fp32 here
    try:
        out = layer(wav.to(_dev()))
        (out * R.to(_dev())).sum().backward()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32_was
    basis = layer.mlfb_layer.mel_basis.detach().cpu().double()
    if window == "param":
        w = layer.stft_layer.window.detach().cpu().double().requires_grad_(True)
        ref = _torch_logmel_ref(wav.double(), w, basis, 128)
        (ref * R.double()).sum().backward()
        pairs = [(layer.stft_layer.window.grad, w.grad)]
    else:
        conv = torch.nn.Conv1d(1, 24, 65, padding=32).double()
        conv.load_state_dict({k: v.detach().cpu().double() for k, v in layer.stft_layer.window_conv[0].state_dict().items()})
        xr = torch.sigmoid(conv(wav.double().unsqueeze(1))).mean(dim=1)
        ref = _torch_logmel_ref(xr, torch.ones(1024, dtype=torch.float64), basis, 128)
        (ref * R.double()).sum().backward()
        pairs = [(layer.stft_layer.window_conv[0].weight.grad, conv.weight.grad),
                 (layer.stft_layer.window_conv[0].bias.grad, conv.bias.grad)]
    e_o = (out.detach().cpu().double() - ref.detach()).abs().max() / ref.detach().abs().max()
    assert e_o <= 1e-4, e_o
    for got, want in pairs:
        assert got is not None
        e = (got.cpu().double() - want).abs().max() / want.abs().max()
        print(f"learnable window {window}: forward {e_o:.2e}, parameter gradient rel err {e:.2e}")
        assert e <= 2e-4, e


@pytest.mark.parametrize("kind", ["radam", "lamb"])
def test_radam_and_lamb_match_the_restated_reference_optimizers(kind):
    """get_optimizer's "radam" / "lamb" (crank/net/trainer/utils.py:44-47) on a weight-normed conv stack: 8 steps (RAdam:
    5 momentum-only steps, then rectified ones) against oracle/optim_port.py on the oracle's per-tensor parameters --
    LAMB's trust ratios must be those of the reference's weight_g / weight_v / bias tensors, not of the flat pack."""
    from crank_b200.net.trainer.optim import FusedLamb, FusedRAdam
    from crank_b200.parallel_wavegan.models import ParallelWaveGANDiscriminator as PD
    from oracle import optim_port as oo
    from oracle.pwg import ParallelWaveGANDiscriminator as OD

    torch.manual_seed(4)
    kw = dict(in_channels=80, out_channels=14, kernel_size=5, layers=4, conv_channels=64)
    o = OD(**kw)
    p = PD(**kw)
    p.load_state_dict(o.state_dict())
    p = p.to(_dev())
    oopt = (oo.RAdam if kind == "radam" else oo.Lamb)(o.parameters(), lr=1e-2)
    popt = (FusedRAdam if kind == "radam" else FusedLamb)(p.parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 80, 200, generator=g)
    for it in range(8):
        tgt = torch.randn(2, 14, 200, generator=g)
        oopt.zero_grad(); popt.zero_grad()
        (o(x) - tgt).square().mean().backward()
        (p(x.to(_dev())) - tgt.to(_dev())).square().mean().backward()
        oopt.step(); popt.step()
    osd, psd = o.state_dict(), p.state_dict()
    worst = 0.0
    for k, v in osd.items():
        e = (psd[k].cpu() - v).abs().max().item() / max(v.abs().max().item(), 1e-12)
        worst = max(worst, e)
    print(f"{kind}: 8 steps, worst parameter rel err {worst:.2e}")
    assert worst <= 2e-4, worst
