"""Shared helpers for the parity tests (oracle = CPU restatement, product = CUDA kernels)."""
import torch


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = b.abs().max().clamp_min(1e-30)
    return ((a - b).abs().max() / den).item()


def assert_close(a, b, tol, what=""):
    e = rel_err(a, b)
    assert e <= tol, f"{what}: max-abs-diff / max-abs-ref = {e:.3e} > {tol:.1e}"
    return e


def compare_conv_grads(prod, oracle, tol, what=""):
    """prod: _PackedConvNet after backward; oracle: the restated module after backward."""
    pg = prod.named_conv_grads()
    worst = 0.0
    for k, p in oracle.named_parameters():
        assert k in pg, f"{what}: missing grad for {k}"
        if p.grad is None:
            continue
        e = rel_err(pg[k], p.grad)
        # a gradient that is identically ~0 in the oracle (e.g. last block's conv1x1_out) is compared absolutely
        if p.grad.abs().max() < 1e-12:
            assert pg[k].abs().max().item() < 1e-6, f"{what}: {k} should be zero"
            continue
        worst = max(worst, e)
        assert e <= tol, f"{what}: grad {k}: rel err {e:.3e} > {tol:.1e}"
    return worst
