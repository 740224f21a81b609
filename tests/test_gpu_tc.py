"""tcgen05 probe: pins the tensor-core kernels' operand layout / descriptors / TMEM path on hardware."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _probe(A, B, N, K, row_shift, mode, split):
    from crank_b200 import lib as L

    D = torch.full((128, N), float("nan"), device="cuda")
    L.call("crk_tc_probe", L.ptr(A), A.stride(0), A.shape[0], L.ptr(B), B.stride(0), B.shape[0], L.ptr(D),
           N, K, row_shift, mode, split)
    torch.cuda.synchronize()
    return D


@pytest.mark.parametrize("split", [1, 0])
@pytest.mark.parametrize("N,K,shift", [(128, 64, 0), (128, 64, 3), (64, 128, 5), (128, 8, 1)])
def test_tc_probe_k_major_with_row_shift(N, K, shift, split):
    g = torch.Generator().manual_seed(N + K + shift)
    A = torch.randn(128 + 8, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    D = _probe(A, B, N, K, shift, 0, split)
    ref = (A[shift : shift + 128].double() @ B.double().T).float()
    assert not torch.isnan(D).any(), "tcgen05 pipeline did not complete (mbarrier timeout)"
    err = ((D - ref).abs().max() / ref.abs().max()).item()
    assert err < (2e-6 if split else 3e-3), err


@pytest.mark.parametrize("split", [1, 0])
@pytest.mark.parametrize("N,frames", [(64, 128), (128, 64), (64, 8)])
def test_tc_probe_mn_major_wgrad_style(N, frames, split):
    g = torch.Generator().manual_seed(N + frames)
    A = torch.randn(frames, 128, generator=g).cuda()
    B = torch.randn(frames, N, generator=g).cuda()
    D = _probe(A, B, N, frames, 0, 1, split)
    ref = (A.double().T @ B.double()).float()
    assert not torch.isnan(D).any(), "tcgen05 pipeline did not complete (mbarrier timeout)"
    err = ((D - ref).abs().max() / ref.abs().max()).item()
    assert err < (2e-6 if split else 3e-3), err
