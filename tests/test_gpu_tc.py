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
@pytest.mark.parametrize("N,K,shift", [(128, 64, 0), (128, 64, 3), (64, 64, 5), (64, 128, 0), (128, 8, 1)])
def test_tc_probe_k_major_with_row_shift(N, K, shift, split):
    g = torch.Generator().manual_seed(N + K + shift)
    if K == 128 and split:
        pytest.skip("probe tile does not fit shared memory with the hi/lo split at K=128")
    A = torch.randn(128 + 8, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    D = _probe(A, B, N, K, shift, 0, split)
    ref = (A[shift : shift + 128].double() @ B.double().T).float()
    assert not torch.isnan(D).any(), "tcgen05 pipeline did not complete (mbarrier timeout)"
    err = ((D - ref).abs().max() / ref.abs().max()).item()
    assert err < (2e-6 if split else 3e-3), err


def test_tc_probe_mn_major_tf32_is_not_usable_without_swizzle():
    """Measured on B200: MN-major tf32 operands in the no-swizzle (INTERLEAVE) layout produce zeros.
    The library therefore never uses them (wgrad stages both operands transposed, K-major); this test
    documents the finding and will flag it if a driver/toolkit change makes the combination work."""
    g = torch.Generator().manual_seed(5)
    A = torch.randn(64, 128, generator=g).cuda()
    B = torch.randn(64, 128, generator=g).cuda()
    D = _probe(A, B, 128, 64, 0, 1, 0)
    ref = (A.double().T @ B.double()).float()
    err = ((D - ref).abs().max() / ref.abs().max()).item()
    assert err > 0.5, "MN-major tf32 INTERLEAVE operands now work: wgrad could drop its transposed staging"


@pytest.fixture
def precision():
    from crank_b200 import lib as L

    def _set(mode):
        L.set_precision(mode)

    yield _set
    L.set_precision("tf32x3")


@pytest.mark.parametrize("mode,tol", [("tf32x3", 1e-4), ("tf32", 2e-2)])
@pytest.mark.parametrize(
    "in_ch,out_ch,aux,k,layers,stacks,B,T",
    [
        (80, 64, 0, 5, 8, 4, 3, 300),     # encoder 0: k5, dilations 1,2
        (64, 64, 0, 3, 6, 3, 2, 130),     # encoder 1 / decoder 1
        (128, 80, 34, 5, 8, 4, 2, 200),   # decoder 0 with aux conditioning
        (80, 64, 0, 3, 2, 1, 1, 17),      # T smaller than a tile
    ],
)
def test_tensor_core_forward_matches_oracle(precision, mode, tol, in_ch, out_ch, aux, k, layers, stacks, B, T):
    from tests.test_gpu_kernels import _pair_generator
    from tests.util import rel_err

    o, p = _pair_generator(in_ch, out_ch, aux, k, layers, stacks, False)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, in_ch, T, generator=g)
    c = torch.randn(B, aux, T, generator=g) if aux > 0 else None
    with torch.no_grad():
        yo = o(x, c)
    precision(mode)
    with torch.no_grad():
        yp = p(x.cuda(), c.cuda() if c is not None else None)
    assert not torch.isnan(yp).any(), "tcgen05 pipeline timed out (poisoned output)"
    e = rel_err(yp, yo)
    print(f"{mode}: forward rel err {e:.2e}")
    assert e <= tol, e


def test_tensor_core_forward_backward_through_saved_gates(precision):
    """forward on tcgen05 (3xTF32) + fp32 backward kernels: gradients must still match the oracle."""
    from tests.test_gpu_kernels import _pair_generator
    from tests.util import assert_close, compare_conv_grads

    o, p = _pair_generator(128, 80, 34, 5, 8, 4, False)
    g = torch.Generator().manual_seed(2)
    B, T = 2, 150
    x = torch.randn(B, 128, T, generator=g)
    c = torch.randn(B, 34, T, generator=g)
    xo, co = x.clone().requires_grad_(True), c.clone().requires_grad_(True)
    yo = o(xo, co)
    precision("tf32x3")
    xp, cp = x.cuda().requires_grad_(True), c.cuda().requires_grad_(True)
    yp = p(xp, cp)
    dy = torch.randn(yo.shape, generator=g)
    yo.backward(dy)
    yp.backward(dy.cuda())
    assert_close(yp, yo, 1e-4, "fwd")
    assert_close(xp.grad, xo.grad, 1e-4, "dx")
    assert_close(cp.grad, co.grad, 1e-4, "dc")
    compare_conv_grads(p, o, 1e-4, "stack (tc fwd)")


# ---- the whole kernel-parity suite again with every dense contraction on the tensor cores -------------
_STACKS = [
    (80, 64, 0, 5, 8, 4, False, 3, 100),
    (64, 64, 0, 3, 6, 3, False, 2, 130),
    (128, 80, 34, 5, 8, 4, False, 2, 75),
    (80, 64, 2, 5, 4, 2, True, 2, 70),
    (80, 64, 0, 3, 2, 1, False, 1, 17),
    (80, 64, 0, 5, 8, 4, False, 2, 500),
]


@pytest.mark.parametrize("in_ch,out_ch,aux,k,layers,stacks,causal,B,T", _STACKS)
def test_tc_wavenet_stack_fwd_bwd_3xtf32(precision, in_ch, out_ch, aux, k, layers, stacks, causal, B, T):
    from tests import test_gpu_kernels as tk

    precision("tf32x3")
    tk.test_wavenet_stack_fwd_bwd(in_ch, out_ch, aux, k, layers, stacks, causal, B, T)


def test_tc_residual_discriminator_3xtf32(precision):
    from tests import test_gpu_kernels as tk

    precision("tf32x3")
    tk.test_residual_discriminator_fwd_bwd_with_injected_dropout()


@pytest.mark.parametrize("in_ch,out_ch,k,layers,B,T", [(80, 14, 5, 8, 2, 150), (128, 14, 3, 3, 3, 70), (80, 12, 5, 8, 1, 64)])
def test_tc_convstack_3xtf32(precision, in_ch, out_ch, k, layers, B, T):
    from tests import test_gpu_kernels as tk

    precision("tf32x3")
    tk.test_convstack_fwd_bwd(in_ch, out_ch, k, layers, B, T)


@pytest.mark.parametrize("kind", ["vqvae", "lsgan", "cyclegan", "stargan"])
def test_tc_train_steps_3xtf32(precision, kind):
    from tests import test_gpu_trainstep as tt

    precision("tf32x3")
    # losses: the same 1e-4 as the fp32 kernels.  Parameter movement after two Adam steps: Adam's first updates are
    # ~lr * sign(g), so the elements whose gradient is cancellation noise move by O(lr) in either direction; measured
    # worst RMS movement error 2.1e-3 (fp32 kernels) / 3.0e-2 (3xTF32, vqvae encoders.0 first block), bound 5e-2.
    # The pre-Adam gradients are compared directly in tests/test_gpu_baseline_shapes.py.
    tt.test_train_steps_match_oracle_and_golden(kind, rms_tol=5e-2)


def test_tc_fast_mode_tf32_is_close(precision):
    """plain TF32 (fast mode): not the parity mode, but must stay within ~1e-2 of the oracle end to end."""
    from tests.test_gpu_kernels import _pair_generator
    from tests.util import rel_err

    o, p = _pair_generator(80, 64, 0, 5, 8, 4, False)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 80, 200, generator=g)
    xo = x.clone().requires_grad_(True)
    yo = o(xo, None)
    precision("tf32")
    xp = x.cuda().requires_grad_(True)
    yp = p(xp, None)
    dy = torch.randn(yo.shape, generator=g)
    yo.backward(dy)
    yp.backward(dy.cuda())
    assert rel_err(yp, yo) < 2e-2 and rel_err(xp.grad, xo.grad) < 0.3
    pg = p.named_conv_grads()
    worst = max(rel_err(pg[k], q.grad) for k, q in o.named_parameters() if q.grad is not None and q.grad.abs().max() > 1e-12)
    print(f"tf32 fast mode: fwd {rel_err(yp, yo):.1e} dx {rel_err(xp.grad, xo.grad):.1e} worst param grad {worst:.1e}")
    assert worst < 0.3


def test_tc_vq_argmin_is_identical_to_fp32_kernel(precision):
    """tensor-core argmin + exact re-score must return the very indices of the fp32 kernel (and e, qx bit-equal), on a
    spread codebook, on the degenerate fresh-init codebook (all near-ties) and with dead codes (|w|~1e5): both the round-2
    kernel (one TF32 pass, resident codebook, persistent CTAs: crk_vq_argmin_fast) and round 1's 3xTF32 kernel."""
    from crank_b200 import ops

    g = torch.Generator().manual_seed(31)
    for B, T, K in ((16, 500, 512), (64, 500, 512), (3, 333, 512), (8, 500, 256), (8, 500, 384), (8, 500, 128)):
        x = torch.randn(B, T, 64, generator=g).cuda()
        for kind in ("spread", "fresh", "dead"):
            W = torch.randn(K, 64, generator=g)
            if kind == "fresh":
                W = (torch.rand(K, 64, generator=g) * 2 - 1) / K
            if kind == "dead":
                W[K // 5:K // 2] *= 1e5
            W = W.cuda()
            precision("fp32")
            e0, q0, i0 = ops.VQFn.apply(x, W)
            precision("tf32x3")
            for variant in (None, "tf32x3"):
                e1, q1, i1 = ops.VQFn.apply(x, W, None, variant)
                assert (i1 >= 0).all(), "tcgen05 pipeline timed out"
                nm = int((i0 != i1).sum())
                assert nm == 0, f"{kind} {B}x{T} K={K} variant {variant}: {nm} indices differ between the fp32 and the tensor-core argmin"
                assert torch.equal(e0, e1) and torch.equal(q0, q1)


def test_fused_ema_matches_round1_kernels(precision):
    """crk_vq_stats_fused + crk_vq_ema_fused (one EMA launch, operand blob refreshed in place) against round 1's
    crk_vq_stats + crk_vq_ema + crk_vq_pack_op: EMA buffers and codebook bit-identical, blob bit-identical to a fresh pack."""
    import ctypes as C

    from crank_b200 import lib as L
    from crank_b200 import ops

    precision("tf32x3")
    g = torch.Generator().manual_seed(5)
    K, D = 512, 64
    x = torch.randn(8, 500, D, generator=g).cuda()
    W = torch.randn(K, D, generator=g).cuda()
    size = torch.rand(K, generator=g).cuda() * 10
    emaw = torch.randn(D, K, generator=g).cuda()
    _, _, idx = ops.VQFn.apply(x, W)
    # round 1 sequence
    W1, s1, w1 = W.clone(), size.clone(), emaw.clone()
    stats = torch.empty(K + D * K, device="cuda")
    ws = torch.empty(L.lib().crk_vq_stats_ws_floats(x.shape[0] * x.shape[1], K, D), device="cuda")
    L.call("crk_vq_stats", L.ptr(x), D, L.ptr(idx), L.ptr(stats[:K]), L.ptr(stats[K:]), L.ptr(ws), x.shape[0] * x.shape[1], K, D)
    L.call("crk_vq_ema", L.ptr(stats[:K]), L.ptr(stats[K:]), L.ptr(s1), L.ptr(w1), L.ptr(W1), C.c_float(0.99), C.c_float(1e-5), K, D)
    # round 2
    W2, s2, w2 = W.clone(), size.clone(), emaw.clone()
    blob = ops.vq_pack_operand(W2)
    for _ in range(2):      # twice: the ticket counter must reset itself
        W2.copy_(W); s2.copy_(size); w2.copy_(emaw)
        ops.vq_ema_update(x, idx, s2, w2, W2, 0.99, 1e-5, opblob=blob)
        assert torch.equal(s1, s2), f"ema_size differs: {(s1 - s2).abs().max().item():.3e}"
        assert torch.equal(w1, w2), f"ema_w differs: {(w1 - w2).abs().max().item():.3e}"
        assert torch.equal(W1, W2), f"codebook differs: {(W1 - W2).abs().max().item():.3e}"
        assert torch.equal(blob, ops.vq_pack_operand(W2)), "operand blob differs from a fresh pack"


@pytest.mark.parametrize("in_ch,out_ch,aux,k,layers,stacks,causal,B,T", [_STACKS[0], _STACKS[2], _STACKS[3], _STACKS[5]])
def test_persistent_pipelined_kernels_opt_in_path(precision, in_ch, out_ch, aux, k, layers, stacks, causal, B, T):
    """The opt-in persistent warp-specialised kernels (k_resblock_fwd_pt: overlapped GEMM / gate / epilogue with the aux
    operand in tensor memory; k_conv_pt: segment-pipelined dgrad with 16x256b epilogues) against the oracle: forward,
    input / conditioning gradients and every parameter gradient at the same 1e-4 as the default kernels."""
    from crank_b200 import lib as L
    from tests import test_gpu_kernels as tk

    L.check(L.lib().crk_debug_opt_enable(3), "opt_enable")
    try:
        precision("tf32x3")
        tk.test_wavenet_stack_fwd_bwd(in_ch, out_ch, aux, k, layers, stacks, causal, B, T)
    finally:
        L.check(L.lib().crk_debug_opt_enable(0), "opt_enable")
