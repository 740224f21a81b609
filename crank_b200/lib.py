"""ctypes binding of libcrank_b200.so (the C ABI declared in include/crank_b200.h).

There is no CPU fallback: every op of this package goes through this library, and `lib()`
raises if it is missing or cannot be loaded.
"""

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcrank_b200.so")

vp = C.c_void_p
i32 = C.c_int
i64 = C.c_longlong
f32 = C.c_float


class WavenetCfg(C.Structure):
    _fields_ = [
        ("in_ch", i32), ("out_ch", i32), ("aux_ch", i32), ("layers", i32), ("stacks", i32),
        ("kernel_size", i32), ("causal", i32), ("first_act", i32), ("head_act", i32),
        ("slope", f32),
    ]


class ConvstackCfg(C.Structure):
    _fields_ = [
        ("in_ch", i32), ("out_ch", i32), ("layers", i32), ("kernel_size", i32),
        ("conv_ch", i32), ("dilation_factor", i32), ("slope", f32),
    ]


class ConvDesc(C.Structure):
    _fields_ = [
        ("g_off", i32), ("v_off", i32), ("b_off", i32), ("cout", i32), ("cin", i32), ("k", i32),
        ("w_off", i32), ("bias_off", i32), ("cin_pad", i32), ("ldw", i32), ("perm", i32),
        ("wt_off", i32), ("wt_rows", i32), ("ldwt", i32), ("tc_off", i32), ("tc_kpad", i32), ("tc_n", i32),
        ("tct_off", i32), ("tct_kpad", i32), ("tct_n", i32),
    ]


MAX_CONVS = 64
PW = C.POINTER(WavenetCfg)
PC = C.POINTER(ConvstackCfg)
PD = C.POINTER(ConvDesc)

# name -> (restype, argtypes); must list every symbol of include/crank_b200.h (tests check this)
SIGNATURES = {
    "crk_strerror": (C.c_char_p, [i32]),
    "crk_last_cuda_error": (C.c_char_p, []),
    "crk_version": (i32, []),
    "crk_set_precision": (i32, [i32]),
    "crk_get_precision": (i32, []),
    "crk_debug_tc_disable": (i32, [i32]),
    "crk_debug_opt_disable": (i32, [i32]),
    "crk_debug_opt_enable": (i32, [i32]),
    "crk_debug_timestamps": (i32, [vp, i32, i32]),
    "crk_launch_count": (C.c_ulonglong, []),
    "crk_timing_enable": (i32, [i32]),
    "crk_timing_read": (i32, [C.POINTER(i32), C.POINTER(f32)]),
    "crk_timing_flops": (C.c_double, []),
    "crk_tc_probe": (i32, [vp, i32, i32, vp, i32, i32, vp, i32, i32, i32, i32, i32, vp]),
    "crk_tc_mma_rate": (i32, [i32, i32, i32, i32, i32, vp, vp]),
    "crk_wavenet_describe": (i32, [PW, PD, C.POINTER(i32), C.POINTER(i64), C.POINTER(i64)]),
    "crk_wavenet_act_floats": (i64, [PW, i32, i32]),
    "crk_wavenet_ws_floats": (i64, [PW, i32, i32]),
    "crk_wavenet_weights": (i32, [PW, vp, vp, vp]),
    "crk_wavenet_fwd": (i32, [PW, vp, vp, i32, vp, i32, vp, vp, i32, vp, i32, i32, vp]),
    "crk_wavenet_infer": (i32, [PW, vp, vp, i32, vp, i32, vp, vp, i32, vp, i32, i32, vp]),
    "crk_wavenet_bwd": (i32, [PW, vp, vp, vp, i32, vp, i32, vp, vp, vp, i32, vp, i32, vp, i32, vp, vp, i32, i32, vp]),
    "crk_convstack_describe": (i32, [PC, PD, C.POINTER(i32), C.POINTER(i64), C.POINTER(i64)]),
    "crk_convstack_act_floats": (i64, [PC, i32, i32]),
    "crk_convstack_ws_floats": (i64, [PC, i32, i32]),
    "crk_convstack_weights": (i32, [PC, vp, vp, vp]),
    "crk_convstack_fwd": (i32, [PC, vp, vp, i32, vp, i32, vp, i32, i32, vp]),
    "crk_convstack_bwd": (i32, [PC, vp, vp, vp, i32, vp, vp, i32, vp, i32, f32, vp, vp, i32, i32, vp]),
    "crk_vq_prepare": (i32, [vp, vp, vp, i32, i32, vp]),
    "crk_vq_argmin": (i32, [vp, i32, vp, vp, vp, vp, vp, i32, vp, i32, i64, i32, i32, vp]),
    "crk_vq_tc_blob_floats": (i64, [i32, i32]),
    "crk_vq_pack_tc": (i32, [vp, vp, i32, i32, vp]),
    "crk_vq_argmin_tc": (i32, [vp, i32, vp, vp, vp, vp, vp, i32, vp, i32, i64, i32, i32, vp]),
    "crk_vq_stats_ws_floats": (i64, [i64, i32, i32]),
    "crk_vq_stats": (i32, [vp, i32, vp, vp, vp, vp, i64, i32, i32, vp]),
    "crk_vq_ema": (i32, [vp, vp, vp, vp, vp, f32, f32, i32, i32, vp]),
    "crk_vq_op_floats": (i64, [i32, i32]),
    "crk_vq_pack_op": (i32, [vp, vp, i32, i32, vp]),
    "crk_vq_argmin_fast": (i32, [vp, i32, vp, vp, vp, i32, vp, i32, i64, i32, i32, vp]),
    "crk_vq_stats_floats": (i64, [i32, i32]),
    "crk_vq_stats_fused": (i32, [vp, i32, vp, vp, vp, i64, i32, i32, vp]),
    "crk_vq_ema_fused": (i32, [vp, vp, vp, vp, vp, f32, f32, i32, i32, vp]),
    "crk_vq_scatter_grad": (i32, [vp, i32, vp, vp, i64, i32, i32, vp]),
    "crk_masked_loss_fwd": (i32, [vp, i32, vp, i32, f32, vp, i32, i32, i32, i32, vp, vp, vp]),
    "crk_masked_loss_ws_floats": (i64, [i32, i32, i32]),
    "crk_masked_loss_bwd": (i32, [vp, i32, vp, i32, f32, vp, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp]),
    "crk_stft_loss_ws_floats": (i64, [i32, i32, i32, i32, i32]),
    "crk_stft_loss_fwd": (i32, [vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "crk_stft_loss_bwd": (i32, [vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, f32, vp, i32, i32, vp]),
    "crk_ce_ws_floats": (i64, [i64]),
    "crk_ce_fwd": (i32, [vp, i32, vp, i64, i32, i64, vp, vp, vp]),
    "crk_ce_bwd": (i32, [vp, i32, vp, i64, i32, i64, vp, vp, vp, i32, vp]),
    "crk_adam_step": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, vp]),
    "crk_radam_step": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, vp]),
    "crk_lamb_step": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, vp, f32, f32, f32, f32, vp]),
    "crk_adam_step_dev": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, vp, vp]),
    "crk_logmel_ws_floats": (i64, [i32, i32, i32]),
    "crk_logmel_fwd": (i32, [vp, i32, i64, vp, vp, i32, i32, i32, f32, vp, vp, vp, vp, vp]),
    "crk_logmel_fused_fwd": (i32, [vp, i32, i64, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp, vp, vp, vp]),
    "crk_logmel_bwd_ws_floats": (i64, [i32, i64, i32]),
    "crk_logmel_fused_bwd": (i32, [vp, i32, i64, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, i32, i32, f32, vp, vp, vp,
                                   vp, vp, vp, vp]),
}

PRECISIONS = {"fp32": 0, "tf32x3": 1, "tf32": 2}
_lib = None
LAUNCHES = 0  # number of C-ABI compute calls issued (each launches >= 1 of our kernels)


def lib():
    """The loaded library; raises (loudly) when the CUDA extension is not built / not loadable."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m crank_b200.build` "
                "(crank_b200 has no CPU / PyTorch fallback)"
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
        # default arithmetic of the dense contractions: the 3xTF32 tensor-core parity mode
        mode = os.environ.get("CRANK_B200_PRECISION", "tf32x3")
        if mode not in PRECISIONS:
            raise RuntimeError(f"CRANK_B200_PRECISION={mode!r}: expected one of {sorted(PRECISIONS)}")
        handle.crk_set_precision(PRECISIONS[mode])
        # A/B switch for optional optimisations (see crk_debug_opt_disable in include/crank_b200.h)
        handle.crk_debug_opt_disable(int(os.environ.get("CRANK_B200_OPT_DISABLE", "0")))
        handle.crk_debug_opt_enable(int(os.environ.get("CRANK_B200_OPT_ENABLE", "0")))
    return _lib


class CrkError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        L = lib()
        msg = L.crk_strerror(rc).decode()
        if rc == -2:
            msg += ": " + L.crk_last_cuda_error().decode()
        raise CrkError(f"crank_b200 {what} failed: {msg} (code {rc})")


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """Invoke a compute entry point on the current torch stream and raise on error."""
    global LAUNCHES
    LAUNCHES += 1
    check(getattr(lib(), name)(*args, stream()), name)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise CrkError(
                "crank_b200 ops need CUDA tensors: the hot path is sm_100a kernels only, "
                "there is no CPU fallback"
            )


def describe_wavenet(cfg):
    descs = (ConvDesc * MAX_CONVS)()
    n = i32()
    th = i64()
    we = i64()
    check(lib().crk_wavenet_describe(C.byref(cfg), descs, C.byref(n), C.byref(th), C.byref(we)),
          "crk_wavenet_describe")
    return [descs[i] for i in range(n.value)], th.value, we.value


def describe_convstack(cfg):
    descs = (ConvDesc * MAX_CONVS)()
    n = i32()
    th = i64()
    we = i64()
    check(lib().crk_convstack_describe(C.byref(cfg), descs, C.byref(n), C.byref(th), C.byref(we)),
          "crk_convstack_describe")
    return [descs[i] for i in range(n.value)], th.value, we.value


def set_precision(mode):
    """Arithmetic of the dense conv contractions: "fp32" (CUDA cores), "tf32x3" (tcgen05, error-compensated,
    ~fp32 accuracy -- the parity mode) or "tf32" (tcgen05, fast mode).  Process-wide."""
    check(lib().crk_set_precision(PRECISIONS[mode] if isinstance(mode, str) else int(mode)), "crk_set_precision")


def get_precision():
    inv = {v: k for k, v in PRECISIONS.items()}
    return inv[lib().crk_get_precision()]
