"""ctypes binding of libigm_b200.so (C ABI declared in include/igm_b200.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libigm_b200.so")

IGM_MAX_MULTS = 8


class UnetCfg(C.Structure):
    _fields_ = [("dim", C.c_int32), ("channels", C.c_int32), ("n_mults", C.c_int32),
                ("dim_mults", C.c_int32 * IGM_MAX_MULTS), ("height", C.c_int32), ("width", C.c_int32),
                ("max_batch", C.c_int32), ("timesteps", C.c_int32), ("loss_type", C.c_int32),
                ("training", C.c_int32)]


class Schedule(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
        "sqrt_recipm1_alphas_cumprod", "posterior_log_variance_clipped", "posterior_mean_coef1",
        "posterior_mean_coef2")]


class ProfileEntry(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_int64), ("ms", C.c_double), ("flops", C.c_double),
                ("bytes", C.c_double)]


# every symbol include/igm_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "igm_version": (C.c_int, []),
    "igm_last_error": (C.c_char_p, [_P]),
    "igm_unet_create": (C.c_int, [C.POINTER(_P), C.POINTER(UnetCfg), C.c_int]),
    "igm_unet_destroy": (None, [_P]),
    "igm_unet_num_params": (C.c_int, [_P]),
    "igm_unet_param_elems": (C.c_int64, [_P]),
    "igm_unet_param_info": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "igm_unet_bind_params": (C.c_int, [_P, _P, _P]),
    "igm_unet_pack_weights": (C.c_int, [_P, _P]),
    "igm_unet_forward": (C.c_int, [_P, _P, _P, _P, C.c_int, _P]),
    "igm_unet_backward": (C.c_int, [_P, _P, _P, _P]),
    "igm_unet_grad_buckets": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int]),
    "igm_unet_bucket_wait": (C.c_int, [_P, C.c_int, _P]),
    "igm_ddpm_set_schedule": (C.c_int, [_P, C.POINTER(Schedule)]),
    "igm_ddpm_q_sample": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, _P]),
    "igm_ddpm_p_losses": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, _P]),
    "igm_ddpm_p_losses_backward": (C.c_int, [_P, _P, C.c_float, _P]),
    "igm_ddpm_sample_loop": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "igm_grad_axpy": (C.c_int, [_P, _P, _P, _P, C.c_float, C.c_int64, _P]),
    "igm_adam_step": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
                                C.c_int, C.c_float, _P]),
    "igm_debug_read_tap": (C.c_int64, [_P, C.c_char_p, _P, C.c_int64, _P]),
    "igm_vq_workspace_floats": (C.c_int, [C.c_int, C.c_int]),
    "igm_vq_forward": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P]),
    "igm_vq_backward": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_float, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "igm_pixelcnn_weight_floats": (C.c_int64, [C.c_int, C.c_int]),
    "igm_pixelcnn_workspace_floats": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "igm_pixelcnn_run": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, _P]),
    "igm_conv2d_workspace_floats": (C.c_int64, [C.c_int] * 13),
    "igm_conv2d_forward": (C.c_int, [_P, _P, _P, _P, _P] + [C.c_int] * 14 + [_P, _P]),
    "igm_conv2d_backward": (C.c_int, [_P, _P, _P, _P, _P, _P] + [C.c_int] * 14 + [_P, _P]),
    "igm_act_forward": (C.c_int, [C.c_int, _P, _P, C.c_int64, _P, C.c_int64, C.c_int, _P]),
    "igm_act_backward": (C.c_int, [C.c_int, _P, _P, C.c_int64, _P, _P, _P, C.c_int64, C.c_int, _P]),
    "igm_ewise": (C.c_int, [C.c_int, _P, _P, _P, C.c_int64, _P]),
    "igm_mse": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, _P]),
    "igm_ce256": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int, _P]),
    "igm_profile_start": (C.c_int, [_P]),
    "igm_profile_stop": (C.c_int, [_P, C.POINTER(ProfileEntry), C.c_int]),
    "igm_launch_count": (C.c_int64, [_P]),
    "igm_ops_launch_count": (C.c_int64, []),
    "igm_debug_pixelcnn_prof": (C.c_int, [_P]),
    "igm_debug_wgrad": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "igm_debug_resample": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, _P]),
    "igm_debug_conv": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, _P]),
    "igm_debug_conv_bench": (C.c_int, [C.c_int] * 11 + [C.POINTER(C.c_float), _P]),
    "igm_set_conv_engine": (C.c_int, [_P, C.c_int]),
    "igm_get_conv_engine": (C.c_int, [_P]),
}

_lib = None


def load():
    """dlopen the CUDA library; there is deliberately no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python image-generation-models_b200/build.py` "
            "(nvcc, sm_100a). This package has no CPU or eager-PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class EngineError(RuntimeError):
    pass


def check(ctx, rc):
    if rc != 0:
        msg = load().igm_last_error(ctx)
        raise EngineError(f"libigm_b200 error {rc}: {msg.decode() if msg else '?'}")
