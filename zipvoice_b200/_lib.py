"""ctypes binding of the C ABI in include/zipvoice_b200.h (libzipvoice_b200.so, built in-tree by
`__graft_entry__.build()` / `zipvoice_b200.build`).  There is no fallback: if the shared library
is missing or no sm_100 device is present, every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# ZVB_LIB: an alternative in-tree build of the same sources (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("ZVB_LIB") or os.path.join(HERE, "libzipvoice_b200.so")

ZVB_ABI_VERSION = 6
ZVB_MAX_STACKS = 8
ZVB_VOC_MAX_LAYERS = 16


class ZvbError(RuntimeError):
    pass


class zvb_linear(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p), ("out_features", C.c_int32),
                ("in_features", C.c_int32), ("k_pitch", C.c_int32), ("rows", C.c_int32)]


class zvb_layer(C.Structure):
    _fields_ = [("attn_in", zvb_linear), ("pos_table", C.c_void_p), ("pos_table_tc", C.c_void_p),
                ("ff_in", zvb_linear * 3), ("ff_out", zvb_linear * 3), ("ff1_attn", zvb_linear),
                ("na_sx", zvb_linear), ("na_y", zvb_linear), ("na_out", zvb_linear),
                ("sa_in", zvb_linear * 2), ("sa_out", zvb_linear * 2),
                ("conv_in", zvb_linear * 2), ("dw_w", C.c_void_p * 2), ("dw_b", C.c_void_p * 2),
                ("conv_out", zvb_linear * 2),
                ("norm_bias", C.c_void_p), ("norm_log_scale", C.c_void_p),
                ("bypass_scale", C.c_void_p), ("bypass_mid_scale", C.c_void_p)]


class zvb_stack(C.Structure):
    _fields_ = [("downsample", C.c_int32), ("num_layers", C.c_int32), ("conv_kernel", C.c_int32),
                ("first_layer", C.c_int32), ("ds_weights", C.c_float * 4),
                ("out_combiner_scale", C.c_void_p), ("time_w", C.c_void_p), ("time_b", C.c_void_p)]


class zvb_model(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("dim", C.c_int32), ("num_heads", C.c_int32),
                ("value_head_dim", C.c_int32), ("in_dim", C.c_int32), ("out_dim", C.c_int32),
                ("ff_dims", C.c_int32 * 3), ("na_hidden", C.c_int32), ("time_dim", C.c_int32),
                ("use_guidance_embed", C.c_int32), ("guidance_dim", C.c_int32), ("num_stacks", C.c_int32),
                ("num_layers", C.c_int32),
                ("in_proj", zvb_linear), ("out_proj", zvb_linear),
                ("time0_w", C.c_void_p), ("time0_b", C.c_void_p),
                ("time2_w", C.c_void_p), ("time2_b", C.c_void_p), ("guidance_w", C.c_void_p),
                ("stacks", zvb_stack * ZVB_MAX_STACKS), ("layers", C.POINTER(zvb_layer))]


class zvb_voc_layer(C.Structure):
    _fields_ = [("dw_w", C.c_void_p), ("dw_b", C.c_void_p), ("ln_w", C.c_void_p), ("ln_b", C.c_void_p),
                ("pw1", zvb_linear), ("pw2", zvb_linear)]


class zvb_vocoder(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("dim", C.c_int32), ("intermediate", C.c_int32), ("n_layers", C.c_int32),
                ("n_mels", C.c_int32), ("n_fft", C.c_int32), ("hop", C.c_int32), ("kernel", C.c_int32),
                ("embed", zvb_linear), ("norm_w", C.c_void_p), ("norm_b", C.c_void_p),
                ("layers", zvb_voc_layer * ZVB_VOC_MAX_LAYERS), ("final_w", C.c_void_p), ("final_b", C.c_void_p),
                ("head", zvb_linear), ("window", C.c_void_p)]


class zvb_io(C.Structure):
    _fields_ = [("xin", C.c_void_p), ("t", C.c_void_p), ("g", C.c_void_p), ("mask", C.c_void_p),
                ("out", C.c_void_p), ("xin_pitch", C.c_int32)]


EXPORTS = {
    "zvb_last_error": (C.c_char_p, []),
    "zvb_abi_version": (C.c_int, []),
    "zvb_source_hash": (C.c_char_p, []),
    "zvb_plan_set_saturation_counter": (C.c_int, [C.c_void_p, C.c_void_p]),
    "zvb_launch_count": (C.c_longlong, []),
    "zvb_plan_workspace_bytes": (C.c_int, [C.POINTER(zvb_model), C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "zvb_plan_create": (C.c_int, [C.POINTER(zvb_model), C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                  C.POINTER(C.c_void_p)]),
    "zvb_plan_destroy": (None, [C.c_void_p]),
    "zvb_plan_io": (C.c_int, [C.c_void_p, C.POINTER(zvb_io)]),
    "zvb_decoder_forward": (C.c_int, [C.c_void_p, C.c_void_p]),
    "zvb_decoder_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "zvb_decoder_forward_f32": (C.c_int, [C.c_void_p] * 7),
    "zvb_sample": (C.c_int, [C.c_void_p] * 8 + [C.c_int] * 5 + [C.c_void_p, C.c_void_p]),
    "zvb_fbank": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                            C.c_float, C.c_void_p, C.c_int, C.c_void_p]),
    "zvb_vocoder_workspace_bytes": (C.c_int, [C.POINTER(zvb_vocoder), C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "zvb_vocoder_create": (C.c_int, [C.POINTER(zvb_vocoder), C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_void_p)]),
    "zvb_vocoder_destroy": (None, [C.c_void_p]),
    "zvb_vocoder_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p]),
    "zvb_vocoder_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_int)]),
    "zvb_test_dwconv_linear": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_void_p]),
    "zvb_test_layernorm": (C.c_int, [C.c_void_p] * 5 + [C.c_longlong, C.c_int, C.c_float, C.c_void_p]),
    "zvb_test_linear_masked": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                         C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "zvb_test_istft": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p]),
    "zvb_test_linear": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_int, C.c_void_p]),
    "zvb_test_attn_weights": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "zvb_test_attn_weights_tc": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "zvb_test_pv": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 7 + [C.c_void_p, C.c_void_p]),
    "zvb_test_gated": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "zvb_test_biasnorm_bypass": (C.c_int, [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 3 +
                                 [C.c_longlong, C.c_int, C.c_void_p]),
    "zvb_test_dwconv": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_void_p]),
    "zvb_test_cfg_euler": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int,
                                     C.c_int, C.c_longlong, C.c_int, C.c_void_p]),
    "zvb_test_linear_t": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "zvb_test_downsample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int,
                                      C.c_void_p]),
    "zvb_test_upsample_combine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_void_p]),
    "zvb_test_stream_prep": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_void_p]),
    "zvb_test_assemble_input": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 7 + [C.c_void_p]),
    "zvb_test_small_linear": (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 5 + [C.c_void_p]),
    "zvb_test_timestep_embedding": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "zvb_test_masks": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "zvb_debug_launch_shape": (C.c_int, [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int,
                                         C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
}

_lib = None


def load():
    """Load libzipvoice_b200.so and bind every symbol the header declares (raises if absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZvbError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for the B200 path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.zvb_abi_version() != ZVB_ABI_VERSION:
        raise ZvbError("libzipvoice_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


CATEGORIES = ["gemm_linear", "gemm_gated", "gemm_pv", "attn_weights", "biasnorm_bypass", "elementwise",
              "resample", "dwconv_swooshr", "other"]


def check(status: int):
    if status != 0:
        raise ZvbError(f"zvb error {status}: {load().zvb_last_error().decode()}")


def launch_count() -> int:
    return int(load().zvb_launch_count())
