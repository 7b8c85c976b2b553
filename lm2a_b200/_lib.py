"""ctypes binding of liblm2a_b200.so (the C ABI declared in include/lm2a_b200.h).

The library is the product: there is no PyTorch / CPU fallback. Importing this module
without the built library raises; calling an op on a non-sm_100 device raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LM2A_LIB_PATH: load another build of the same ABI (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("LM2A_LIB_PATH") or os.path.join(_HERE, "liblm2a_b200.so")

c_void_p, c_int32, c_int64, c_float = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float


class ConvSeg(ctypes.Structure):
    _fields_ = [("x", c_void_p), ("rows", c_int64), ("ld", c_int32), ("cin", c_int32),
                ("taps", c_int32), ("_pad", c_int32)]


class ConvDesc(ctypes.Structure):
    _fields_ = [("seg", ConvSeg * 2), ("w", c_void_p), ("n_pad", c_int32), ("n_valid", c_int32),
                ("m", c_int64), ("tp", c_int32), ("t_valid", c_int32), ("bias", c_void_p),
                ("film", c_void_p), ("film_ld", c_int32), ("film_shift_off", c_int32),
                ("residual", c_void_p), ("res_ld", c_int32), ("out_mode", c_int32),
                ("out", c_void_p), ("out_ld", c_int32), ("block_n", c_int32),
                ("stats", c_void_p), ("stats_pitch", c_int32), ("stats_cg", c_int32),
                ("cta_group", c_int32), ("stats_c0", c_int32),
                ("in_gn_stats", c_void_p), ("in_gn_gamma", c_void_p), ("in_gn_beta", c_void_p),
                ("in_gn_pitch", c_int32), ("in_gn_groups", c_int32), ("in_gn_eps", c_float),
                ("in_gn_silu", c_int32), ("in_up_tp", c_int32), ("in_up_t", c_int32),
                ("k_order", c_int32), ("_pad0", c_int32)]


ABI_VERSION = 11
TAPS_K1, TAPS_K3, TAPS_K4S2 = 0, 1, 2
OUT_BF16_SLAB, OUT_F32_NCT = 0, 1

# name -> (restype, argtypes); must list every symbol include/lm2a_b200.h declares
SIGNATURES = {
    "lm2a_abi_version": (c_int32, []),
    "lm2a_last_error": (ctypes.c_char_p, []),
    "lm2a_check_device": (c_int32, []),
    "lm2a_launch_count": (c_int64, []),
    "lm2a_reset_launch_count": (None, []),
    "lm2a_conv1d_bf16": (c_int32, [c_void_p, ctypes.POINTER(ConvDesc)]),
    "lm2a_conv1d_f32": (c_int32, [c_void_p, ctypes.POINTER(ConvDesc)]),
    "lm2a_cross_attn_f32": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p,
                                      c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32,
                                      c_int32, c_int32, c_int32, c_int32, c_int32]),
    "lm2a_bias_add_f32": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p,
                                    c_int64, c_int32, c_int32, c_int32, c_void_p, c_int32, c_int32,
                                    c_int32]),
    "lm2a_ingest_x_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                    c_int32, c_int32, c_int32, c_void_p, c_int64]),
    "lm2a_ingest_seq_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                      c_int32, c_int32]),
    "lm2a_gn_silu_bf16": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p,
                                    c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                    c_float, c_int32]),
    "lm2a_cross_attn_bf16": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p,
                                       c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                       c_int32]),
    "lm2a_cross_attn_streams_bf16": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32,
                                               c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                               c_int32, c_void_p, c_int32, c_int32, c_int32,
                                               c_int32, c_int32, c_int32, c_int32, c_int32]),
    "lm2a_cross_attn_cond_supported": (c_int32, [c_int32]),
    "lm2a_cross_attn_cond_bf16": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32,
                                            c_void_p, c_void_p, c_int32, c_void_p, c_int32,
                                            c_int32, c_int32, c_int32, c_int32, c_int32, c_int32]),
    "lm2a_cross_attn_tail_bf16": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32,
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                            c_int32, c_void_p, c_int32, c_int32, c_int32,
                                            c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                            c_int32]),
    "lm2a_transpose_kv_bf16": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int32,
                                         c_int32, c_int32]),
    "lm2a_time_mlp": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                c_int32]),
    "lm2a_time_embed": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                  c_int32, c_int32]),
    "lm2a_film": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                            c_int32]),
    "lm2a_ingest_x": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                c_int32, c_int32, c_int32, c_void_p, c_int64]),
    "lm2a_ingest_seq": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                  c_int32, c_int32]),
    "lm2a_resample_seq": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                    c_int32, c_int32, c_int32, c_int32, c_int32]),
    "lm2a_upsample2x_bf16": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int32,
                                       c_int32, c_int32, c_int32, c_int32]),
    "lm2a_bias_add_bf16": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p,
                                     c_int64, c_int32, c_int32, c_int32, c_void_p, c_int32,
                                     c_int32, c_int32]),
    "lm2a_gn_apply_bf16": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p,
                                     c_int32, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                     c_int32, c_int32, c_float, c_int32]),
    "lm2a_cfg_posterior": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int32, c_void_p, c_int32, c_int64, c_float, c_int32,
                                     c_int32, c_void_p]),
    "lm2a_cfg_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_float,
                                c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32, c_void_p,
                                c_int64, c_void_p]),
    "lm2a_philox_normal": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                     ctypes.c_uint32]),
    "lm2a_cfg_ddim": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int64, c_float,
                                c_int32, c_int32, c_void_p]),
    "lm2a_mel_metrics": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                   c_int32, c_float, c_float]),
    "lm2a_adan_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                 c_int32, c_void_p]),
}

_lib = None


def load():
    """Load the shared library (building it is `__graft_entry__.build()`'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension is required (no fallback). "
            "Run `python -m lm2a_b200.build` (needs nvcc).")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.lm2a_abi_version() != ABI_VERSION:
        raise RuntimeError("liblm2a_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().lm2a_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed ({status}): {msg}")
