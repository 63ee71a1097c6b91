"""ctypes binding of libsdb200.so (C ABI: include/sdb200.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsdb200.so")

c_void_p, c_int, c_ll, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("kind", c_int), ("a0", c_void_p), ("a1", c_void_p), ("w", c_void_p), ("bias", c_void_p),
        ("residual", c_void_p), ("out", c_void_p), ("out2", c_void_p), ("workspace", c_void_p), ("M", c_int),
        ("NB", c_int), ("HI", c_int), ("WI", c_int), ("C0", c_int), ("C1", c_int), ("Cout", c_int),
        ("lda0", c_ll), ("lda1", c_ll), ("ldw", c_ll), ("ldo", c_ll), ("ldr", c_ll),
        ("out_fp32", c_int), ("res_fp32", c_int), ("bias_per_row", c_int), ("act", c_int), ("block_n", c_int),
        ("nsplit", c_int), ("smem_budget", c_int), ("cta_pair", c_int), ("out_f16", c_int),
        ("epi_mode", c_int), ("gn_part", c_void_p), ("gn_hw", c_int),
        ("ax0", c_void_p), ("ax1", c_void_p), ("Cx0", c_int), ("Cx1", c_int), ("up_phase", c_int),
        ("ab_f16", c_int),
    ]


class AttnArgs(ctypes.Structure):
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("vt", c_void_p), ("out", c_void_p), ("NB", c_int),
        ("heads", c_int), ("d", c_int), ("S", c_int), ("Skv", c_int), ("Skv_pad", c_int), ("vt_ld", c_int),
        ("ldq", c_ll), ("ldk", c_ll), ("ldo", c_ll), ("causal", c_int), ("scale", c_float), ("variant", c_int), ("sum_row", c_int), ("p_f16", c_int),
        ("exp_poly", c_int), ("q_prescaled", c_int), ("qk_cols", c_int), ("qk_fold", c_int),
    ]


# name -> argtypes (every function returns int unless listed in _RESTYPES)
SIGNATURES = {
    "sdb_abi_version": [],
    "sdb_last_error": [],
    "sdb_launch_count": [],
    "sdb_args_size": [c_int],
    "sdb_read_fault": [ctypes.POINTER(ctypes.c_uint)],
    "sdb_gemm_tc": [ctypes.POINTER(GemmArgs), c_void_p],
    "sdb_debug_gemm_trace": [c_int, ctypes.POINTER(ctypes.c_longlong)],
    "sdb_attention": [ctypes.POINTER(AttnArgs), c_void_p],
    "sdb_groupnorm_stats_bytes": [c_int, c_int],
    "sdb_groupnorm_stats": [c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_int, c_int, c_int, c_int,
                            c_void_p],
    "sdb_groupnorm_apply": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_ll,
                            c_int, c_int, c_int, c_float, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "sdb_groupnorm_reduce_partials": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_void_p],
    "sdb_gemm_gn_slabs": [c_int, c_int, c_int, c_int, c_int, c_int],
    "sdb_gemm_conv_a3_bytes": [c_int, c_int, c_int],
    "sdb_groupnorm_fused_supported": [c_ll, c_int, c_int, c_int],
    "sdb_groupnorm_fused": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_int, c_int,
                            c_float, c_int, c_int, c_void_p],
    "sdb_layernorm": [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_float, c_int, c_int, c_void_p],
    "sdb_softmax_rows": [c_void_p, c_void_p, c_ll, c_int, c_float, c_void_p],
    "sdb_fill_zero": [c_void_p, c_ll, c_void_p],
    "sdb_nchw_f32_to_nhwc": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p],
    "sdb_nhwc_to_nchw_f32": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "sdb_upsample2x_nhwc": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "sdb_conv_direct": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                        c_int, c_int, c_int, c_int, c_void_p],
    "sdb_small_linear": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "sdb_cfg_ddpm_step": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_int, c_void_p, c_int,
                          c_int, c_int, c_int, c_int, c_int, c_void_p],
    "sdb_vae_attn_scramble_add": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_void_p],
    "sdb_f32_to_bf16": [c_void_p, c_void_p, c_ll, c_int, c_void_p],
    "sdb_vae_encode_tail": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "sdb_axpby": [c_void_p, c_void_p, c_void_p, c_float, c_float, c_ll, c_void_p],
    "sdb_image_to_uint8": [c_void_p, c_void_p, c_ll, c_void_p],
    "sdb_uint8_to_image": [c_void_p, c_void_p, c_ll, c_int, c_void_p],
    "sdb_clip_embed": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "sdb_resample_u8": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                        c_void_p],
    "sdb_copy_bytes": [c_void_p, c_void_p, c_ll, c_void_p],
    "sdb_matmul_f64": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p],
}
_RESTYPES = {"sdb_last_error": ctypes.c_char_p, "sdb_launch_count": ctypes.c_ulonglong,
             "sdb_groupnorm_stats_bytes": ctypes.c_longlong}

_lib = None


class SdbError(RuntimeError):
    pass


def lib():
    """Returns the loaded library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SdbError(
                f"{LIB_PATH} is missing: build it with `python -m pytorch_stable_diffusion_b200.csrc.build` "
                "(or __graft_entry__.build()); there is no fallback path")
        l = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        for which, struct in ((0, GemmArgs), (1, AttnArgs)):
            if l.sdb_args_size(which) != ctypes.sizeof(struct):
                raise SdbError(f"{LIB_PATH}: sizeof({struct.__name__}) is {l.sdb_args_size(which)} in the library but "
                               f"{ctypes.sizeof(struct)} in _ext.py - rebuild the library (stale .so?)")
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().sdb_last_error().decode("utf-8", "replace")
        if rc in (-1, -2):
            raise ValueError(f"{what}: {msg} (rc={rc})")
        raise SdbError(f"{what}: {msg} (rc={rc})")


def launch_count():
    return int(lib().sdb_launch_count())


def read_fault():
    v = ctypes.c_uint(0)
    check(lib().sdb_read_fault(ctypes.byref(v)), "sdb_read_fault")
    return v.value
