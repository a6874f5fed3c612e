"""ctypes binding of libu2pl_b200.so (the C ABI declared in include/u2pl_b200.h).

There is no fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libu2pl_b200.so")

_P = c_void_p       # device pointer
_S = c_void_p       # cudaStream_t

# name -> (restype, argtypes).  tests/test_cabi_symbols.py checks this table against the header.
SIGNATURES = {
    "u2pl_abi_version": (c_int, []),
    "u2pl_last_error": (c_char_p, []),
    "u2pl_launch_count": (c_int64, []),
    "u2pl_entropy_ws_bytes": (c_size_t, [c_int64, c_int64]),
    "u2pl_entropy_thresholds": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, POINTER(c_float), c_int,
                                        _P, _P, _P, _P, c_size_t, _S]),
    "u2pl_entropy_fast_ws_bytes": (c_size_t, [c_int64, c_int64]),
    "u2pl_entropy_thresholds_fast": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, POINTER(c_float), c_int,
                                             _P, _P, _P, _P, c_size_t, _S]),
    "u2pl_entropy_partition_fused": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, POINTER(c_float), c_int, c_int,
                                             _P, _P, _P, _P, _P, _P, _P, c_size_t, _S]),
    "u2pl_partition_target": (c_int, [_P, _P, c_int64, c_int64, _P, c_int, _P, _P, _S]),
    "u2pl_entropy_masks": (c_int, [_P, _P, _P, c_int64, c_int64, _P, c_int, c_int, _P, _P, _S]),
    "u2pl_upsample_fused_supported": (c_int, [c_int64]),
    "u2pl_upce_ws_bytes": (c_size_t, []),
    "u2pl_up_softmax_max": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P, _S]),
    "u2pl_upce_forward": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, c_size_t, _S]),
    "u2pl_upce_backward": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P, _S]),
    "u2pl_ce_ws_bytes": (c_size_t, [c_int64, c_int64]),
    "u2pl_ce_forward": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, c_size_t, _S]),
    "u2pl_ce_backward": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, _P, _P, _S]),
    "u2pl_unsup_finalize": (c_int, [_P, _P, c_int64, _P, _P, _P, _S]),
    "u2pl_ohem_select": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, c_float, c_int64, _P, _P, _P, _P, c_size_t, _S]),
    "u2pl_onehot_to_bits": (c_int, [_P, c_int64, c_int64, c_int64, _P, _S]),
    "u2pl_contra_prep_lowres": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                        c_int64, c_int64, c_int, _P, _P, _P, _S]),
    "u2pl_contra_num_blocks": (c_int64, [c_int64]),
    "u2pl_contra_classify": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, c_int64, c_int64, c_float, c_float,
                                     c_int, c_int, _P, _P, _P, _P, _S]),
    "u2pl_contra_proto_parts": (c_int64, []),
    "u2pl_contra_proto": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, _P, _S]),
    "u2pl_contra_pack_keys": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, _P, _S]),
    "u2pl_bank_append": (c_int, [_P, _P, c_int64, _P, c_int, c_int64, _S]),
    "u2pl_infonce_forward": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P,
                                     c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P, _S]),
    "u2pl_infonce_forward_sharded": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P,
                                             c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P, _S]),
    "u2pl_shard_alloc": (c_int, [c_int64, _P, _P]),
    "u2pl_shard_open": (c_int, [_P, _P]),
    "u2pl_shard_close": (c_int, [_P, c_int]),
    "u2pl_infonce_backward": (c_int, [_P, _P, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P, _S]),
    "u2pl_gemm_bf16_tn": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, _P, _P, c_int, _S]),
    "u2pl_conv_bf16_nhwc": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int, _P, _P, _P, c_int, _S]),
    "u2pl_conv_stat_parts": (c_int64, [c_int64, c_int64, c_int64, c_int]),
    "u2pl_conv_bf16_nhwc_stats": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int, _P, _P, _S]),
    "u2pl_conv_bf16_nhwc_ex": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int, _P, _P, c_int,
                                       _P, _P, _P, c_int, _P, _P, _S]),
    "u2pl_conv_wgrad_splits": (c_int, [c_int64, c_int64, c_int64, c_int64, c_int64]),
    "u2pl_conv_wgrad_bf16_nhwc": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, _S]),
    "u2pl_maxpool3s2_out": (c_int64, [c_int64]),
    "u2pl_maxpool3s2_forward": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_int64, _S]),
    "u2pl_maxpool3s2_backward": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_int64, _S]),
    "u2pl_peer_region_bytes": (c_int64, []),
    "u2pl_peer_max_floats": (c_int64, []),
    "u2pl_peer_allreduce_f32": (c_int, [_P, c_int64, POINTER(c_void_p), c_int, c_int, c_uint32, _S]),
    "u2pl_sgd_tensor_bytes": (c_int64, []),
    "u2pl_sgd_chunk_elems": (c_int64, []),
    "u2pl_sgd_ema_step": (c_int, [_P, _P, c_int64, c_float, c_float, c_float, c_int, _S]),
    "u2pl_bn_parts": (c_int64, []),
    "u2pl_bn_stats": (c_int, [_P, c_int64, c_int64, _P, _P, _S]),
    "u2pl_bn_finalize": (c_int, [_P, c_int64, c_double, _P, _P, _P, _P, c_float, c_float, _P, _P, _P, _P, _S]),
    "u2pl_bn_fold": (c_int, [c_int64, _P, _P, _P, _P, c_float, _P, _P, _S]),
    "u2pl_bn_apply": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int, _P, _S]),
    "u2pl_bn_backward_reduce": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, _P, _P, _S]),
    "u2pl_bn_backward_elemt": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_double, c_int64, c_int64, _P, _P, _P, _S]),
}

_lib = None


class U2PLNativeError(RuntimeError):
    pass


def load(build_if_missing=True):
    """Load (building first if needed) libu2pl_b200.so and attach prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if build_if_missing and not _build.is_current():
        # stale or missing library: rebuild (under build.py's file lock, atomically replaced).  A failed rebuild is
        # fatal -- loading a stale library against new Python signatures is worse than stopping.  The only tolerated
        # case is a machine without nvcc whose library travelled with its stamp removed (then the ABI version and the
        # symbol table below are the check).
        import shutil
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            _build.build()
        elif not os.path.exists(LIB_PATH):
            raise U2PLNativeError(f"{LIB_PATH} is missing and nvcc is not available (there is no fallback path)")
    if not os.path.exists(LIB_PATH):
        raise U2PLNativeError(f"{LIB_PATH} is missing; run `python -m u2pl_b200.build` (there is no fallback path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library drift: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.u2pl_abi_version() != 1:
        raise U2PLNativeError("libu2pl_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().u2pl_last_error()
        raise U2PLNativeError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def launch_count():
    return int(load().u2pl_launch_count())
