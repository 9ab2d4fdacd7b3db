"""ctypes binding of the C ABI declared in include/r4r_b200.h.

The shared library is built in-tree by ``__graft_entry__.build()`` / ``make -C reviews4rec_b200/csrc``.
Import fails loudly (ImportError) when it is missing -- there is no fallback implementation.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# R4R_LIB: load another build of the SAME library (kernel experiments under scripts/); never a fallback
LIB_PATH = os.environ.get("R4R_LIB") or os.path.join(_HERE, "libr4r_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "reviews4rec_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C reviews4rec_b200/csrc`; there is no CPU/PyTorch fallback." % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

c_int, c_i64, c_f32, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

# name -> (restype, argtypes); must list every symbol of include/r4r_b200.h (tests check this)
SIGNATURES = {
    "r4r_abi_version": (c_int, []),
    "r4r_last_error": (ctypes.c_char_p, []),
    "r4r_device_info": (c_int, [ctypes.POINTER(c_int)] * 3 + [ctypes.POINTER(c_i64)]),
    "r4r_word_gather_f32": (c_int, [c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_vp]),
    "r4r_docs_expand": (c_int, [c_vp, c_vp, c_i64, c_int, c_i64, c_vp, c_vp]),
    "r4r_docs_assemble": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_i64, c_int,
                                  c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "r4r_shadow_build": (c_int, [c_vp, c_i64, c_int, c_vp, c_int, c_int, c_vp]),
    "r4r_conv_pool_simt": (c_int, [c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    "r4r_conv_wpack_bytes": (c_i64, [c_int, c_int]),
    "r4r_conv_pack_weights": (c_int, [c_vp, c_int, c_int, c_vp, c_int, c_vp]),
    "r4r_conv_stream_ws_bytes": (c_i64, [c_i64, c_int]),
    "r4r_conv_pool_tc": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_i64, c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "r4r_doc_plan_ws_bytes": (c_i64, [c_i64, c_int]),
    "r4r_doc_plan": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp]),
    "r4r_conv_pool_tc_ragged": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "r4r_doc_plan_ragged": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp]),
    "r4r_conv_wgrad_argmax_h_ragged": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    "r4r_conv_debug_profile": (c_int, [c_vp]),
    "r4r_conv_set_clusters": (c_int, [c_int]),
    "r4r_conv_wgrad_argmax": (c_int, [c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    "r4r_conv_refine": (c_int, [c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp]),
    "r4r_conv_wgrad_argmax_h": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    "r4r_conv_dgrad_scatter": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_i64, c_vp]),
    "r4r_linear_fwd": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp]),
    "r4r_linear_bwd": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "r4r_fm_fwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp]),
    "r4r_fm_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "r4r_deepconn_head_fwd": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_f32, ctypes.c_uint64, c_vp]),
    "r4r_deepconn_head_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_f32, c_vp]),
    "r4r_mse_fwd": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "r4r_mse_bwd": (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "r4r_rows_argmax": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    "r4r_rows_gather": (c_int, [c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_vp]),
    "r4r_rows_scatter_add": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_i64, c_vp]),
    "r4r_adam_step": (c_int, [c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp] + [c_f32] * 5 + [c_vp]),
    "r4r_counter_inc": (c_int, [c_vp, c_vp]),
    "r4r_shard_mark": (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp]),
    "r4r_shard_plan": (c_int, [c_vp, c_i64, c_int, c_i64, c_vp, c_vp]),
    "r4r_shard_bucket": (c_int, [c_vp, c_i64, c_i64, c_int, c_i64, c_vp, c_vp, c_vp]),
    "r4r_shard_serve": (c_int, [c_vp, c_i64, c_int, c_vp, c_int, c_i64, c_vp, c_vp]),
    "r4r_shard_serve_p2p": (c_int, [c_vp, c_i64, c_int, c_vp, c_int, c_i64, c_vp, c_vp]),
    "r4r_shard_place": (c_int, [c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_i64, c_vp]),
    "r4r_shard_scatter_add": (c_int, [c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_i64, c_f32, c_vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args

R4R_DT_F16, R4R_DT_BF16 = 0, 1
ABI_VERSION = 5

if lib.r4r_abi_version() != ABI_VERSION:
    raise ImportError("reviews4rec_b200: libr4r_b200.so ABI %d != expected %d -- rebuild" % (lib.r4r_abi_version(), ABI_VERSION))

# number of kernel launches issued through this binding (bench.py reports it as gpu_launches)
launch_count = 0
_LAUNCHES_PER_CALL = {"r4r_conv_pool_simt": 2, "r4r_linear_bwd": 2, "r4r_shard_bucket": 2, "r4r_shard_plan": 2, "r4r_doc_plan": 3, "r4r_doc_plan_ragged": 3,
                      "r4r_conv_pool_tc": 3, "r4r_conv_pool_tc_ragged": 3}


def check(rc, name="r4r"):
    if rc != 0:
        msg = lib.r4r_last_error()
        raise RuntimeError("%s failed (code %d): %s" % (name, rc, msg.decode() if msg else "?"))


def call(name, *args):
    """Invoke a C-ABI entry point and raise RuntimeError on a non-zero return."""
    global launch_count
    launch_count += _LAUNCHES_PER_CALL.get(name, 1)
    check(getattr(lib, name)(*args), name)
