"""ctypes binding of libjabble_b200.so (the C ABI of include/jabble_b200.h).

Loading fails loudly: there is no CPU path behind this package.
"""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# JABBLE_B200_LIB: load another build of the same ABI (kernel tuning experiments)
LIB_PATH = os.environ.get("JABBLE_B200_LIB") or os.path.join(_HERE, "libjabble_b200.so")

JB_ABI_VERSION = 1
JB_OK = 0
JB_ERR_BAD_ARG, JB_ERR_UNSUPPORTED, JB_ERR_CUDA = -1, -2, -3
JB_ERR_OUT_OF_RANGE, JB_ERR_WINDOW, JB_ERR_NONFINITE = -4, -5, -6
JB_COMP_SPLINE, JB_COMP_NORM_SPLINE = 0, 1
JB_PLAN_GENERAL_KERNEL = 1
JB_PLAN_FLOAT32 = 2
JB_PLAN_NO_STRIP = 4

c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)
c_i64p = C.POINTER(C.c_int64)


class ComponentDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("p_val", C.c_int32),
        ("xp0", C.c_double), ("dx", C.c_double),
        ("n_knots", C.c_int64), ("ctrl_offset", C.c_int64), ("n_ctrl_keys", C.c_int64),
        ("ctrl_key", c_i32p),
        ("shift_offset", C.c_int64), ("n_shift", C.c_int64), ("shift_key", c_i32p),
        ("stretch_offset", C.c_int64), ("n_stretch", C.c_int64), ("stretch_key", c_i32p),
    ]


class PlanDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32),
        ("n_rows", C.c_int64), ("n_pix", C.c_int64),
        ("xs", c_f64p), ("ys", c_f64p), ("yivar", c_f64p), ("mask", c_u8p),
        ("n_components", C.c_int32), ("components", C.POINTER(ComponentDesc)),
        ("n_params", C.c_int64), ("coefficient", C.c_double),
        ("rows_per_group", C.c_int32), ("flags", C.c_int32),
    ]


class LbfgsOptions(C.Structure):
    """jb_lbfgs_options (include/jabble_b200.h): scipy's fmin_l_bfgs_b options."""
    _fields_ = [("m", C.c_int32), ("maxls", C.c_int32), ("maxiter", C.c_int64), ("maxfun", C.c_int64),
                ("factr", C.c_double), ("pgtol", C.c_double)]


class LbfgsResult(C.Structure):
    """jb_lbfgs_result."""
    _fields_ = [("f", C.c_double), ("pg_norm", C.c_double), ("nit", C.c_int64), ("funcalls", C.c_int64),
                ("warnflag", C.c_int32), ("task", C.c_int32), ("task_msg", C.c_int32),
                ("n_out_of_range", C.c_int32), ("n_skipped", C.c_int32), ("n_restarts", C.c_int32)]


class PlanInfo(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_rows", "n_pix", "n_pix_padded", "n_params", "n_fit", "n_tasks")] + \
               [("rows_per_group", C.c_int32), ("tile_pixels", C.c_int32)] + \
               [(n, C.c_int64) for n in ("device_bytes", "scratch_bytes", "algorithmic_bytes",
                                         "kernel_launches", "evaluations", "serial_rows", "tiles_smooth",
                                         "tiles_general", "tiles_serial", "tiles_empty")] + \
               [("window_knots", C.c_int32), ("fused_variant", C.c_int32), ("strip_reason", C.c_int32), ("strip_rows_per_block", C.c_int32)]


#: every symbol include/jabble_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("jb_last_error", C.c_char_p, []),
    ("jb_abi_version", C.c_int, []),
    ("jb_device_count", C.c_int, []),
    ("jb_plan_create", C.c_int, [C.POINTER(PlanDesc), C.POINTER(C.c_void_p)]),
    ("jb_plan_destroy", None, [C.c_void_p]),
    ("jb_plan_info", C.c_int, [C.c_void_p, C.POINTER(PlanInfo)]),
    ("jb_plan_set_params", C.c_int, [C.c_void_p, c_f64p, C.c_int64]),
    ("jb_plan_set_fit", C.c_int, [C.c_void_p, c_i64p, C.c_int64]),
    ("jb_plan_eval", C.c_int, [C.c_void_p, c_f64p, C.c_int64, c_f64p, c_f64p]),
    ("jb_plan_eval_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("jb_plan_check", C.c_int, [C.c_void_p]),
    ("jb_plan_forward", C.c_int, [C.c_void_p, c_f64p, C.c_int64, c_f64p]),
    ("jb_plan_fisher", C.c_int, [C.c_void_p, C.c_int32, c_f64p, C.c_int64]),
    ("jb_plan_grad_all", C.c_int, [C.c_void_p, c_f64p, C.c_int64]),
    ("jb_plan_curvature", C.c_int, [C.c_void_p, C.c_int32, c_f64p, C.c_int64]),
    ("jb_plan_grid_search", C.c_int, [C.c_void_p, C.c_int32, c_f64p, C.c_int64, C.c_int64, c_f64p]),
    ("jb_batch_create", C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(C.c_void_p)]),
    ("jb_batch_destroy", None, [C.c_void_p]),
    ("jb_batch_bind", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    ("jb_batch_eval_device", C.c_int, [C.c_void_p, C.c_void_p]),
    ("jb_batch_eval_device_checked", C.c_int, [C.c_void_p, C.c_void_p]),
    ("jb_batch_eval_host", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    ("jb_batch_eval_host_begin", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    ("jb_batch_eval_host_end", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    ("jb_batch_host_buffers", C.c_int, [C.c_void_p, C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.POINTER(C.c_double)),
                                        C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("jb_batch_eval_pinned", C.c_int, [C.c_void_p]),
    ("jb_batch_eval_pinned_begin", C.c_int, [C.c_void_p]),
    ("jb_batch_eval_pinned_end", C.c_int, [C.c_void_p]),
    ("jb_batch_soft_range", C.c_int, [C.c_void_p, C.c_int]),
    ("jb_batch_check", C.c_int, [C.c_void_p]),
    ("jb_batch_launches", C.c_int64, [C.c_void_p]),
    ("jb_nccl_unique_id", C.c_int, [C.c_void_p]),
    ("jb_nccl_version", C.c_int, []),
    ("jb_group_create", C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    ("jb_group_destroy", None, [C.c_void_p]),
    ("jb_group_bind", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    ("jb_group_eval_device", C.c_int, [C.c_void_p, C.c_void_p]),
    ("jb_group_eval_device_checked", C.c_int, [C.c_void_p, C.c_void_p]),
    ("jb_group_info", C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    ("jb_group_launches", C.c_int64, [C.c_void_p]),
    ("jb_double_peak", C.c_int, [C.c_int32, c_f64p]),
    ("jb_host_checksum", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("jb_lbfgs_create", C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(LbfgsOptions), C.POINTER(C.c_void_p)]),
    ("jb_lbfgs_destroy", None, [C.c_void_p]),
    ("jb_lbfgs_add_reg", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double, c_i64p, C.c_int64]),
    ("jb_lbfgs_run", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(LbfgsResult)]),
    ("jb_lbfgs_rounds", C.c_int64, [C.c_void_p]),
    ("jb_plan_stream", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    ("jb_plan_profile", C.c_int, [C.c_void_p, C.c_int]),
    ("jb_plan_profile_read", C.c_int, [C.c_void_p, c_f64p, c_i64p]),
]

_lib = None


class JabbleB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libjabble_b200 error {code}: {message}")
        self.code = code


def load():
    """Load the library (once).  Raises if it has not been built: see __graft_entry__.build()."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python wobble_jax_b200/csrc/build.py` "
            "(nvcc, sm_100a). This package has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.jb_abi_version() != JB_ABI_VERSION:
        raise ImportError("libjabble_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(rc):
    if rc != JB_OK:
        msg = load().jb_last_error().decode("utf-8", "replace")
        if rc in (JB_ERR_BAD_ARG, JB_ERR_UNSUPPORTED, JB_ERR_OUT_OF_RANGE, JB_ERR_WINDOW):
            raise ValueError(f"libjabble_b200 error {rc}: {msg}")
        raise JabbleB200Error(rc, msg)
