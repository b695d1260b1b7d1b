"""ctypes binding of libxesn_b200.so (C ABI declared in include/xesn_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails the
error is raised immediately.  Build with `python __graft_entry__.py` or
`bash xesn_b200/csrc/build.sh`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# XESN_B200_LIB: another build of the same library (the -DXESN_RECUR_TRACE debug build used by tools/*trace*.py)
LIB_PATH = os.environ.get("XESN_B200_LIB") or os.path.join(_HERE, "libxesn_b200.so")

c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)


class Series(C.Structure):
    """`xesn_series` of include/xesn_b200.h: time-last rows, optionally index-gathered."""
    _fields_ = [
        ("data_dev", C.c_void_p),
        ("ld", C.c_int64),
        ("map_dev", C.c_void_p),
        ("fill_dev", C.c_void_p),
        ("zero_row_dev", C.c_void_p),
    ]


class FlowPlan(C.Structure):
    """`xesn_flow_plan` of include/xesn_b200.h."""
    _fields_ = [
        ("n_import", C.c_int32),
        ("src_dev", C.c_void_p),
        ("import_cell_dev", C.c_void_p),
        ("exp_ptr_dev", C.c_void_p),
        ("exp_peer_dev", C.c_void_p),
        ("exp_slot_dev", C.c_void_p),
        ("world", C.c_int32),
        ("rank", C.c_int32),
        ("ring_ptrs", C.c_uint64 * 16),
        ("launches_done", C.c_uint32),
    ]


# name -> (restype, argtypes); every symbol include/xesn_b200.h declares
SIGNATURES = {
    "xesn_last_error": (C.c_char_p, []),
    "xesn_abi_version": (C.c_int, []),
    "xesn_launch_count": (C.c_int64, []),
    "xesn_reservoir_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]),
    "xesn_reservoir_destroy": (C.c_int, [C.c_void_p]),
    "xesn_reservoir_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "xesn_bank_layout": (C.c_int64, [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                     C.POINTER(C.c_int64)]),
    "xesn_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "xesn_forced_scratch_bytes": (C.c_int64, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "xesn_forced_run": (C.c_int, [C.c_void_p, C.POINTER(Series), C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "xesn_train_scratch_bytes": (C.c_int64, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "xesn_train_accumulate": (C.c_int, [C.c_void_p, C.POINTER(Series), C.POINTER(Series), C.c_int32, C.c_int32,
                                        C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_void_p]),
    "xesn_ridge_solve": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "xesn_ridge_solve_pivoted": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_int64, C.c_void_p]),
    "xesn_predict_scratch_bytes": (C.c_int64, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64]),
    "xesn_predict": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                               C.c_void_p]),
    "xesn_predict_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                    C.c_void_p]),
    "xesn_p2p_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]),
    "xesn_p2p_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "xesn_p2p_close": (C.c_int, [C.c_void_p]),
    "xesn_memcpy_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "xesn_p2p_free": (C.c_int, [C.c_void_p]),
    "xesn_predict_p2p": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_uint32,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "xesn_flow_geometry": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "xesn_predict_flow": (C.c_int, [C.c_void_p, C.POINTER(FlowPlan), C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_void_p]),
    "xesn_memset_async": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]),
    "xesn_gemm_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_void_p]),
    "xesn_set_fused": (C.c_int, [C.c_void_p, C.c_int32]),
    "xesn_cost_psd_nrmse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_void_p, C.c_void_p]),
    "xesn_cost_nrmse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xesn_psd_radial": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.c_void_p, C.c_void_p]),
    "xesn_nrmse_skipnan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "xesn_reservoir_set_group_params": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xesn_set_recur_mode": (C.c_int, [C.c_void_p, C.c_int32]),
    "xesn_set_gram_mode": (C.c_int, [C.c_void_p, C.c_int32]),
    "xesn_get_gram_mode": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "xesn_set_poll_timeout_ms": (C.c_int, [C.c_void_p, C.c_int32]),
    "xesn_set_nan_input_mask": (C.c_int, [C.c_void_p, C.c_int32]),
    "xesn_profile_enable": (C.c_int, [C.c_int32]),
    "xesn_profile_read": (C.c_int, [C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "xesn_profile_regions": (C.c_int64, [C.c_int32, C.POINTER(C.c_double), C.c_int64]),
    "xesn_gram_i8": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "xesn_readout": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
}

_lib = None


class XesnError(RuntimeError):
    pass


def load():
    """Load the shared library once; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise XesnError(
                f"{LIB_PATH} not found: the sm_100a CUDA library has not been built "
                f"(run `python __graft_entry__.py` or `bash xesn_b200/csrc/build.sh`). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().xesn_last_error().decode("utf-8", "replace")
        raise XesnError(f"libxesn_b200 call failed (code {rc}): {msg}")


def launch_count():
    return int(load().xesn_launch_count())


PROFILE_CLASSES = ("drive_gemm", "recurrence", "gram_syrk", "yr_gemm", "ridge_solve", "closed_loop", "halo_gather",
                   "fused_chunk")


def profile_enable(on=True):
    check(load().xesn_profile_enable(1 if on else 0))


def profile_regions(name, capacity=4096):
    """ms of every timed region of one class, in launch order."""
    buf = (C.c_double * capacity)()
    n = int(load().xesn_profile_regions(PROFILE_CLASSES.index(name), buf, capacity))
    if n < 0:
        check(n)
    return [float(buf[i]) for i in range(min(n, capacity))]


def profile_read():
    """{class name: (regions, total_ms)} of the phases timed since profile_enable(True)."""
    lib = load()
    out = {}
    for k, name in enumerate(PROFILE_CLASSES):
        n, ms = C.c_int64(0), C.c_double(0.0)
        check(lib.xesn_profile_read(k, C.byref(n), C.byref(ms)))
        out[name] = (int(n.value), float(ms.value))
    return out
