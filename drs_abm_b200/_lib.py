"""ctypes binding of libdrs_b200.so (include/drs_b200.h).

There is no CPU fallback: if the library is missing or a call fails, an
exception is raised.  Status codes map to the exceptions the reference's igraph
calls would raise (ValueError for bad vertices / arguments).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DRS_B200_LIB") or os.path.join(_HERE, "libdrs_b200.so")   # env override: kernel-variant experiments

MODE_F32 = 0
MODE_I32 = 1
FW_TILE = 128
APSP_FW, APSP_SSSP, APSP_AUTO, APSP_SEP = 0, 1, 2, 3
PRED_LAZY, PRED_MATRIX = 0, 1
I32_INF = 0x3FFFFFFF

QUEUE_TRACK_CLD = 1
ERR_ARG, ERR_CUDA, ERR_STATE, ERR_NOMEM, ERR_RANGE, ERR_CAPACITY = -1, -2, -3, -4, -5, -6


class DrsError(RuntimeError):
    """A libdrs_b200 call failed for a reason that is not the caller's argument."""


_lib = None

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p

_i8p = C.POINTER(C.c_int8)
_u8p = C.POINTER(C.c_uint8)
_i64p = C.POINTER(C.c_int64)
_u64p = C.POINTER(C.c_uint64)


class VehicleBatch(C.Structure):
    """drs_vehicle_batch (include/drs_b200.h)."""
    _fields_ = [("n_veh", C.c_int32), ("start_node", _i32p), ("start_delay", _f64p), ("capacity", _i32p),
                ("onboard", _i32p), ("stop_off", _i32p), ("stop_node", _i32p), ("stop_rid", _i32p),
                ("stop_pod", _i8p), ("stop_prev", _i8p), ("stop_lb", _f64p), ("stop_ub", _f64p)]


class RequestBatch(C.Structure):
    """drs_request_batch (include/drs_b200.h)."""
    _fields_ = [("n_req", C.c_int32), ("rid", _i32p), ("o_node", _i32p), ("d_node", _i32p), ("lb_p", _f64p),
                ("ub_p", _f64p), ("lb_d", _f64p), ("ub_d", _f64p)]


class PlanBatch(C.Structure):
    """drs_plan_batch (include/drs_b200.h)."""
    _fields_ = [("n_veh", C.c_int32), ("start_node", _i32p), ("start_delay", _f64p), ("onboard", _i32p),
                ("clock", _f64p), ("capacity", _i32p), ("cost", _f64p), ("stop_off", _i32p), ("stop_node", _i32p),
                ("stop_req", _i32p), ("stop_pod", _i8p), ("pos_node", _i32p), ("keep_node", _i32p), ("keep_t", _f64p),
                ("keep_et", _f64p)]


class RequestTable(C.Structure):
    """drs_request_table (include/drs_b200.h)."""
    _fields_ = [("n_rows", C.c_int32), ("o_node", _i32p), ("d_node", _i32p), ("Cep", _f64p), ("Clp", _f64p),
                ("Cld", _f64p), ("Ts", _f64p), ("Tr", _f64p), ("pwt", _f64p)]


class InsertionParams(C.Structure):
    """drs_insertion_params (include/drs_b200.h)."""
    _fields_ = [("max_detour", C.c_double), ("coef_inveh", C.c_double), ("coef_wait", C.c_double),
                ("coef_extra_wait", C.c_double)]


# name -> (restype, argtypes); kept in one table so tests can check that every
# symbol declared in include/drs_b200.h is exported and bound.
SIGNATURES = {
    "drs_last_error": (C.c_char_p, []),
    "drs_version": (C.c_int, []),
    "drs_graph_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _i32p, _i32p, _f64p, _f64p, C.POINTER(_vp)]),
    "drs_graph_set_weights": (C.c_int, [_vp, _f64p]),
    "drs_locality_order": (C.c_int, [C.c_int32, _f64p, _f64p, _i32p]),
    "drs_graph_destroy": (C.c_int, [_vp]),
    "drs_pool_trim": (C.c_int, []),
    "drs_graph_info": (C.c_int, [_vp, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p]),
    "drs_apsp_build": (C.c_int, [_vp, C.c_int32]),
    "drs_apsp_matrices": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_int64)]),
    "drs_apsp_adopt": (C.c_int, [_vp, C.c_int32, _vp, _vp, C.c_int64]),
    "drs_apsp_rows_to_host": (C.c_int, [_vp, C.c_int32, C.c_int32, _f64p]),
    "drs_fw_init": (C.c_int, [_vp, C.c_int32, _vp, C.c_int64, C.c_int32, C.c_int32, _vp]),
    "drs_fw_pivot_row": (C.c_int, [C.c_int32, _vp, C.c_int64, C.c_int32, _vp]),
    "drs_fw_update": (C.c_int, [C.c_int32, _vp, C.c_int64, C.c_int32, _vp, C.c_int32, C.c_int32, C.c_int32,
                                C.c_int32, _vp]),
    "drs_p2p_alloc": (C.c_int, [C.c_int64, C.POINTER(_vp), C.POINTER(C.c_ubyte)]),
    "drs_p2p_open": (C.c_int, [C.POINTER(C.c_ubyte), C.POINTER(_vp)]),
    "drs_p2p_close": (C.c_int, [_vp]),
    "drs_p2p_free": (C.c_int, [_vp]),
    "drs_fw_pivot_row_p2p": (C.c_int, [C.c_int32, _vp, C.c_int64, C.c_int32, C.c_int32, C.POINTER(_vp), C.POINTER(_vp), C.c_int32, _vp]),
    "drs_fw_post_words": (C.c_int, [C.c_int32, C.POINTER(_vp), C.c_int32, _vp]),
    "drs_fw_wait_words": (C.c_int, [C.c_int32, C.POINTER(_vp), C.c_int32, C.c_int64, _vp, _vp]),
    "drs_pred_build": (C.c_int, [_vp, C.c_int32, _vp, C.c_int64, C.c_int32, C.c_int32, _vp, _vp]),
    "drs_query_pairs": (C.c_int, [_vp, C.c_int32, _i32p, _i32p, _f64p, _f64p, _i32p]),
    "drs_query_pair": (C.c_int, [_vp, C.c_int32, C.c_int32, _f64p, _f64p, _i32p]),
    "drs_query_matrix": (C.c_int, [_vp, C.c_int32, _i32p, _i32p, _f64p]),
    "drs_path": (C.c_int, [_vp, C.c_int32, C.c_int32, _i32p, _i32p, _i32p]),
    "drs_alu_peak": (C.c_int, [C.c_int32, C.c_int32, _f64p]),
    "drs_graph_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "drs_apsp_set_algorithm": (C.c_int, [_vp, C.c_int32]),
    "drs_apsp_set_pred_mode": (C.c_int, [_vp, C.c_int32]),
    "drs_sssp_batch_dev": (C.c_int, [_vp, C.c_int32, C.c_int32, _i32p, _vp, C.c_int64, C.c_double, _vp]),
    "drs_sssp_batch": (C.c_int, [_vp, C.c_int32, C.c_int32, _i32p, _f64p]),
    "drs_apsp_rows_sssp": (C.c_int, [_vp, C.c_int32, _vp, C.c_int64, C.c_int32, C.c_int32, C.c_double, _vp]),
    "drs_partition_order": (C.c_int, [C.c_int32, C.c_int32, _i32p, _i32p, _f64p, _f64p, C.c_int32, _i32p, _i32p, C.c_int32, _i32p,
                                      _i32p, _i32p]),
    "drs_graph_set_partition": (C.c_int, [_vp, C.c_int32, _i32p, C.c_int32, _i32p]),
    "drs_sep_info2": (C.c_int, [_vp, _i32p, _i32p, _i32p]),
    "drs_sep_work": (C.c_int, [_vp, C.c_int32, C.c_int32, _f64p]),
    "drs_sep_info": (C.c_int, [_vp, _i32p, _i32p, _i32p, _i32p, _i32p]),
    "drs_apsp_rows_sep": (C.c_int, [_vp, C.c_int32, _vp, C.c_int64, C.c_int32, C.c_int32, _vp]),
    "drs_apsp_sep_profile": (C.c_int, [_vp, _f64p]),
    "drs_launch_count": (C.c_int64, []),
    "drs_tick_last_ms": (C.c_int, [_vp, _f64p]),
    "drs_apsp_last_profile": (C.c_int, [_vp, _f64p, _i32p]),
    "drs_tick_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "drs_tick_destroy": (C.c_int, [_vp]),
    "drs_tick_set_vehicles": (C.c_int, [_vp, C.POINTER(VehicleBatch)]),
    "drs_tick_set_requests": (C.c_int, [_vp, C.POINTER(RequestBatch)]),
    "drs_tick_rv_eval": (C.c_int, [_vp, C.c_double, C.c_int32, _i32p, _i64p]),
    "drs_tick_rv_fetch": (C.c_int, [_vp, C.c_int32, _i32p, _i32p, _f64p, _u8p, _i32p]),
    "drs_tick_set_plans": (C.c_int, [_vp, C.POINTER(PlanBatch), C.POINTER(RequestTable), C.POINTER(InsertionParams)]),
    "drs_tick_insertion_eval": (C.c_int, [_vp, C.c_int32, _i32p, _i32p, _i32p, _i32p, _f64p, _i64p]),
    "drs_tick_insertion_queue": (C.c_int, [_vp, C.c_int32, _i32p, C.c_double, C.c_int32, _i32p, _i32p, _i32p, _f64p, _i64p]),
    "drs_tick_commit_insertion": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double]),
    "drs_tick_fetch_plans": (C.c_int, [_vp, _f64p, _i32p, _i32p, _i8p, C.c_int32, _f64p, _f64p]),
    "drs_alloc_count": (C.c_int64, []),
    "drs_tick_last_insertion_ms": (C.c_int, [_vp, _f64p]),
    "drs_tick_trip_eval": (C.c_int, [_vp, C.c_double, C.c_int32, C.c_int32, _i32p, _i32p, _i32p, _f64p, _u8p,
                                     _i32p]),
    "drs_rtv_level": (C.c_int, [C.c_int32, C.c_int32, _i32p, _i32p, _u8p, C.c_int32, _u64p, C.c_int64, _i32p, _i32p,
                                _i64p, _i64p]),
}


def load():
    """Load the shared library, binding every entry point.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DrsError(
            "libdrs_b200.so is not built (%s). Run `python -m drs_abm_b200.build`; "
            "there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, lib=None):
    if rc == 0:
        return
    lib = lib or load()
    msg = lib.drs_last_error().decode("utf-8", "replace")
    if rc in (ERR_ARG, ERR_RANGE, ERR_CAPACITY):
        raise ValueError(msg)
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    raise DrsError("libdrs_b200 error %d: %s" % (rc, msg))


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr_i32(a):
    return a.ctypes.data_as(_i32p)


def ptr_f64(a):
    return a.ctypes.data_as(_f64p)
