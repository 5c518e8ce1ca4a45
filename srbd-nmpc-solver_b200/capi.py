"""ctypes mirror of include/srbd_b200.h (the C-ABI) and the loader of libsrbd_b200.so.

There is no CPU fallback: `lib()` raises if the CUDA library has not been built, and every compute
entry point returns SRBD_ERR_CUDA without a usable device.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SRBD_LIB: A/B experiments with another build of the same library (development knob)
LIB_PATH = os.environ.get("SRBD_LIB") or os.path.join(HERE, "libsrbd_b200.so")

SRBD_NX, SRBD_NU, SRBD_NG, SRBD_STAT_M, SRBD_HIST_BINS = 12, 12, 24, 18, 64
SRBD_BARRIER_SOFT, SRBD_HARD_INEQ = 0, 1

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_u8_p = C.POINTER(C.c_uint8)


class ModelParams(C.Structure):
    _fields_ = [("mass", C.c_double), ("dt", C.c_double), ("inertia_inv", C.c_double * 9),
                ("foot_pos", C.c_double * 6), ("foot_rot", C.c_double * 18), ("mu", C.c_double),
                ("Lfx", C.c_double), ("Lfz", C.c_double), ("fmax", C.c_double), ("fmin", C.c_double),
                ("gravity", C.c_double * 3), ("Q", C.c_double * 12), ("Qf", C.c_double * 12),
                ("R", C.c_double), ("mu_b", C.c_double), ("theta_b", C.c_double),
                ("swing_fmax", C.c_double)]


class IpmArgs(C.Structure):
    _fields_ = [("iter_max", C.c_int), ("alpha_min", C.c_double), ("mu0", C.c_double),
                ("tol_stat", C.c_double), ("tol_eq", C.c_double), ("tol_ineq", C.c_double),
                ("tol_comp", C.c_double), ("reg_prim", C.c_double), ("warm_start", C.c_int),
                ("pred_corr", C.c_int), ("ric_alg", C.c_int), ("split_step", C.c_int),
                ("cond_pred_corr", C.c_int), ("cond_factor", C.c_double), ("thr0", C.c_double),
                ("lam_min", C.c_double), ("t_min", C.c_double), ("tau_min", C.c_double),
                ("t_lam_min", C.c_int), ("alpha_shorten", C.c_int),
                ("itref_pred_max", C.c_int), ("itref_corr_max", C.c_int), ("itref_abs", C.c_double),
                ("itref_rel", C.c_double)]


class QpDims(C.Structure):
    _fields_ = [("N", C.c_int), ("nx", C.c_int), ("nu", C.c_int), ("nbx", C.c_int), ("nbu", C.c_int),
                ("ng", C.c_int), ("ngN", C.c_int)]


_QP_FIELDS = ["A", "Bm", "b", "Q", "S", "R", "q", "r", "idxbx", "lbx", "ubx", "lbx_mask", "ubx_mask",
              "idxbu", "lbu", "ubu", "lbu_mask", "ubu_mask", "C", "D", "lg", "ug", "lg_mask", "ug_mask",
              "CN", "lgN", "ugN", "lgN_mask", "ugN_mask", "x0", "x_init", "u_init"]


class QpHost(C.Structure):
    _fields_ = [(n, c_int_p if n in ("idxbx", "idxbu") else c_double_p) for n in _QP_FIELDS]


class SolHost(C.Structure):
    _fields_ = [(n, c_double_p) for n in ["x", "u", "pi", "lam", "t", "P", "p", "K", "k"]]


class StatsHost(C.Structure):
    _fields_ = [("iter", c_int_p), ("status", c_int_p), ("res_max", c_double_p), ("stat", c_double_p)]


class BatchStats(C.Structure):
    _fields_ = [("solves", C.c_longlong), ("iter_sum", C.c_longlong),
                ("iter_hist", C.c_longlong * SRBD_HIST_BINS), ("status_count", C.c_longlong * 5),
                ("res_max", C.c_double * 4)]


def dptr(a):
    """double* of a C-contiguous float64 array (None -> NULL)."""
    if a is None:
        return c_double_p()
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "need C-contiguous float64"
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    if a is None:
        return c_int_p()
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int_p)


def u8ptr(a):
    if a is None:
        return c_u8_p()
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_u8_p)


def qp_nct(d):
    n = 0
    for k in range(d.N + 1):
        nb = (d.nbu if k < d.N else 0) + (d.nbx if k > 0 else 0)
        ng = d.ng if k < d.N else d.ngN
        n += 2 * (nb + ng)
    return n


def make_qp_host(arrays):
    """arrays: dict name -> ndarray (float64 / int32) or None.  Returns (QpHost, keepalive)."""
    qp = QpHost()
    keep = []
    for n in _QP_FIELDS:
        a = arrays.get(n)
        if a is None:
            continue
        if n in ("idxbx", "idxbu"):
            a = np.ascontiguousarray(a, dtype=np.int32)
            setattr(qp, n, iptr(a))
        else:
            a = np.ascontiguousarray(a, dtype=np.float64)
            setattr(qp, n, dptr(a))
        keep.append(a)
    return qp, keep


_EXPORTS = [
    "srbd_model_params_default", "srbd_ipm_args_default", "srbd_ipm_args_set_mode", "srbd_qp_nct", "srbd_ctx_create",
    "srbd_ctx_destroy", "srbd_last_error", "srbd_set_model", "srbd_set_ipm_args", "srbd_set_outputs",
    "srbd_ctx_stat_rows",
    "srbd_ctx_stream", "srbd_ctx_device_ptr", "srbd_ctx_sync", "srbd_ctx_launch_count",
    "srbd_upload_traj", "srbd_download_traj", "srbd_linearize", "srbd_assemble",
    "srbd_download_linearization", "srbd_download_qp", "srbd_qp_upload", "srbd_qp_upload_layout", "srbd_qp_solve",
    "srbd_download_solution", "srbd_download_stats", "srbd_download_ric_lr0", "srbd_batch_stats_get", "srbd_line_search",
    "srbd_download_sqp_state", "srbd_reset_sqp_state", "srbd_sqp_iterate", "srbd_sqp_solve", "srbd_download_sqp_iters", "srbd_solve_host", "srbd_solve_host_async", "srbd_wait", "srbd_solve_host_graph",
    "srbd_fp64_peak", "srbd_mpc_run", "srbd_out_layout", "srbd_download_packed", "srbd_host_alloc", "srbd_host_free",
]

_lib = None


def lib():
    """Load libsrbd_b200.so (built in-tree by __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback for the product path.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.srbd_model_params_default.argtypes = [C.POINTER(ModelParams), C.c_int]
    L.srbd_model_params_default.restype = None
    L.srbd_ipm_args_default.argtypes = [C.POINTER(IpmArgs)]
    L.srbd_ipm_args_default.restype = None
    L.srbd_ipm_args_set_mode.argtypes = [C.POINTER(IpmArgs), C.c_int]
    L.srbd_qp_nct.argtypes = [C.POINTER(QpDims)]
    L.srbd_qp_nct.restype = C.c_size_t
    L.srbd_ctx_create.argtypes = [C.c_int, C.c_int, C.POINTER(QpDims), vp, C.POINTER(vp)]
    L.srbd_ctx_destroy.argtypes = [vp]
    L.srbd_last_error.argtypes = [vp]
    L.srbd_last_error.restype = C.c_char_p
    L.srbd_set_model.argtypes = [vp, C.POINTER(ModelParams)]
    L.srbd_set_ipm_args.argtypes = [vp, C.POINTER(IpmArgs)]
    L.srbd_set_outputs.argtypes = [vp, C.c_int, C.c_int]
    L.srbd_ctx_stat_rows.argtypes = [vp]
    L.srbd_ctx_stream.argtypes = [vp]
    L.srbd_ctx_stream.restype = vp
    L.srbd_ctx_device_ptr.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.srbd_ctx_sync.argtypes = [vp]
    L.srbd_ctx_launch_count.argtypes = [vp]
    L.srbd_ctx_launch_count.restype = C.c_longlong
    L.srbd_upload_traj.argtypes = [vp, c_double_p, c_double_p, c_double_p, c_double_p, c_u8_p]
    L.srbd_download_traj.argtypes = [vp, c_double_p, c_double_p]
    L.srbd_linearize.argtypes = [vp]
    L.srbd_assemble.argtypes = [vp, C.c_int]
    L.srbd_download_linearization.argtypes = [vp] + [c_double_p] * 4
    L.srbd_download_qp.argtypes = [vp] + [c_double_p] * 8
    L.srbd_qp_upload.argtypes = [vp, C.POINTER(QpHost)]
    L.srbd_qp_upload_layout.argtypes = [vp, C.c_int]
    L.srbd_qp_solve.argtypes = [vp]
    L.srbd_download_solution.argtypes = [vp, C.POINTER(SolHost)]
    L.srbd_download_stats.argtypes = [vp, C.POINTER(StatsHost)]
    L.srbd_batch_stats_get.argtypes = [vp, C.POINTER(BatchStats)]
    L.srbd_line_search.argtypes = [vp]
    L.srbd_download_sqp_state.argtypes = [vp, c_double_p, c_int_p, c_double_p]
    L.srbd_reset_sqp_state.argtypes = [vp]
    L.srbd_sqp_iterate.argtypes = [vp, C.c_int, C.c_int]
    L.srbd_sqp_solve.argtypes = [vp, C.c_int, C.c_int]
    L.srbd_download_sqp_iters.argtypes = [vp, c_int_p]
    L.srbd_solve_host.argtypes = [vp, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_u8_p,
                                  c_double_p, c_double_p, c_int_p, c_int_p]
    L.srbd_solve_host_async.argtypes = [vp, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_u8_p,
                                  c_double_p, c_double_p, c_int_p, c_int_p]
    L.srbd_solve_host_graph.argtypes = [vp, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_u8_p,
                                        c_double_p, c_double_p, c_int_p, c_int_p]
    L.srbd_wait.argtypes = [vp]
    L.srbd_fp64_peak.argtypes = [vp, c_double_p]
    L.srbd_out_layout.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.srbd_download_packed.argtypes = [vp, c_double_p, C.c_int]
    L.srbd_download_ric_lr0.argtypes = [vp, c_double_p]
    L.srbd_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.srbd_host_free.argtypes = [vp]
    L.srbd_mpc_run.argtypes = [vp, c_double_p, c_double_p, c_double_p, C.c_int, c_double_p, C.c_int, c_double_p,
                               c_double_p, c_int_p, c_int_p]
    for n in _EXPORTS:
        getattr(L, n)  # every symbol include/srbd_b200.h declares must be exported
    _lib = L
    return L
