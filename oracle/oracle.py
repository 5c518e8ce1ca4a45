"""ctypes front-end of the CPU ORACLE (test infrastructure, never imported by the product path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use this.
The struct definitions are shared with the product's C-ABI mirror (capi.py) because the oracle takes
the same POD structs (include/srbd_b200.h).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import srbd_pkg  # noqa: E402

_pkg = srbd_pkg.load()
capi = _pkg.capi

LIB_PATH = os.path.join(HERE, "_build", "libsrbd_oracle.so")
QUAD_PATH = os.path.join(HERE, "_build", "libsrbd_oracle_quad.so")
_lib = None
_qlib = None


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("srbd_model.c", "ocp_qp_ipm.c", "pipeline.c", "srbd_oracle.h")]
    srcs.append(os.path.join(ROOT, "include", "srbd_b200.h"))
    if (not force and all(os.path.exists(p) and all(os.path.getmtime(p) >= os.path.getmtime(s) for s in srcs
                                                      if os.path.exists(s)) for p in (LIB_PATH, QUAD_PATH))):
        return LIB_PATH
    subprocess.check_call(["make", "-C", HERE, "-B", "all"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def quad_lib():
    """The higher-precision ARBITER: ocp_qp_ipm.c compiled with __float128 arithmetic (same algorithm, constants,
    operation order; inputs / outputs double)."""
    global _qlib
    if _qlib is None:
        if not os.path.exists(QUAD_PATH):
            build(force=True)
        L = C.CDLL(QUAD_PATH)
        L.orc_qp_solve_batch_q.argtypes = [C.POINTER(capi.QpDims), C.POINTER(capi.IpmArgs), C.POINTER(capi.QpHost),
                                           C.POINTER(capi.SolHost), C.POINTER(capi.StatsHost), C.c_int, C.c_int, C.c_int]
        _qlib = L
    return _qlib


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    dp, ip, u8 = capi.c_double_p, capi.c_int_p, capi.c_u8_p
    MP, IA, QD = C.POINTER(capi.ModelParams), C.POINTER(capi.IpmArgs), C.POINTER(capi.QpDims)
    L.orc_model_params_default.argtypes = [MP, C.c_int]
    L.orc_model_params_default.restype = None
    L.orc_ipm_args_default.argtypes = [IA]
    L.orc_ipm_args_default.restype = None
    for f in ("orc_skew", "orc_expm", "orc_jl", "orc_jlt", "orc_djl", "orc_djlt"):
        getattr(L, f).argtypes = [dp, dp]
        getattr(L, f).restype = None
    L.orc_continuous.argtypes = [MP, dp, dp, dp, dp, dp]
    L.orc_continuous.restype = None
    L.orc_shooting.argtypes = [MP, dp, dp, dp, dp, dp, dp, dp]
    L.orc_shooting.restype = None
    L.orc_constraint.argtypes = [MP, dp, u8, dp, dp]
    L.orc_constraint.restype = None
    L.orc_barrier.argtypes = [C.c_double, C.c_double, C.c_double, dp, dp, dp]
    L.orc_barrier.restype = None
    L.orc_assemble_batch.argtypes = [MP, C.c_int, C.c_int, C.c_int, dp, dp, dp, u8] + [dp] * 13 + [C.c_int]
    L.orc_line_search.argtypes = [MP, C.c_int, dp, dp, dp, u8, dp, dp, dp, dp]
    L.orc_line_search_mode.argtypes = [MP, C.c_int, C.c_int, dp, dp, dp, u8, dp, dp, dp, dp]
    L.orc_qp_solve_batch.argtypes = [QD, IA, C.POINTER(capi.QpHost), C.POINTER(capi.SolHost),
                                     C.POINTER(capi.StatsHost), C.c_int, C.c_int, C.c_int]
    L.orc_pipeline_batch.argtypes = [MP, IA, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, u8, dp, dp, dp,
                                     dp, dp, ip, ip, dp, C.c_int]
    L.orc_num_threads.restype = C.c_int
    L.orc_qp_nct.argtypes = [QD]
    L.orc_qp_nct.restype = C.c_size_t
    _lib = L
    return L


# ---------------------------------------------------------------------------------------------------
def model_params(horizon=20):
    p = capi.ModelParams()
    lib().orc_model_params_default(C.byref(p), horizon)
    return p


def ipm_args(**kw):
    a = capi.IpmArgs()
    lib().orc_ipm_args_default(C.byref(a))
    for k, v in kw.items():
        if not hasattr(a, k):
            raise AttributeError(k)
        setattr(a, k, v)
    return a


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def so3(name, r):
    out = np.zeros(27 if name in ("djl", "djlt") else 9)
    getattr(lib(), "orc_" + name)(capi.dptr(_f(r)), capi.dptr(out))
    if out.size == 27:
        return out.reshape(3, 3, 3).transpose(0, 2, 1)  # [k][i][j]
    return out.reshape(3, 3).T


def continuous(m, x, u, jac=True):
    dx = np.zeros(12)
    jfx = np.zeros(144) if jac else None
    jfu = np.zeros(144) if jac else None
    lib().orc_continuous(C.byref(m), capi.dptr(_f(x)), capi.dptr(_f(u)), capi.dptr(dx), capi.dptr(jfx),
                         capi.dptr(jfu))
    if jac:
        return dx, jfx.reshape(12, 12).T.copy(), jfu.reshape(12, 12).T.copy()
    return dx


def shooting(m, x, xn, u):
    A, B, b, f = np.zeros(144), np.zeros(144), np.zeros(12), np.zeros(12)
    lib().orc_shooting(C.byref(m), capi.dptr(_f(x)), capi.dptr(_f(xn)), capi.dptr(_f(u)), capi.dptr(A),
                       capi.dptr(B), capi.dptr(b), capi.dptr(f))
    return A.reshape(12, 12).T.copy(), B.reshape(12, 12).T.copy(), b, f


def constraint(m, u, stance=None):
    Ac, f = np.zeros(288), np.zeros(24)
    st = None if stance is None else np.ascontiguousarray(stance, dtype=np.uint8)
    lib().orc_constraint(C.byref(m), capi.dptr(_f(u)), capi.u8ptr(st), capi.dptr(Ac), capi.dptr(f))
    return Ac.reshape(12, 24).T.copy(), f


def barrier(v, mu, theta):
    o = [np.zeros(1) for _ in range(3)]
    lib().orc_barrier(v, mu, theta, *[capi.dptr(a) for a in o])
    return tuple(float(a[0]) for a in o)


def assemble(m, N, mode, x, u, xref, contact=None, threads=0):
    """x [B][N+1][12] ... -> dict of column-major batched arrays (flat per stage)."""
    x, u, xref = _f(x), _f(u), _f(xref)
    B = x.shape[0]
    ct = None if contact is None else np.ascontiguousarray(contact, dtype=np.uint8)
    o = dict(A=np.zeros((B, N, 144)), Bm=np.zeros((B, N, 144)), b=np.zeros((B, N, 12)),
             Q=np.zeros((B, N + 1, 144)), S=np.zeros((B, N, 144)), R=np.zeros((B, N, 144)),
             q=np.zeros((B, N + 1, 12)), r=np.zeros((B, N, 12)), D=np.zeros((B, N, 288)),
             lg=np.zeros((B, N, 24)), lg_mask=np.zeros((B, N, 24)), defect=np.zeros((B, N, 12)),
             fcon=np.zeros((B, N, 24)))
    lib().orc_assemble_batch(C.byref(m), N, mode, B, capi.dptr(x), capi.dptr(u), capi.dptr(xref),
                             capi.u8ptr(ct), *[capi.dptr(o[k]) for k in
                                               ("A", "Bm", "b", "Q", "S", "R", "q", "r", "D", "lg",
                                                "lg_mask", "defect", "fcon")], threads)
    return o


def line_search(m, N, x, u, xref, dx, du, alpha, contact=None, mode=0):
    """One QP.  Returns (x_new, u_new, alpha_new, converged, merit[3])."""
    x, u = _f(x).copy(), _f(u).copy()
    al = np.array([alpha], dtype=np.float64)
    merit = np.zeros(3)
    ct = None if contact is None else np.ascontiguousarray(contact, dtype=np.uint8)
    conv = lib().orc_line_search_mode(C.byref(m), N, mode, capi.dptr(x), capi.dptr(u), capi.dptr(_f(xref)),
                                 capi.u8ptr(ct), capi.dptr(_f(dx)), capi.dptr(_f(du)), capi.dptr(al),
                                 capi.dptr(merit))
    return x, u, float(al[0]), int(conv), merit


def qp_solve(dims, args, arrays, batch, stat_rows=0, threads=0, want=("x", "u", "pi", "lam", "t", "P", "p", "K", "k"),
             quad=False):
    """arrays: dict of hpipm-cpp OcpQp fields batched [B][stage][...] (column-major blocks).
    quad=True: the __float128 arbiter build of the same source (about 50x slower; inputs / outputs double)."""
    qp, keep = capi.make_qp_host(arrays)
    N, nx, nu = dims.N, dims.nx, dims.nu
    nct = capi.qp_nct(dims)
    shapes = dict(x=(batch, N + 1, nx), u=(batch, N, nu), pi=(batch, N + 1, nx), lam=(batch, nct),
                  t=(batch, nct), P=(batch, N + 1, nx * nx), p=(batch, N + 1, nx), K=(batch, N, nu * nx),
                  k=(batch, N, nu))
    out = {k: np.zeros(shapes[k]) for k in want}
    sol = capi.SolHost()
    for k in want:
        setattr(sol, k, capi.dptr(out[k]))
    st = capi.StatsHost()
    out["iter"] = np.zeros(batch, dtype=np.int32)
    out["status"] = np.zeros(batch, dtype=np.int32)
    out["res_max"] = np.zeros((batch, 4))
    st.iter, st.status, st.res_max = capi.iptr(out["iter"]), capi.iptr(out["status"]), capi.dptr(out["res_max"])
    if stat_rows > 0:
        out["stat"] = np.zeros((batch, stat_rows, capi.SRBD_STAT_M))
        st.stat = capi.dptr(out["stat"])
    fn = quad_lib().orc_qp_solve_batch_q if quad else lib().orc_qp_solve_batch
    fn(C.byref(dims), C.byref(args), C.byref(qp), C.byref(sol), C.byref(st), stat_rows, batch, threads)
    del keep
    return out


def pipeline(m, args, N, mode, x, u, xref, x0, contact=None, threads=0, duals=True):
    x, u, xref, x0 = _f(x), _f(u), _f(xref), _f(x0)
    B = x.shape[0]
    ct = None if contact is None else np.ascontiguousarray(contact, dtype=np.uint8)
    out = dict(x=np.zeros((B, N + 1, 12)), u=np.zeros((B, N, 12)), pi=np.zeros((B, N + 1, 12)),
               iter=np.zeros(B, dtype=np.int32), status=np.zeros(B, dtype=np.int32),
               res_max=np.zeros((B, 4)))
    if duals and mode == capi.SRBD_HARD_INEQ:
        out["lam"] = np.zeros((B, N * 48))
        out["t"] = np.zeros((B, N * 48))
    lib().orc_pipeline_batch(C.byref(m), C.byref(args), N, mode, B, capi.dptr(x), capi.dptr(u),
                             capi.dptr(xref), capi.dptr(x0), capi.u8ptr(ct), capi.dptr(out["x"]),
                             capi.dptr(out["u"]), capi.dptr(out["pi"]), capi.dptr(out.get("lam")),
                             capi.dptr(out.get("t")), capi.iptr(out["iter"]), capi.iptr(out["status"]),
                             capi.dptr(out["res_max"]), threads)
    return out


def num_threads():
    return int(lib().orc_num_threads())
