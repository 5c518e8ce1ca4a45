"""Thin Python binding over the C-ABI (include/srbd_b200.h) used by tests/ and bench.py.

The reference's host language is C++, so the drop-in facades live in host/*.hpp; this binding only
moves numpy buffers through the same entry points.  PyTorch is used for plumbing only (streams,
device tensors over context buffers, torch.distributed).
"""
import ctypes as C

import numpy as np

from . import capi


class SrbdError(RuntimeError):
    pass


def default_model_params(horizon=20):
    p = capi.ModelParams()
    capi.lib().srbd_model_params_default(C.byref(p), horizon)
    return p


def default_ipm_args(**kw):
    a = capi.IpmArgs()
    capi.lib().srbd_ipm_args_default(C.byref(a))
    for k, v in kw.items():
        if not hasattr(a, k):
            raise AttributeError(k)
        setattr(a, k, v)
    return a


def make_dims(N=20, nx=12, nu=12, nbx=0, nbu=0, ng=24, ngN=0):
    d = capi.QpDims()
    d.N, d.nx, d.nu, d.nbx, d.nbu, d.ng, d.ngN = N, nx, nu, nbx, nbu, ng, ngN
    return d


_BUF_DTYPE = {4: np.uint8, 10: np.int32, 11: np.int32}


class _CudaArray:
    """Minimal __cuda_array_interface__ carrier so torch.as_tensor can view a context buffer."""

    def __init__(self, ptr, nbytes, dtype):
        dt = np.dtype(dtype)
        self.__cuda_array_interface__ = {"shape": (nbytes // dt.itemsize,), "typestr": dt.str,
                                         "data": (ptr, False), "version": 2}


class Context:
    """Owns one srbd_ctx (device + pinned host buffers for `batch` QPs of dimensions `dims`)."""

    def __init__(self, batch, dims=None, device=0, stream=None):
        self._L = capi.lib()
        self.dims = dims if dims is not None else make_dims()
        self.batch = int(batch)
        h = C.c_void_p()
        rc = self._L.srbd_ctx_create(int(device), self.batch, C.byref(self.dims),
                                     C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise SrbdError(f"srbd_ctx_create failed with {rc} (no usable CUDA device?)")
        self._h = h
        self.nct = capi.qp_nct(self.dims)

    # -- plumbing ------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            msg = self._L.srbd_last_error(self._h)
            raise SrbdError(f"srbd error {rc}: {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "_h", None):
            self._L.srbd_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        self._ck(self._L.srbd_ctx_sync(self._h))

    @property
    def stream(self):
        return self._L.srbd_ctx_stream(self._h)

    @property
    def launch_count(self):
        return int(self._L.srbd_ctx_launch_count(self._h))

    @property
    def stat_rows(self):
        return int(self._L.srbd_ctx_stat_rows(self._h))

    def device_tensor(self, buf):
        """torch view (no copy) of a context device buffer."""
        import torch
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self._L.srbd_ctx_device_ptr(self._h, int(buf), C.byref(p), C.byref(n)))
        return torch.as_tensor(_CudaArray(p.value, n.value, _BUF_DTYPE.get(int(buf), np.float64)),
                               device="cuda")

    def set_model(self, p):
        self._model = p
        self._ck(self._L.srbd_set_model(self._h, C.byref(p)))

    def set_ipm_args(self, a):
        self._args = a
        self._ck(self._L.srbd_set_ipm_args(self._h, C.byref(a)))

    def set_outputs(self, export_ric=False, export_stat=False):
        self._ck(self._L.srbd_set_outputs(self._h, int(bool(export_ric)), int(bool(export_stat))))

    # -- NMPC level ------------------------------------------------------------------------------
    def upload_traj(self, x, u, xref, x0, contact=None):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        ct = None if contact is None else np.ascontiguousarray(contact, dtype=np.uint8)
        self._ck(self._L.srbd_upload_traj(self._h, capi.dptr(f(x)), capi.dptr(f(u)), capi.dptr(f(xref)),
                                          capi.dptr(f(x0)), capi.u8ptr(ct)))

    def download_traj(self):
        B, N = self.batch, self.dims.N
        x, u = np.zeros((B, N + 1, 12)), np.zeros((B, N, 12))
        self._ck(self._L.srbd_download_traj(self._h, capi.dptr(x), capi.dptr(u)))
        return x, u

    def linearize(self):
        self._ck(self._L.srbd_linearize(self._h))

    def assemble(self, mode):
        self._ck(self._L.srbd_assemble(self._h, int(mode)))

    def download_linearization(self):
        B, N = self.batch, self.dims.N
        o = dict(A=np.zeros((B, N, 144)), Bm=np.zeros((B, N, 144)), b=np.zeros((B, N, 12)),
                 defect=np.zeros((B, N, 12)))
        self._ck(self._L.srbd_download_linearization(self._h, *[capi.dptr(o[k]) for k in ("A", "Bm", "b", "defect")]))
        return o

    def download_qp(self):
        B, N = self.batch, self.dims.N
        o = dict(Q=np.zeros((B, N + 1, 144)), S=np.zeros((B, N, 144)), R=np.zeros((B, N, 144)),
                 q=np.zeros((B, N + 1, 12)), r=np.zeros((B, N, 12)), D=np.zeros((B, N, 288)),
                 lg=np.zeros((B, N, 24)), lg_mask=np.zeros((B, N, 24)))
        self._ck(self._L.srbd_download_qp(self._h, *[capi.dptr(o[k]) for k in
                                                    ("Q", "S", "R", "q", "r", "D", "lg", "lg_mask")]))
        return o

    # -- QP level ---------------------------------------------------------------------------------
    def qp_upload(self, arrays, d_shared=False):
        """d_shared: arrays["D"] is ONE ng x nu matrix shared by every stage of every QP (srbd_qp_upload_layout)."""
        qp, keep = capi.make_qp_host(arrays)
        self._ck(self._L.srbd_qp_upload_layout(self._h, 1 if d_shared else 0))
        self._ck(self._L.srbd_qp_upload(self._h, C.byref(qp)))
        del keep

    def qp_solve(self):
        self._ck(self._L.srbd_qp_solve(self._h))

    def download_solution(self, want=("x", "u", "pi", "lam", "t", "P", "p", "K", "k")):
        B, d = self.batch, self.dims
        N, nx, nu = d.N, d.nx, d.nu
        shapes = dict(x=(B, N + 1, nx), u=(B, N, nu), pi=(B, N + 1, nx), lam=(B, self.nct), t=(B, self.nct),
                      P=(B, N + 1, nx * nx), p=(B, N + 1, nx), K=(B, N, nu * nx), k=(B, N, nu))
        out = {k: np.zeros(shapes[k]) for k in want}
        sol = capi.SolHost()
        for k in want:
            setattr(sol, k, capi.dptr(out[k]))
        self._ck(self._L.srbd_download_solution(self._h, C.byref(sol)))
        return out

    def download_stats(self, with_table=False):
        B = self.batch
        out = dict(iter=np.zeros(B, dtype=np.int32), status=np.zeros(B, dtype=np.int32),
                   res_max=np.zeros((B, 4)))
        st = capi.StatsHost()
        st.iter, st.status, st.res_max = capi.iptr(out["iter"]), capi.iptr(out["status"]), capi.dptr(out["res_max"])
        if with_table:
            out["stat"] = np.zeros((B, self.stat_rows, capi.SRBD_STAT_M))
            st.stat = capi.dptr(out["stat"])
        self._ck(self._L.srbd_download_stats(self._h, C.byref(st)))
        return out

    def batch_stats(self):
        bs = capi.BatchStats()
        self._ck(self._L.srbd_batch_stats_get(self._h, C.byref(bs)))
        return dict(solves=int(bs.solves), iter_sum=int(bs.iter_sum), iter_hist=list(bs.iter_hist),
                    status_count=list(bs.status_count), res_max=list(bs.res_max))

    # -- SQP level ---------------------------------------------------------------------------------
    def line_search(self):
        self._ck(self._L.srbd_line_search(self._h))

    def reset_sqp_state(self):
        self._ck(self._L.srbd_reset_sqp_state(self._h))

    def download_sqp_state(self):
        B = self.batch
        alpha, conv, merit = np.zeros(B), np.zeros(B, dtype=np.int32), np.zeros((B, 3))
        self._ck(self._L.srbd_download_sqp_state(self._h, capi.dptr(alpha), capi.iptr(conv), capi.dptr(merit)))
        return alpha, conv, merit

    def sqp_iterate(self, mode, do_line_search=False):
        self._ck(self._L.srbd_sqp_iterate(self._h, int(mode), int(bool(do_line_search))))

    def sqp_solve(self, mode, max_iter):
        """The outer SQP loop on the device (srbd_sqp_solve); returns the SQP iterations each problem took."""
        self._ck(self._L.srbd_sqp_solve(self._h, int(mode), int(max_iter)))
        it = np.zeros(self.batch, dtype=np.int32)
        self._ck(self._L.srbd_download_sqp_iters(self._h, capi.iptr(it)))
        return it

    def solve_host(self, mode, x, u, xref, x0, contact, sol_x, sol_u, it, status):
        """End-to-end call on caller-owned (ideally pinned) host buffers; nothing is allocated here."""
        self._ck(self._L.srbd_solve_host(self._h, int(mode), capi.dptr(x), capi.dptr(u), capi.dptr(xref),
                                         capi.dptr(x0), capi.u8ptr(contact), capi.dptr(sol_x),
                                         capi.dptr(sol_u), capi.iptr(it), capi.iptr(status)))

    def solve_host_graph(self, mode, x, u, xref, x0, contact, sol_x, sol_u, it, status):
        """The low-latency form: pinned staging inside the context, the whole pipeline replayed as one CUDA graph."""
        self._ck(self._L.srbd_solve_host_graph(self._h, int(mode), capi.dptr(x), capi.dptr(u), capi.dptr(xref),
                                               capi.dptr(x0), capi.u8ptr(contact), capi.dptr(sol_x),
                                               capi.dptr(sol_u), capi.iptr(it), capi.iptr(status)))

    def solve_host_async(self, mode, x, u, xref, x0, contact, sol_x, sol_u, it, status):
        """End-to-end call on caller-owned (ideally pinned) host buffers; nothing is allocated here."""
        self._ck(self._L.srbd_solve_host_async(self._h, int(mode), capi.dptr(x), capi.dptr(u), capi.dptr(xref),
                                         capi.dptr(x0), capi.u8ptr(contact), capi.dptr(sol_x),
                                         capi.dptr(sol_u), capi.iptr(it), capi.iptr(status)))


    def wait(self):
        self._ck(self._L.srbd_wait(self._h))

    def mpc_run(self, A, Bm, b, x_start, steps):
        """Closed-loop batched MPC on the device (srbd_mpc_run): A, Bm, b = the plant, row-major numpy matrices, shared
        ([nx,nx], [nx,nu], [nx]) or one per robot ([B,nx,nx], ...).  Returns x_traj [steps+1,B,nx], u_traj [steps,B,nu],
        iter / status [steps,B]."""
        nx, nu, B = self.dims.nx, self.dims.nu, self.batch
        A, Bm, b = (np.asarray(v, dtype=np.float64) for v in (A, Bm, b))
        shared = A.ndim == 2
        Ac = np.ascontiguousarray(np.swapaxes(A, -1, -2))    # column-major blocks
        Bc = np.ascontiguousarray(np.swapaxes(Bm, -1, -2))
        bc = np.ascontiguousarray(b)
        xs = np.ascontiguousarray(x_start, dtype=np.float64).reshape(B, nx)
        xt, ut = np.zeros((steps + 1, B, nx)), np.zeros((steps, B, nu))
        it, st = np.zeros((steps, B), dtype=np.int32), np.zeros((steps, B), dtype=np.int32)
        self._ck(self._L.srbd_mpc_run(self._h, capi.dptr(Ac), capi.dptr(Bc), capi.dptr(bc), 1 if shared else 0,
                                      capi.dptr(xs), int(steps), capi.dptr(xt), capi.dptr(ut), capi.iptr(it), capi.iptr(st)))
        return xt, ut, it, st

    def fp64_peak(self):
        v = C.c_double()
        self._ck(self._L.srbd_fp64_peak(self._h, C.byref(v)))
        return float(v.value)
