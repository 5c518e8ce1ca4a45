"""CPU tests (-m "not gpu"): the C-ABI library loads, exports every symbol include/srbd_b200.h declares,
its defaults agree with the oracle's, and the product path fails loudly without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(pkg):
    import __graft_entry__ as g
    g.build()
    return pkg.capi.lib()


def test_every_declared_symbol_is_exported(pkg, lib):
    hdr = open(os.path.join(ROOT, "include", "srbd_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(srbd_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(pkg.capi._EXPORTS)


def test_defaults_match_oracle(pkg, orc, lib):
    mp, om = pkg.default_model_params(20), orc.model_params(20)
    assert bytes(mp) == bytes(om)
    ia, oa = pkg.default_ipm_args(), orc.ipm_args()
    assert bytes(ia) == bytes(oa)
    d = pkg.capi.QpDims(10, 12, 4, 3, 4, 0, 0)
    assert lib.srbd_qp_nct(C.byref(d)) == orc.lib().orc_qp_nct(C.byref(d)) == pkg.capi.qp_nct(d) == 2 * (10 * 4 + 10 * 3)


def test_struct_sizes_match_header(pkg):
    # srbd_model_params: 2+9+6+18+5+3+12+12+1+2+1 doubles
    assert C.sizeof(pkg.capi.ModelParams) == 71 * 8
    assert C.sizeof(pkg.capi.QpDims) == 7 * 4
    assert C.sizeof(pkg.capi.QpHost) == 32 * 8
    assert C.sizeof(pkg.capi.BatchStats) == (2 + 64 + 5) * 8 + 4 * 8


def test_no_cpu_fallback(pkg, lib):
    """Without a usable CUDA device the product path must fail loudly (no oracle / CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from srbd_nmpc_solver_b200.binding import SrbdError
    with pytest.raises(SrbdError, match="srbd_ctx_create failed"):
        pkg.Context(4)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pk = os.path.join(ROOT, "srbd-nmpc-solver_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "__init__.py" and False, os.path.join(dirpath, f)


def test_workload_is_shard_invariant(pkg):
    """counter-based generator: a shard is identical regardless of how the batch is split over ranks."""
    full = pkg.workload.srbd_batch(64, N=20, contact_mode="gait")
    a = pkg.workload.srbd_batch(32, N=20, contact_mode="gait", start=0)
    b = pkg.workload.srbd_batch(32, N=20, contact_mode="gait", start=32)
    for k in full:
        assert np.array_equal(full[k], np.concatenate([a[k], b[k]]))
    assert full["contact"].sum(axis=2).min() >= 1
    assert np.linalg.norm(full["x0"][:, :3], axis=1).min() >= 1e-2 - 1e-15


def test_ipm_args_set_mode(pkg, lib):
    """srbd_ipm_args_set_mode: the hidden constants of d_ocp_qp_ipm_arg_set_default(mode) this implementation honours
    (hpipm-cpp's HpipmMode: 0 SpeedAbs, 1 Speed, 2 Balance, 3 Robust); the public fields are left alone."""
    want = {0: (0, 0, 0), 1: (1, 0, 0), 2: (1, 0, 2), 3: (1, 0, 4)}
    for mode, (cpc, ip, ic) in want.items():
        a = pkg.default_ipm_args(iter_max=77, tol_stat=3e-5)
        assert lib.srbd_ipm_args_set_mode(C.byref(a), mode) == 0
        assert (a.cond_pred_corr, a.itref_pred_max, a.itref_corr_max) == (cpc, ip, ic)
        assert a.iter_max == 77 and a.tol_stat == 3e-5 and a.itref_abs == 1.0 and a.itref_rel == 1e-3
    a = pkg.default_ipm_args()
    assert lib.srbd_ipm_args_set_mode(C.byref(a), 7) != 0


def test_workload_spread(pkg):
    """spread pulls the start states towards the reference's upright pose; spread = 1 is the config-2/3 generator itself."""
    a = pkg.workload.srbd_batch(16, N=5, contact_mode="gait")
    b = pkg.workload.srbd_batch(16, N=5, contact_mode="gait", spread=1.0)
    c = pkg.workload.srbd_batch(16, N=5, contact_mode="gait", spread=0.25)
    for k in a:
        assert np.array_equal(a[k], b[k])
    up = np.zeros(12); up[8] = 1.0
    assert np.allclose(c["x0"] - up, 0.25 * (a["x0"] - up)) and np.array_equal(c["contact"], a["contact"])
    assert np.allclose(c["xref"][:, 0, [2, 6, 7]], 0.25 * a["xref"][:, 0, [2, 6, 7]])
