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


def test_every_hpipm_symbol_is_exported(lib):
    """include/hpipm_b200_compat.h (SURVEY 8(b) "Option A"): the 54 HPIPM C symbols hpipm-cpp links against -- 7 dim, 10 qp,
    8 sol, 15 arg, 14 ws / solve / getters -- are exported by the same library, and the ABI-relevant struct sizes are
    those of the reference's vendored headers on LP64 (13 pointers + int + size_t, ...)."""
    hdr = open(os.path.join(ROOT, "include", "hpipm_b200_compat.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(d_ocp_qp_[a-z0-9_A-Z]+)\s*\(", hdr)))
    groups = {"dim": 0, "sol": 0, "arg": 0, "ipm": 0, "qp": 0}
    for n in names:
        assert hasattr(lib, n), n
        key = ("dim" if n.startswith("d_ocp_qp_dim_") else "sol" if n.startswith("d_ocp_qp_sol_") else
               "arg" if n.startswith("d_ocp_qp_ipm_arg_") else "ipm" if n.startswith("d_ocp_qp_ipm_") else "qp")
        groups[key] += 1
    assert groups == {"dim": 7, "qp": 10, "sol": 8, "arg": 15, "ipm": 14}, groups
    assert hasattr(lib, "hpipm_b200_last_error") and hasattr(lib, "hpipm_b200_pool_size")
    # no device needed for the host-side objects: memsize / create / setters / getters of the dimensions
    lib.d_ocp_qp_dim_memsize.restype = C.c_size_t
    n = lib.d_ocp_qp_dim_memsize(10)
    assert n >= 13 * 11 * 4

    class Dim(C.Structure):
        _fields_ = [(f, C.POINTER(C.c_int)) for f in ("nx", "nu", "nb", "nbx", "nbu", "ng", "ns", "nsbx", "nsbu", "nsg",
                                                      "nbxe", "nbue", "nge")] + [("N", C.c_int), ("memsize", C.c_size_t)]
    assert C.sizeof(Dim) == 13 * 8 + 8 + 8
    dim, mem = Dim(), C.create_string_buffer(n)
    lib.d_ocp_qp_dim_create(10, C.byref(dim), mem)
    arr = (C.c_int * 11)
    nx, nu, z = arr(*([12] * 11)), arr(*([4] * 10 + [0])), arr(*([0] * 11))
    nbx, nbu = arr(*([3] * 11)), arr(*([4] * 10 + [0]))
    lib.d_ocp_qp_dim_set_all(nx, nu, nbx, nbu, z, z, z, z, C.byref(dim))
    lib.d_ocp_qp_dim_set_nx(0, 0, C.byref(dim)); lib.d_ocp_qp_dim_set_nbx(0, 0, C.byref(dim)); lib.d_ocp_qp_dim_set_nsbx(0, 0, C.byref(dim))
    assert dim.N == 10 and dim.nx[0] == 0 and dim.nx[1] == 12 and dim.nb[0] == 4 and dim.nb[1] == 7 and dim.nu[10] == 0
    lib.d_ocp_qp_memsize.restype = C.c_size_t
    lib.d_ocp_qp_sol_memsize.restype = C.c_size_t
    # the quadcopter QP of the reference's test: 10 stages of A, B, Q, S, R (12 x 12, 12 x 4, ...) plus boxes
    assert lib.d_ocp_qp_memsize(C.byref(dim)) > 10 * (144 + 48 + 144 + 48 + 16) * 8
    assert lib.d_ocp_qp_sol_memsize(C.byref(dim)) > (11 * 12 * 2 + 10 * 4) * 8
    assert lib.hpipm_b200_pool_size() == 0   # nothing touched a device


def test_hpipm_struct_layout_matches_the_vendored_headers(tmp_path):
    """ABI of include/hpipm_b200_compat.h: sizes and member offsets of the five HPIPM structs (and the enum values) equal
    those of the reference's vendored headers (hpipm-cpp/include/include/hpipm_d_ocp_qp{_dim,,_sol,_ipm}.h, LP64).  The
    expected lines were produced in the build container by compiling tests/cpu_progs/hpipm_abi_layout.c with -DUSE_VENDORED
    against those headers (the reference tree is not available where the tests run)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "hpipm_abi_layout")
    r = subprocess.run([gcc, "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpu_progs", "hpipm_abi_layout.c"),
                        "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True).stdout.strip().splitlines()
    assert out == [
        "dim 120 N 104 memsize 112",
        "qp 120 BAbt 8 idxb 80 diag_H_flag 104 memsize 112",
        "sol 56 ux 8 misc 40 memsize 48",
        "arg 160 mu0 0 tau_min 72 iter_max 80 stat_max 84 pred_corr 88 warm_start 104 square_root_alg 108 lq_fact 112 "
        "split_step 132 t_lam_min 140 mode 144 memsize 152",
        "ws 304 qp_res 0 core_workspace 32 dim 40 stat 232 iter 256 stat_max 260 stat_m 264 status 272 valid_ric_p 292 memsize 296",
        "enums 0 1 2 3 | 0 1 2 3 4",
    ], out


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
