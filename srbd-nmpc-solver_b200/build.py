"""Builds libsrbd_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsrbd_b200.so")
SOURCES = ["capi.cu", "hpipm_compat.cu"]
HEADERS = ["layout.cuh", "ipm_solve.cuh", "ipm_srbd.cuh", "srbd_model.cuh", "aux_kernels.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "srbd_b200.h"))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "hpipm_b200_compat.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed")
    return OUT


HOST_TEST = os.path.join(HERE, "host", "tests", "test_facades")
HOST_BENCH = os.path.join(HERE, "host", "tests", "bench_facade")
HOST_COMPAT = os.path.join(HERE, "host", "tests", "test_hpipm_compat")


def build_host_tests(force=False):
    """C++ host facades (host/*.hpp) + the reference's own host tests re-written against them."""
    src = os.path.join(HERE, "host", "tests", "test_facades.cpp")
    deps = [src] + [os.path.join(HERE, "host", f) for f in ("NMPC_solver.hpp", "SRBD_model.hpp", "eigen_shim.hpp",
                                                             os.path.join("hpipm-cpp", "hpipm-cpp.hpp"))]
    csrc = os.path.join(HERE, "host", "tests", "test_hpipm_compat.cpp")
    if (not force and os.path.exists(HOST_TEST) and
            os.path.exists(HOST_BENCH) and os.path.exists(HOST_COMPAT) and
            all(min(os.path.getmtime(HOST_TEST), os.path.getmtime(HOST_BENCH), os.path.getmtime(HOST_COMPAT)) >= os.path.getmtime(d)
                for d in deps + [OUT, csrc, os.path.join(HERE, "host", "tests", "bench_facade.cpp"),
                                 os.path.join(os.path.dirname(HERE), "include", "hpipm_b200_compat.h")])):
        return HOST_TEST
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-std=c++17", "-O2", "-Wall", "-Wextra", src, "-o", HOST_TEST, "-L" + HERE, "-lsrbd_b200",
           "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/../..", "-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed for the host facade test")
    if r.stderr.strip():
        sys.stderr.write(r.stderr)
    # the facade benchmark (bench.py: `facade_e2e`)
    bsrc = os.path.join(HERE, "host", "tests", "bench_facade.cpp")
    cmd = [gxx, "-std=c++17", "-O2", "-Wall", "-Wextra", bsrc, "-o", HOST_BENCH, "-L" + HERE, "-lsrbd_b200",
           "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/../..", "-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed for the facade benchmark")
    # the HPIPM C symbols (include/hpipm_b200_compat.h) driven in the reference wrapper's call order
    cmd = [gxx, "-std=c++17", "-O2", "-Wall", "-Wextra", csrc, "-o", HOST_COMPAT, "-L" + HERE, "-lsrbd_b200",
           "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/../..", "-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed for the HPIPM-symbol test")
    if r.stderr.strip():
        sys.stderr.write(r.stderr)
    return HOST_TEST


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    build_host_tests(force=True)
    print(OUT)
