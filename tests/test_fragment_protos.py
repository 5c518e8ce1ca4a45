"""CPU checks of the fragment bookkeeping of the DMMA kernels (csrc/ipm_srbd.cuh).

scripts/proto_dmma_factor.py and scripts/proto_dmma_vec.py emulate, lane by lane, the mma.sync m8n8k4 f64 fragment
layout, the row-permuted B operands, the blocked 4-column Cholesky panels with the appended identity rows, the
factor-panel export and the gather offsets of the vector sweeps (blocked triangular solves, G / G^T / P fragments from
the BLASFEO panel-major records), and compare against plain numpy.  The CUDA code is a transcription of these
prototypes; the GPU parity tests (tests/test_gpu_parity.py) check the transcription against the oracle.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
import proto_dmma_factor as pf  # noqa: E402
import proto_dmma_vec as pv  # noqa: E402


def test_dmma_emulator_is_a_matrix_product():
    rng = np.random.default_rng(3)
    A, B, C = rng.normal(size=(8, 4)), rng.normal(size=(4, 8)), rng.normal(size=(8, 8))
    a = A[pf.R_, pf.T_]
    b = B[pf.T_, pf.R_]
    d0, d1 = pf.dmma(C[pf.R_, 2 * pf.T_], C[pf.R_, 2 * pf.T_ + 1], a, b)
    D = A @ B + C
    assert np.allclose(d0, D[pf.R_, 2 * pf.T_]) and np.allclose(d1, D[pf.R_, 2 * pf.T_ + 1])


def test_permuted_b_operand_yields_panel_fragments():
    """with B rows permuted by pi the accumulator halves are the 4-column panels 2J and 2J+1 of the product"""
    rng = np.random.default_rng(4)
    X, Y = rng.normal(size=(8, 4)), rng.normal(size=(8, 4))
    c0, c1 = pf.dmma(np.zeros(32), np.zeros(32), pf.frag(X, 0, 0), pf.pfrag(Y, 0, 0))
    Pm = X @ Y.T
    assert np.allclose(c0, Pm[pf.R_, pf.T_]) and np.allclose(c1, Pm[pf.R_, 4 + pf.T_])


def test_factor_stage_prototype(capsys):
    pf.main()
    assert "rel err" in capsys.readouterr().out


def test_vector_sweep_prototype(capsys):
    pv.main()
    assert "OK" in capsys.readouterr().out
