"""Generates tests/golden/quadcopter_sol.npz from the reference's own golden vectors.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Source: hpipm-cpp/test/sol0.txt ... sol14.txt (OSQP solutions at eps 1e-10 of the closed-loop
quadcopter MPC, generated upstream by hpipm-cpp/test/osqp_gen.py), checked by the reference at 1e-9 in
hpipm-cpp/test/ocp_qp_ipm_solver.cpp:300-314.  Each vector is [x0..xN (11*12), u0..uN-1 (10*4)].
"""
import os

import numpy as np

REF = "/root/reference/hpipm-cpp/test"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    sols = np.stack([np.loadtxt(os.path.join(REF, f"sol{t}.txt")) for t in range(15)])
    assert sols.shape == (15, 172)
    np.savez(os.path.join(HERE, "quadcopter_sol.npz"), sol=sols)
    np.savetxt(os.path.join(HERE, "quadcopter_sol.txt"), sols, fmt="%.18e")  # same data for the C++ facade test
    print("wrote", sols.shape)
