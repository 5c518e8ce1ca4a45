#!/bin/bash
# A/B of the pipelined solveBatch (hpipm-cpp facade): test_facades, then bench_facade at several chunk sizes / copy modes.
# Usage (GPU box): bash scripts/facade_ab.sh  -> gpurun_out/facade_ab.txt
mkdir -p gpurun_out
H=srbd-nmpc-solver_b200/host/tests
export LD_LIBRARY_PATH=srbd-nmpc-solver_b200:$LD_LIBRARY_PATH
{
  nproc
  timeout 600 $H/test_facades tests/golden/quadcopter_sol.txt | tail -12
  for nt in 1 0; do for c in 0 1024; do
    echo "== SRBD_FACADE_CHUNK=$c SRBD_FACADE_NT=$nt, 4096 QPs"
    SRBD_FACADE_NT=$nt SRBD_FACADE_CHUNK=$c timeout 600 $H/bench_facade 4096 3 | cut -c1-260
  done; done
  for nt in 1 0; do
    echo "== SRBD_FACADE_CHUNK=1024 SRBD_FACADE_NT=$nt, 16384 QPs"
    SRBD_FACADE_NT=$nt SRBD_FACADE_CHUNK=1024 timeout 900 $H/bench_facade 16384 2 | cut -c1-260
  done
} > gpurun_out/facade_ab.txt 2>&1
cat gpurun_out/facade_ab.txt
