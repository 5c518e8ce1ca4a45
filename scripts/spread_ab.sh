#!/bin/bash
# A/B of the even spreading of sparse batches (148 <= B < 1776 resident warps): previous library / SRBD_K3_SPREAD=0 / new.
P=$PWD/srbd-nmpc-solver_b200/libsrbd_b200_prev.so
{
for B in 148 296 512 1024 1500 1776; do
  echo "== B=$B"
  [ -f $P ] && SRBD_LIB=$P python scripts/run_k3.py $B 5 | sed 's/^/prev    /'
  SRBD_K3_SPREAD=0 python scripts/run_k3.py $B 5 | sed 's/^/packed  /'
  python scripts/run_k3.py $B 5 | sed 's/^/spread  /'
done
echo "== burst 16384"
[ -f $P ] && SRBD_LIB=$P python scripts/run_k3.py 16384 3 | sed 's/^/prev    /'
python scripts/run_k3.py 16384 3 | sed 's/^/new     /'
[ -f $P ] && SRBD_LIB=$P python scripts/run_k3.py 16384 3 | sed 's/^/prev    /'
python scripts/run_k3.py 16384 3 | sed 's/^/new     /'
echo "== sustained"
[ -f $P ] && SRBD_LIB=$P python scripts/run_k3_sustained.py | tail -2 | sed 's/^/prev    /'
python scripts/run_k3_sustained.py | tail -2 | sed 's/^/new     /'
} > gpurun_out/spread_ab.txt 2>&1
cat gpurun_out/spread_ab.txt
