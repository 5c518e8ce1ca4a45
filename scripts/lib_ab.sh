#!/bin/bash
# A/B of library builds (build/libsrbd_NAME.so, scripts/build_variant.sh): burst (16384 QPs) and sustained (16 x 65536) K3 rates.
# Usage: bash scripts/lib_ab.sh NAME [NAME ...]   -> gpurun_out/lib_ab.txt
{
for rep in 1 2; do
  python scripts/run_k3.py 16384 3 | sed 's/^/default  /'
  for n in "$@"; do SRBD_LIB=$PWD/build/libsrbd_$n.so python scripts/run_k3.py 16384 3 | sed "s/^/$n  /"; done
done
python scripts/run_k3_sustained.py | tail -1 | sed 's/^/default  /'
for n in "$@"; do SRBD_LIB=$PWD/build/libsrbd_$n.so python scripts/run_k3_sustained.py | tail -1 | sed "s/^/$n  /"; done
python scripts/run_k3_sustained.py | tail -1 | sed 's/^/default  /'
} > gpurun_out/lib_ab.txt 2>&1
cat gpurun_out/lib_ab.txt
