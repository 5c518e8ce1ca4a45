#!/bin/bash
# A/B of an environment knob of the library: bash scripts/env_ab.sh VAR=VALUE  (burst 16384 QPs and sustained 16 x 65536)
{
for rep in 1 2; do
  python scripts/run_k3.py 16384 3 | sed 's/^/default  /'
  env "$1" python scripts/run_k3.py 16384 3 | sed "s/^/$1  /"
done
python scripts/run_k3_sustained.py | tail -1 | sed 's/^/default  /'
env "$1" python scripts/run_k3_sustained.py | tail -1 | sed "s/^/$1  /"
python scripts/run_k3_sustained.py | tail -1 | sed 's/^/default  /'
env "$1" python scripts/run_k3_sustained.py | tail -1 | sed "s/^/$1  /"
} > gpurun_out/env_ab.txt 2>&1
cat gpurun_out/env_ab.txt
