#!/bin/bash
# The measurement bundle committed under profiles/ at the end of a round (tag = $1, e.g. r2f).
T=${1:-r2f}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_ref.err
python scripts/bench_configs.py > gpurun_out/${T}_configs_1_2_4_5.json 2> gpurun_out/${T}_configs.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-facade --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
tail -c 600 gpurun_out/${T}_bench.json; echo; tail -c 1500 gpurun_out/${T}_configs_1_2_4_5.json
