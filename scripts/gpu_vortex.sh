#!/usr/bin/env bash
# BASELINE configs[3] family on one GPU: isentropic vortex on jittered, id-shuffled triangulations (2 M and 16 M cells).
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
for n in ${SIZES:-1024 2828}; do
  MLB_PREP_TIMING=1 $T ${TMO:-1200} python bench.py --workload vortex --nx $n --ny $n --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_vortex_$n.json 2> gpurun_out/bench_vortex_$n.err
  echo "vortex $n rc=$?"; tail -c 2600 gpurun_out/bench_vortex_$n.json; tail -14 gpurun_out/bench_vortex_$n.err
  nvidia-smi --query-gpu=memory.used,memory.total --format=csv,noheader
done
