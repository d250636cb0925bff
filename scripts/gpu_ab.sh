#!/usr/bin/env bash
# A/B visit: parity tests, bench with the main library, bench with the alternative build (MLB_LIB), host-table timing.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
$T 1200 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
MLB_PREP_TIMING=1 $T 600 python bench.py --steps 20 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for alt in ${ALTS:-alt}; do
  if [ -f mallard_b200/libmallard_b200_$alt.so ]; then
    MLB_LIB=$PWD/mallard_b200/libmallard_b200_$alt.so $T 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$alt.json 2> gpurun_out/bench_$alt.err; echo "bench $alt rc=$?"
    python -c "import json;d=json.load(open('gpurun_out/bench_$alt.json'));print(d['value'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})"
  fi
done
if [ "${HOSTTAB:-0}" = 1 ]; then
  MLB_HOST_TABLES=1 MLB_PREP_TIMING=1 $T 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_hosttab.json 2> gpurun_out/bench_hosttab.err; echo "bench hosttab rc=$?"
  tail -3 gpurun_out/bench_hosttab.err
fi
if [ "${NCU:-0}" = 1 ]; then
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
$T 900 ncu --set full --clock-control none --import-source on -k regex:'teno_|face_flux|gather_stage|cfl_kernel' -s 9 -c 8 -f -o gpurun_out/prof \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi
ls -la gpurun_out
