#!/usr/bin/env bash
# A/B visit: parity tests, bench with the main library, bench under environment switches (ENVS="MLB_X=1 MLB_Y=2"),
# bench with alternative builds (ALTS="name ..." -> mallard_b200/libmallard_b200_<name>.so), optional ncu of the main build.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
short() { python -c "import json,sys;d=json.load(open(sys.argv[1]));print(sys.argv[1], round(d['value']/1e6,1), 'M/s', {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})" "$1"; }
if [ "${PYTEST:-1}" = 1 ]; then
$T 1200 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
fi
MLB_PREP_TIMING=1 $T 600 python bench.py --steps 20 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
i=0
for e in ${ENVS:-}; do
  i=$((i+1))
  env $e $T 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} > gpurun_out/bench_env$i.json 2> gpurun_out/bench_env$i.err; echo "bench $e rc=$?"
  short gpurun_out/bench_env$i.json
done
for alt in ${ALTS:-}; do
  if [ -f mallard_b200/libmallard_b200_$alt.so ]; then
    MLB_LIB=$PWD/mallard_b200/libmallard_b200_$alt.so $T 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} > gpurun_out/bench_$alt.json 2> gpurun_out/bench_$alt.err; echo "bench $alt rc=$?"
    short gpurun_out/bench_$alt.json
  fi
done
if [ "${NCU:-0}" = 1 ]; then
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/ncu_launches.log 2>&1
$T 900 ncu --set full --clock-control none --import-source on -k regex:'teno_stream|face_flux|gather_stage|cfl_kernel' -s 8 -c 8 -f -o gpurun_out/prof \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/ncu_full.log 2>&1
fi
ls -la gpurun_out
