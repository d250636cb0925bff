#!/usr/bin/env bash
# Round 2, visit A (one GPU): full parity suite incl. the new multi-tile / benchmarked-config / drop-in cases, face-flux
# kernel A/B (register budgets), first-order bench, drift study.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
$T 2400 python -m pytest tests -m gpu -q -rs --durations=15 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -5
$T 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench.json'));print(d['value'], d['e2e']['value'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})"
for alt in ${ALTS:-}; do
  if [ -f mallard_b200/libmallard_b200_$alt.so ]; then
    MLB_LIB=$PWD/mallard_b200/libmallard_b200_$alt.so $T 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$alt.json 2> gpurun_out/bench_$alt.err; echo "bench $alt rc=$?"
    python -c "import json;d=json.load(open('gpurun_out/bench_$alt.json'));print('$alt', d['value'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})"
    MLB_LIB=$PWD/mallard_b200/libmallard_b200_$alt.so $T 600 python bench.py --recon FO --nx 4096 --ny 4096 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_fo_$alt.json 2> gpurun_out/bench_fo_$alt.err
    python -c "import json;d=json.load(open('gpurun_out/bench_fo_$alt.json'));print('$alt FO', d['value'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})"
  fi
done
$T 600 python bench.py --recon FO --nx 4096 --ny 4096 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_fo.json 2> gpurun_out/bench_fo.err; echo "bench FO rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_fo.json'));print('FO', d['value'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})"
$T 900 python scripts/drift_study.py 2000 100 96 > gpurun_out/drift.txt 2> gpurun_out/drift.err; echo "drift rc=$?"; tail -8 gpurun_out/drift.txt
if [ "${NCU:-0}" = 1 ]; then
$T 900 ncu --set full --clock-control none --import-source on -k regex:'face_flux|gather_stage' -s 4 -c 4 -f -o gpurun_out/prof_flux \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_flux.log 2>&1
fi
ls -la gpurun_out
