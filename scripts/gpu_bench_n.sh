#!/usr/bin/env bash
# One bench.py line on N GPUs exactly as the driver launches it (N=1: plain python), strong-scaling records included.
set -u
mkdir -p gpurun_out
N=${NGPU:-8}
T="timeout -s KILL"
nvidia-smi -L | head -8; nproc; free -g | head -2
if [ "$N" = 1 ]; then
  MLB_PREP_TIMING=1 $T 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ${BENCH_ARGS:-} > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench N=1 rc=$?"
else
  MLB_PREP_TIMING=1 $T 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 ${BENCH_ARGS:-} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
fi
python - <<PY
import json
for l in open('gpurun_out/bench_n$N.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print('main', d['value'], d['ms_per_step'], 'e2e', d['e2e'] and d['e2e']['value'], 'roof', d['roofline'] and d['roofline']['frac'])
        for r in d.get('strong') or []:
            print('strong', {k: r.get(k) for k in ('workload', 'value', 'ms_per_step', 'efficiency', 'setup_seconds', 'mesh_seconds', 'preprocess_seconds', 'host_rss_gb_per_rank_max', 'skipped', 'error')} if 'value' in r else r)
            if 'roofline' in r: print('   roof', r['roofline']['frac'], r['roofline']['ms_per_stage_slowest_rank'], r.get('halo'))
PY
tail -5 gpurun_out/bench_n$N.err
