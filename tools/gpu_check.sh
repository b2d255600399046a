#!/bin/bash
# One gpurun call's worth of validation + evidence, with every step under its own timeout so a stuck kernel cannot
# hold the box.  Run ON the GPU box from the repo root:
#
#   gpurun --timeout 600 -- 'bash tools/gpu_check.sh quick'      # tests + bench summary        (~2.5 GPU-min)
#   gpurun --timeout 900 -- 'bash tools/gpu_check.sh profile r2' # + launch list + ncu --set full (~5 GPU-min)
#
# Outputs land in gpurun_out/ (merged back by gpurun); copy what should be judged into profiles/ by hand.
# Multi-GPU: never under ncu, always `timeout`, e.g.
#   gpurun --gpus 2 --timeout 300 -- 'timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
#       --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5'
mode=${1:-quick}
tag=${2:-run}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 120 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_${tag}.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('value %.4g  steady %.2f us/iter  e2e %.4g (%.3f ms/step)  frac %.3f' % (
        d['value'], d['steady_state']['us_per_iteration'], d['e2e']['value'], d['e2e']['ms_per_step'], r['frac']))
    print('stages', r['stage_us_serialised'])
    print('micro_k', r.get('dominant_kernel'))
    print('cpu', d.get('cpu_baseline', {}).get('value'), 'clocks', d.get('clocks'))
except Exception as e:
    print('bench line unreadable:', e)
PY
[ "$mode" = profile ] || exit 0
# launch list (cold-cache, serialised: shares matter) -- only after the same command has exited 0 without ncu
timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_${tag}.log 2>&1 || exit 1
timeout 180 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_launches_${tag}.log 2>&1
# one full capture of each sweep kernel (skip the warm-up launches)
timeout 240 ncu --set full --clock-control none --import-source on \
    -k regex:"k_accumulate_rm4|micro_k|upper_k|k_fanin" -s 12 -c 6 -o gpurun_out/full_${tag} \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${tag}.log 2>&1
tail -1 gpurun_out/ncu_full_${tag}.log
ls -la gpurun_out | tail -8
