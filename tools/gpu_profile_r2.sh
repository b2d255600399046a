#!/bin/bash
# Round-2 evidence run (ON the GPU box, one GPU): launch list of the default bench and one full ncu capture each of the
# group-fused iteration kernel and the group-fused best-response kernel, each only after the same command ran clean.
mkdir -p gpurun_out
set -x
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --repeats 0 > gpurun_out/plain_r2.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --repeats 0 \
    > gpurun_out/ncu_launches_r2.log 2>&1
RLP_FUSED_DUMP=/root/repo/gpurun_out/cfr_fused_k.cu timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:cfr_fused_k -s 4 -c 1 -o gpurun_out/r2_fused_full python tools/fused_profile_target.py 13 6 > gpurun_out/ncu_full_r2.log 2>&1
timeout 120 python tools/br_probe.py 13 > gpurun_out/br_probe_r2.log 2>&1
RLP_FUSED_DUMP=/root/repo/gpurun_out/cfr_fused_k.cu timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:br_fused_k -s 2 -c 1 -o gpurun_out/r2_br_full python tools/br_probe.py 13 > gpurun_out/ncu_br_r2.log 2>&1
tail -2 gpurun_out/ncu_full_r2.log gpurun_out/ncu_br_r2.log
ls -la gpurun_out | tail -8
# sampled lanes (config 3): launch list of the level-by-level kernels and one full capture each of a mid-level expand / collect
timeout 120 python tools/lanes_launches.py 1 > gpurun_out/lanes_plain_r2.log 2>&1 || exit 1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_front --csv \
    --log-file gpurun_out/lanes_launches_r2.csv python tools/lanes_launches.py 1 > gpurun_out/ncu_lanes_launches_r2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_front_expand -s 20 -c 1 \
    -o gpurun_out/r2_lanes_expand python tools/lanes_launches.py 1 > gpurun_out/ncu_lanes_expand_r2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_front_collect -s 17 -c 1 \
    -o gpurun_out/r2_lanes_collect python tools/lanes_launches.py 1 > gpurun_out/ncu_lanes_collect_r2.log 2>&1
ls -la gpurun_out | tail -8
