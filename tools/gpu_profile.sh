#!/bin/bash
# usage (on the GPU box): bash tools/gpu_profile.sh <tag>
#   1. the default bench (value, e2e, cpu_baseline)                 -> gpurun_out/bench_<tag>_full.json
#   2. ncu launch list of the same device-resident command          -> gpurun_out/launches_<tag>.csv
#   3. one ncu --set full capture of every kernel, 64-frame launch  -> gpurun_out/prof_<tag>.ncu-rep
# Nothing printed under ncu is a bench value.
T=$1
if [ "$2" != nobench ]; then python bench.py > gpurun_out/bench_${T}_full.json 2> gpurun_out/bench_${T}_full.err; tail -c 600 gpurun_out/bench_${T}_full.json; fi
K='regex:mbvar_kernel|fdct_quant_kernel|huffman_kernel|entropy_walk_kernel|scan_place_kernel|stuff_kernel|pack_offsets_kernel|pack_kernel|convert_pad'
ncu -k "$K" --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --no-e2e --no-cpu-baseline --no-overlap --no-extras --no-other-configs --sustain-seconds 0 --parity-frames 0 > gpurun_out/ncu_launches_$T.log 2>&1
ncu -k "$K" --set full --clock-control none --import-source on -s 7 -c 7 -o gpurun_out/prof_$T -f \
    python bench.py --no-e2e --no-cpu-baseline --no-overlap --no-extras --no-other-configs --sustain-seconds 0 --parity-frames 0 --frames 64 --steps 1 --warmup 1 > gpurun_out/ncu_$T.log 2>&1
tail -3 gpurun_out/ncu_$T.log | cut -c1-200
