#!/bin/bash
# usage (on the GPU box): bash tools/ab.sh <tag> <variant> [<variant> ...]  -- the default device-resident bench once per
# library variant (lib/variants/<variant>.so, "default" = the in-tree library), one summary line each; extra bench flags in $AB_FLAGS
T=$1; shift
for v in "$@"; do
  if [ "$v" = default ]; then unset H2J_B200_LIB; else export H2J_B200_LIB=$PWD/h264-h265-to-jpeg_b200/lib/variants/$v.so; fi
  timeout 240 python bench.py --no-e2e --no-cpu-baseline --no-overlap --no-extras --no-other-configs --sustain-seconds 0 $AB_FLAGS > gpurun_out/ab_${T}_$v.json 2> gpurun_out/ab_${T}_$v.err
  python -c "
import json,sys;d=json.load(open(sys.argv[1]));print(sys.argv[2], round(d['value']), d['config'].get('parity_sampled',{}).get('ok'), {k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})" gpurun_out/ab_${T}_$v.json $v || tail -3 gpurun_out/ab_${T}_$v.err
done
