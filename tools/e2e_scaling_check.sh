#!/bin/bash
# usage (on an 8-GPU box): bash tools/e2e_scaling_check.sh <tag>   -> the bench's host-to-host pass at 8 / 4 / 2 ranks with the
# probe-driven device plan and link-rate weighted shards (bench.py device_plan), one summary line each
T=$1; O=gpurun_out/$T; mkdir -p $O
B="--steps 5 --warmup 3 --no-other-configs --no-extras --no-cpu-baseline --sustain-seconds 0 --no-overlap --parity-frames 4"
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n $B > $O/bench_n$n.json 2> $O/bench_n$n.err
  python -c "
import json;d=json.load(open('$O/bench_n$n.json'));e=d['e2e'];print('N=$n value',round(d['value']),'e2e',round(e['value']),'h2d_gbs',round(e['h2d_gbs'],1),'per rank',e['per_rank']['h2d_gbs'],'frames',e['per_rank']['frames_per_step'],'gpus',e['per_rank']['gpu'],'parity',e.get('parity_sampled',{}).get('ok'), d['config']['parity_sampled']['ok'])" || tail -5 $O/bench_n$n.err
done
