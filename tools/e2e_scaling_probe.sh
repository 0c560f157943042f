#!/bin/bash
# usage (on an 8-GPU box): bash tools/e2e_scaling_probe.sh <tag>
# Raw concurrent H2D probe (tools/microbench/pcie_probe), then the bench's host-to-host pass at 2 / 4 / 8 ranks with the
# ranks on neighbouring GPUs (H2J_BENCH_SPREAD=0) and spread over the box (default) -> gpurun_out/<tag>/
T=$1; O=gpurun_out/$T; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1; lscpu > $O/lscpu.txt 2>&1; numactl -H > $O/numa.txt 2>&1 || true
./tools/microbench/pcie_probe 199 30 > $O/pcie_probe.jsonl 2> $O/pcie_probe.err
B="--steps 5 --warmup 3 --no-other-configs --no-extras --no-cpu-baseline --sustain-seconds 0 --no-overlap --parity-frames 4"
run() { # n spread
  H2J_BENCH_SPREAD=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1 + 10 * $2)) bench.py --gpus $1 $B > $O/bench_n$1_spread$2.json 2> $O/bench_n$1_spread$2.err
  python -c "
import json,sys;d=json.load(open('$O/bench_n$1_spread$2.json'));e=d['e2e'];print('N=$1 spread=$2 value',round(d['value']),'e2e',round(e['value']),'h2d_gbs',round(e['h2d_gbs'],1),'per rank',e['per_rank']['h2d_gbs'],'gpus',e['per_rank']['gpu'])"
}
run 8 1; run 4 0; run 4 1; run 2 0; run 2 1
