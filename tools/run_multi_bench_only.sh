#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
for rep in 1 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$rep bench.py --gpus $N --steps 10 --warmup 3 --no-inproc > gpurun_out/bench_n${N}_rep$rep.json 2> gpurun_out/bench_n${N}_rep$rep.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n${N}_rep$rep.json').read().strip().splitlines()[-1])
a=d['also']; print("rep$rep", round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), d['e2e']['ms_per_step_by_rank'], "pageable", a['e2e_pageable'], "1kb", round(a['e2e_1kb']['value']/1e6,1))
PY
done
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|^CPU\(s\)"
