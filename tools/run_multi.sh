#!/bin/bash
# GPU box with N GPUs (gpurun --gpus N): the multi-device tests and bench.py under torchrun at N
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -q -m gpu --tb=short -k "shards_over or current_device or concurrent" > gpurun_out/pytest_multi_n$N.log 2>&1; tail -6 gpurun_out/pytest_multi_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_reference_n$N.json 2>> gpurun_out/bench_n$N.err; tail -c 400 gpurun_out/bench_reference_n$N.json
echo finished
