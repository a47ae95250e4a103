#!/bin/bash
# Runs on the GPU box: GPU parity tests, smoke, both bench arms, secondary configs, the ncu launch list and one full
# capture per kernel.  Results land in gpurun_out/; tools/make_profiles.py turns them into profiles/.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 600 gpurun_out/bench_reference.json
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" != "noprof" ]; then
python tools/bench_configs.py 2>&1 | tail -1 > gpurun_out/bench_configs.json
python tools/bench_ragged.py 2>&1 | tail -1 > gpurun_out/bench_ragged.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
for k in k_verify k_verify_front k_x25519 k_sign k_genpub; do
ncu --set full --clock-control none --import-source on -k regex:${k}\$ -s 1 -c 1 -o gpurun_out/prof_${k} -f python tools/prof_driver.py > gpurun_out/prof_${k}.log 2>&1
done
fi
echo finished
