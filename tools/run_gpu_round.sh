#!/bin/bash
# Runs on the GPU box: GPU parity tests, smoke, both bench arms, secondary configs, the dynamic constant-time check, the
# ncu launch list and one full capture per kernel.  Results land in gpurun_out/; tools/make_profiles.py turns them into profiles/.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --maxfail=8 --tb=short --durations=12 > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 600 gpurun_out/bench_reference.json
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ -d libeddsa_b200/variants ]; then bash tools/run_variants.sh; fi
if [ "$1" != "noprof" ]; then
python tools/bench_configs.py 2>&1 | tail -1 > gpurun_out/bench_configs.json
python tools/bench_ragged.py 2>&1 | tail -1 > gpurun_out/bench_ragged.json
python tools/ct_dynamic.py --json gpurun_out/ct_dynamic.json 2>&1 | tail -20 > gpurun_out/ct_dynamic.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-inproc > gpurun_out/bench_under_ncu.log 2>&1
for k in ${KERNELS:-k_verify k_verify_points k_verify_scalars k_x25519 k_comb k_sign_nonce k_sign_finish k_expand_key}; do
ncu --set full --clock-control none --import-source on -k regex:${k}\$ -s 1 -c 1 -o gpurun_out/prof_${k} -f python tools/prof_driver.py > gpurun_out/prof_${k}.log 2>&1
done
fi
echo finished
