mkdir -p gpurun_out
python -m pytest tests -q -m gpu --maxfail=8 --tb=short > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
bash tools/run_variants.sh
python tools/bench_ragged.py 2>&1 | tail -1 | tee gpurun_out/bench_ragged.json
