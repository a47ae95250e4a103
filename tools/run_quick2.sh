mkdir -p gpurun_out
python -m pytest tests -q -m gpu --maxfail=5 --tb=short > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
a=d['also']
print("value", d['value']/1e6, "e2e", d['e2e']['value']/1e6, "pageable", a['e2e_pageable']['value']/1e6, "1kb", a['e2e_1kb']['value']/1e6)
print({k: round(v['value']/1e6,1) for k,v in a.items() if 'roofline_frac_executed' in v})
print(a['inproc']['config4_keygen'])
c5=a['inproc']['config5_verify_1kb_10pct_mutated']; print(c5['ops_per_s']/1e6, c5['parity_sample'], c5['sign_1kb_ops_per_s_e2e']/1e6)
PY
