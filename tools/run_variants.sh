#!/bin/bash
# GPU box: time the shipped library and every tuning variant under libeddsa_b200/variants/
( python tools/tune_bench.py 2>&1 | tail -1
for so in libeddsa_b200/variants/lib_*.so; do
  LIBEDDSA_B200_SO=$PWD/$so python tools/tune_bench.py 2>&1 | tail -1
done ) | tee gpurun_out/variants.jsonl
