#!/bin/bash
# GPU box: one ncu --set full capture per kernel on the short driver (2^18 ops), reports into gpurun_out/
mkdir -p gpurun_out
for k in ${KERNELS:-k_verify k_verify_front k_x25519 k_sign k_genpub}; do
ncu --set full --clock-control none --import-source on -k regex:${k}\$ -s 1 -c 1 -o gpurun_out/prof_${k} -f python tools/prof_driver.py > gpurun_out/prof_${k}.log 2>&1
done
echo finished
