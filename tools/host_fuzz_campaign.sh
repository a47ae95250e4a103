#!/bin/bash
# Long random walk over the C-ABI on the CUDA runtime simulator (tests/host_sim): seeds $1..$2, 100 steps each, under all four
# stream schedules, 1-4 simulated devices, every third seed with a 1 MB staging budget.  No GPU needed.  Prints failures.
cd "$(dirname "$0")/.."
make -s -C tests/host_sim || exit 1
fail=0
for sched in lazy others-first random eager; do
for seed in $(seq ${1:-1} ${2:-16}); do
  extra=""
  if [ $((seed % 3)) = 0 ]; then extra="EDDSA_B200_CHUNK_MB=1"; fi
  out=$(env CUDASIM_DEVICES=$((1 + seed % 4)) CUDASIM_SMS=1 CUDASIM_RESIDENT=32 CUDASIM_SCHEDULE=$sched $extra timeout 300 python tests/host_sim/sim_scenarios.py fuzz $seed 100 2>&1 | tail -4)
  case "$out" in *"OK fuzz"*) ;; *) echo "FAIL sched=$sched seed=$seed $extra"; echo "$out" | cut -c1-600; fail=1;; esac
done
done
echo "campaign done: seeds ${1:-1}..${2:-16} x 4 schedules, fail=$fail"
