#!/bin/bash
# Sanitizers over the REAL kernels on the SIMT emulator (tests/host_sim/simt_emul.h; lanes are OS threads).  No GPU.
#   tools/host_sanitize.sh thread     ThreadSanitizer: the shared-memory protocols of the kernels — tile sort, comb table staging, the
#                                     exchange areas of the tensor-core lookup, the staging of the verify loop, the permutation
#                                     counters — checked for data races (the CPU counterpart of compute-sanitizer racecheck)
#   tools/host_sanitize.sh address    AddressSanitizer + UBSan over the real_kernels scenario: every access of the kernels against the
#                                     exact bounds of the simulated device allocations (counterpart of memcheck), no undefined arithmetic
set -e
mode=${1:-thread}
cd "$(dirname "$0")/../tests/host_sim"
make -s -j4
D=$(mktemp -d)
if [ "$mode" = thread ]; then SAN="-fsanitize=thread"; RT=libtsan.so; SCEN=tsan_workload
else SAN="-fsanitize=address,undefined -fno-sanitize-recover=undefined"; RT=libasan.so; SCEN=real_kernels; fi
FL="-g -O1 $SAN -fPIC"
gcc $FL -std=gnu11 -fvisibility=hidden -DEDDSA_BUILD -I../../include -I/usr/local/cuda/include -c ../../libeddsa_b200/csrc/host.c -o $D/host.o &
g++ $FL -std=c++17 -Wno-unknown-pragmas -DCUDASIM_REAL_KERNELS -I/usr/local/cuda/include -c cudasim.cpp -o $D/cudasim.o &
for f in x25519 fixedbase verify; do
  g++ $FL -std=c++17 -fno-gnu-unique -Wno-unknown-pragmas -D__CUDA_ARCH__=1000 -D__CUDACC__ -Ifake_cuda -include simt_emul.h -c _ptx/libeddsa_b200/csrc/kernels_$f.cu.cpp -o $D/k_$f.o &
done
wait
g++ -shared $SAN -o $D/libeddsa_sim_kernels_san.so $D/cudasim.o $D/host.o $D/k_*.o -lpthread -Wl,--allow-multiple-definition
CUDASIM_SO=$D/libeddsa_sim_kernels_san.so LD_PRELOAD=$(gcc -print-file-name=$RT) CUDASIM_DEVICES=1 EDDSA_B200_VERIFY_WAVES=1 \
  TSAN_OPTIONS="report_signal_unsafe=0 exitcode=0 history_size=4" ASAN_OPTIONS=detect_leaks=0:detect_odr_violation=0 \
  python sim_scenarios.py $SCEN > $D/out.log 2>&1 || true
echo "$mode: $(grep -c -E 'WARNING: ThreadSanitizer|ERROR: AddressSanitizer|runtime error' $D/out.log) reports"
tail -2 $D/out.log
grep -A14 -E 'WARNING: ThreadSanitizer|ERROR: AddressSanitizer|runtime error' $D/out.log | head -60
rm -rf $D
