#!/bin/bash
# ThreadSanitizer over the REAL kernels on the SIMT emulator (tests/host_sim/simt_emul.h: lanes are OS threads, so the shared-memory
# protocols of the kernels — tile sort, comb table staging, the exchange areas of the tensor-core lookup, the staging of the verify
# loop, the permutation counters — are checked for data races the way compute-sanitizer racecheck checks them on the GPU).  No GPU.
set -e
cd "$(dirname "$0")/../tests/host_sim"
make -s -j4
D=$(mktemp -d)
TS="-g -O1 -fsanitize=thread -fPIC"
gcc $TS -std=gnu11 -fvisibility=hidden -DEDDSA_BUILD -I../../include -I/usr/local/cuda/include -c ../../libeddsa_b200/csrc/host.c -o $D/host.o &
g++ $TS -std=c++17 -Wno-unknown-pragmas -DCUDASIM_REAL_KERNELS -I/usr/local/cuda/include -c cudasim.cpp -o $D/cudasim.o &
for f in x25519 fixedbase verify; do
  g++ $TS -std=c++17 -fno-gnu-unique -Wno-unknown-pragmas -D__CUDA_ARCH__=1000 -D__CUDACC__ -Ifake_cuda -include simt_emul.h -c _ptx/libeddsa_b200/csrc/kernels_$f.cu.cpp -o $D/k_$f.o &
done
wait
g++ -shared -fsanitize=thread -o $D/libsim.so $D/cudasim.o $D/host.o $D/k_*.o -lpthread -Wl,--allow-multiple-definition
CUDASIM_SO=$D/libsim.so LD_PRELOAD=$(gcc -print-file-name=libtsan.so) TSAN_OPTIONS="report_signal_unsafe=0 exitcode=0 history_size=4" CUDASIM_DEVICES=1 \
  python sim_scenarios.py tsan_workload > $D/out.log 2>&1 || true
echo "ThreadSanitizer warnings: $(grep -c 'WARNING: ThreadSanitizer' $D/out.log)"
tail -2 $D/out.log
grep -A14 'WARNING: ThreadSanitizer' $D/out.log | head -60
rm -rf $D
