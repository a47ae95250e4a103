#!/bin/bash
# usage: tools/sass_size.sh <object-or-so> ...   — SASS instruction count and code bytes per kernel
for o in "$@"; do
cuobjdump -sass "$o" 2>/dev/null | grep -E "Function :|^\s+/\*[0-9a-f]{4,}\*/" | awk '/Function :/{fn=$3; next} {t[fn]++} END{for (f in t) printf "%7d instr %7.1f KB  %s\n", t[f], t[f]*16/1024, f}' | sed -E 's/_ZN[0-9]+_GLOBAL__N__[0-9a-f]+_[0-9]+_kernels_[a-z0-9]+_cu_[0-9a-f]+[0-9]+//' | sort -k5
done
