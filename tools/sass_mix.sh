#!/bin/bash
# usage: tools/sass_mix.sh <object-or-so> [min-count]   — per-kernel SASS opcode histogram
cuobjdump -sass "$1" 2>/dev/null | grep -E "Function :|^\s+/\*[0-9a-f]{4,}\*/" | awk -v min="${2:-100}" '/Function :/{fn=$3; next} {op=$2; if (op ~ /^@/) op=$3; gsub(/;/,"",op); c[fn" "op]++; t[fn]++} END{for(k in c) if (c[k]>=min) print c[k], k; for (f in t) print t[f], f, "TOTAL"}' | sort -k2,2 -k1,1nr | sed 's/_ZN.*_GLOBAL__N__[0-9a-f_]*kernels_[a-z0-9]*_cu_[0-9a-f]*//'
