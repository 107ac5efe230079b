#!/bin/bash
# code size (KB) per kernel / device function in a cubin-bearing .so
cuobjdump -sass "$1" 2>/dev/null | awk '/Function :/ {name=$3} /^ +\/\*[0-9a-f]+\*\/ / {cnt[name]++} END {for (n in cnt) printf "%8.1f KB  %s\n", cnt[n]*16/1024, n}' | sort -rn | sed -E 's/_ZN[0-9a-zA-Z_]*GLOBAL__N__[0-9a-f]+_[0-9]+_([a-z]+)_cu_[0-9a-f]+/\1::/' | c++filt 2>/dev/null | cut -c1-150
