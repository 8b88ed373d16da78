#!/bin/bash
# static SASS size, registers and spills of kernels matching a pattern in the built library (proxy for loop instruction counts)
lib=${2:-/root/repo/clair_torch_b200/lib/libclair_b200.so}
cuobjdump -sass -fun "$1" $lib 2>/dev/null | awk '/Function :/{name=$3} /^\s+\/\*[0-9a-f]{4}\*\//{n[name]++} END{for(k in n) print n[k], k}'
cuobjdump -res-usage $lib 2>/dev/null | grep -A1 "$1" | grep -o "REG:[0-9]*\|STACK:[0-9]*" | paste - -
