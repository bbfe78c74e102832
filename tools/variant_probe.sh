#!/bin/bash
# perf_probe.py (peak config, identical subjects) for the shipped library and every kernel-variant library under
# build/variants (development aid). usage: tools/variant_probe.sh [extra env assignments...]
for lib in cudasw4_b200/libsw4b200.so build/variants/*.so; do
  echo "== $lib $*"
  env SW4B200_LIB=$PWD/$lib "$@" timeout 300 python tools/perf_probe.py 1000000 256 pseudo 2>&1 | grep -E "q=  144|q=  375|q= 1000|q= 2504|q= 5478|total"
done
