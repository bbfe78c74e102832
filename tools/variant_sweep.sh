#!/bin/bash
# class_sweep.py for every kernel-variant library under build/variants (development aid)
for lib in cudasw4_b200/libsw4b200.so build/variants/*.so; do
  echo "== $lib"
  SW4B200_LIB=$PWD/$lib timeout 300 python tools/class_sweep.py ${QI:-9} "$@" 2>&1 | grep GCUPS
done
