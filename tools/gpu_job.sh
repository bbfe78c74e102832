#!/bin/bash
# development aid: the measurement batch of the current iteration (run on the GPU box through gpurun)
( tools/variant_probe.sh
  for lib in cudasw4_b200/libsw4b200.so build/variants/ff.so; do echo "-- class sweep $lib"; SW4B200_LIB=$PWD/$lib python tools/class_sweep.py 9 32 64 96 128 192 256 320 384 448 512 640 768 1024 2048; done
  SW4B200_LIB=$PWD/build/variants/ff.so python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "mixed_lengths or every_length or c1_full or long_subjects" 2>&1 | tail -2
) > gpurun_out/r2_probe6.log 2>&1
cat gpurun_out/r2_probe6.log
