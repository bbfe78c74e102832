#!/bin/bash
# development aid: the measurement batch of the current iteration (run on the GPU box through gpurun)
python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=8 2>&1 | tail -25 > gpurun_out/r2_test5.log; tail -4 gpurun_out/r2_test5.log
( python tools/perf_probe.py 1000000 256 pseudo | grep -E "q=  144|q=  375|q= 1000|q= 2504|q= 5478|total"
  echo "-- 8x32 for every query"; SW4_CLASS256_CROSSOVER=100000 python tools/perf_probe.py 1000000 256 pseudo | grep -E "q= 1000|q= 2504|q= 5478|total"
  echo "-- 16x16 for every query"; SW4_CLASS256_CROSSOVER=0 python tools/perf_probe.py 1000000 256 pseudo | grep -E "q=  144|q=  375|q= 1000|q= 2504|q= 5478|total"
  echo "-- class sweep"; python tools/class_sweep.py 9
  echo "-- class sweep, one-row kernels above 512"; SW4_NO_TWO_ROW_MULTI=1 python tools/class_sweep.py 9 576 640 704 768 832 896 960 1024 2048
  python tools/c3c5_probe.py ) > gpurun_out/r2_probe5.log 2>&1
cat gpurun_out/r2_probe5.log
