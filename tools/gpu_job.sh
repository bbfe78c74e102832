#!/bin/bash
( timeout 500 compute-sanitizer --tool racecheck python tools/sanitize_long2_target.py 2>&1 | tail -4
  timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_long2_target.py 2>&1 | tail -2
  timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "long or c5 or saturating or beyond_16 or border or mixed_lengths or c3_shaped" 2>&1 | tail -3
  python tools/c3c5_probe.py c5
) > gpurun_out/r2_probe11.log 2>&1
cat gpurun_out/r2_probe11.log
