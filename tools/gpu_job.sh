#!/bin/bash
( python tools/sanitize_target.py streaming-only
  timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_target.py streaming-only 2>&1 | tail -5
  timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_target.py streaming-only 2>&1 | tail -5
  python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "benchmark_mode or streaming" 2>&1 | tail -2
) > gpurun_out/r2_probe8.log 2>&1
cat gpurun_out/r2_probe8.log
