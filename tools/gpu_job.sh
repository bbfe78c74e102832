#!/bin/bash
( python tools/perf_probe.py 1000000 256 pseudo | grep -E "q=  144|q=  189|q=  375|q= 1000|q= 5478|total"
  timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=5 2>&1 | tail -3
) > gpurun_out/r2_probe13.log 2>&1
cat gpurun_out/r2_probe13.log
