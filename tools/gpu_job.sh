#!/bin/bash
( for d in 2 4 6; do echo "-- SW4_PIPELINE=$d"; SW4_PIPELINE=$d python tools/c3c5_probe.py c5 c3; done
  echo "-- C2 scan_many by depth"; for d in 2 3 4; do SW4_PIPELINE=$d python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import cudasw4_b200 as sw
from cudasw4_b200 import synth
qs = [q for _, q in synth.load_queries()]
with sw.CudaSW4(deviceIds=[0], numTop=10, blosumType=62) as eng:
    eng.setPseudoDatabase(1000000, 256)
    eng.prefetchDBToGpus()
    eng.scanMany(qs)
    best = max(eng.scanMany(qs)[1].gcups for _ in range(3))
    print("depth", os.environ["SW4_PIPELINE"], "C2 scan_many", round(best, 1), flush=True)
PY
done ) > gpurun_out/r2_probe12.log 2>&1
cat gpurun_out/r2_probe12.log
