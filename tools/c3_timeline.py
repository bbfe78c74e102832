"""C3 per-class timeline (development aid): SW4_DEBUG_PARTITION=1 python tools/c3_timeline.py QIDX"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudasw4_b200 as sw
from cudasw4_b200 import synth
qi = int(sys.argv[1])
db = synth.config_c3()
q = synth.load_queries()[qi][1]
with sw.CudaSW4(deviceIds=[0], numTop=10, blosumType=62) as eng:
    eng.setDatabase(db); eng.prefetchDBToGpus()
    eng.scan(q); eng.scan(q)
    sys.stderr.write("==== timed scan\n"); sys.stderr.flush()
    r = eng.scan(q)
    sys.stderr.write(f"q{qi} len {len(q)}: {r.stats.gcups:.1f} GCUPS, {r.stats.seconds*1e3:.3f} ms, kernels {r.stats.kernelSeconds*1e3:.3f} ms\n")
