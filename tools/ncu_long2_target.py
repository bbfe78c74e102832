"""Fixed workload for ncu captures of the two-row array kernel: 800 C5-shaped long subjects, one 2504-aa query."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudasw4_b200 as sw
from cudasw4_b200 import synth
db, _ = synth.config_c5(n_subjects=800)
q = synth.load_queries()[12][1]
with sw.CudaSW4(deviceIds=[0], numTop=10, blosumType=62) as eng:
    eng.setDatabase(db); eng.prefetchDBToGpus()
    for _ in range(2):
        r = eng.scan(q)
    print(len(q), r.stats.gcups, r.stats.numOverflows)
