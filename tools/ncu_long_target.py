"""Small fixed workload for ncu captures of the array kernels: C5-shaped long subjects, one long query (with a planted
copy, so the exact 32-bit re-scoring runs too)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudasw4_b200 as sw
from cudasw4_b200 import synth, dbformat
qi = int(sys.argv[1]) if len(sys.argv) > 1 else 2
db, queries = synth.config_c5(n_subjects=400)
with sw.CudaSW4(deviceIds=[0], numTop=10, blosumType=45) as eng:
    eng.setGapScores(-13, -2); eng.setDatabase(db); eng.prefetchDBToGpus()
    for _ in range(2):
        r = eng.scan(dbformat.decode(queries[qi]))
    print(len(queries[qi]), r.stats.gcups, r.stats.numOverflows)
