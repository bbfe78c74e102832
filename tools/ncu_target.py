"""Small fixed workload for ncu captures: pseudo DB n x L, a few scans of one query."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudasw4_b200 as sw
from cudasw4_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 256
qi = int(sys.argv[3]) if len(sys.argv) > 3 else 9
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
q = synth.load_queries()[qi][1]
with sw.CudaSW4(deviceIds=[0], numTop=10, blosumType=62) as eng:
    eng.setPseudoDatabase(n, L)
    for _ in range(reps):
        r = eng.scan(q)
    print(len(q), r.stats.gcups, r.stats.kernelSeconds)
