"""C3 (Swiss-Prot-shaped) probe: per-query GCUPS and, under ncu, per-class kernel durations."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cudasw4_b200 as sw
from cudasw4_b200 import synth, dbformat
qsel = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 9, 19]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
db = synth.config_c3()
L = db.lengths.astype(np.int64)
print("seqs", db.num_sequences, "residues", L.sum(), ">1024:", (L > 1024).sum(), "res frac", L[L > 1024].sum() / L.sum(), flush=True)
queries = synth.load_queries()
with sw.CudaSW4(deviceIds=[0], numTop=10, blosumType=62, verbose=True) as eng:
    eng.setDatabase(db); eng.prefetchDBToGpus()
    for qi in qsel:
        for _ in range(reps):
            r = eng.scan(queries[qi][1])
        print(f"q{qi} len {len(queries[qi][1])}: {r.stats.gcups:.1f} GCUPS kernel-only {r.stats.cells/1e9/r.stats.kernelSeconds:.1f} launches {r.stats.kernelLaunches} ovf {r.stats.numOverflows}", flush=True)
    if os.environ.get("C3_CHECK"):
        from tests import oracle_lib
        orc = oracle_lib.load()
        for qi in qsel:
            q = dbformat.encode(queries[qi][1])
            r = eng.scan(queries[qi][1])
            ref = [orc.score(62, q, db.sequence(i), -11, -1) for i in r.referenceIds]
            print("check q", qi, "scores", r.scores, "oracle", ref, "ovf", r.stats.numOverflows, "lens", [int(db.lengths[i]) for i in r.referenceIds], flush=True)
