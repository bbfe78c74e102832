"""Small mixed workload for compute-sanitizer (memcheck / racecheck): every kernel family runs at least once."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cudasw4_b200 as sw
from cudasw4_b200 import synth, dbformat
rng = np.random.default_rng(5)
q = synth.random_residues(rng, 5200); q[::4] = 17
seqs = [synth.random_residues(rng, int(n)) for n in rng.integers(20, 1400, 260)]
seqs += [synth.random_residues(rng, int(n)) for n in rng.integers(1500, 7000, 24)]
seqs += [q.copy(), synth.mutate(rng, q, 0.05), q[:3000].copy()]
db = dbformat.from_sequences(seqs)
with sw.CudaSW4(deviceIds=[0], numTop=10, blosumType=62) as eng:
    eng.setDatabase(db)
    for ql in (40, 333, 5200):
        r = eng.scan(dbformat.decode(q[:ql]))
        print(ql, r.scores[:3], r.stats.numOverflows, flush=True)
