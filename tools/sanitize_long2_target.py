"""Tiny workload for compute-sanitizer on the two-row array kernel: 60 subjects of 1100..5000 aa, queries of 333 and 700 aa
(4- and 8-warp arrays), one with a planted hit so that block borders carry real scores."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cudasw4_b200 as sw
from cudasw4_b200 import synth, dbformat
rng = np.random.default_rng(8)
q = synth.random_residues(rng, 700)
seqs = [synth.random_residues(rng, int(n)) for n in rng.integers(1100, 5000, 58)]
seqs += [np.concatenate([synth.random_residues(rng, 900), q, synth.random_residues(rng, 700)]), np.concatenate([q[:333], synth.random_residues(rng, 1500)])]
db = dbformat.from_sequences(seqs)
with sw.CudaSW4(deviceIds=[0], numTop=5, blosumType=62) as eng:
    eng.setDatabase(db)
    for ql in (333, 700):
        r = eng.scan(dbformat.decode(q[:ql]))
        print(ql, r.scores[:3], flush=True)
