"""Ad-hoc GPU parity probe (development aid): random DBs per length class vs the oracle, prints mismatch patterns."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudasw4_b200 as sw
from cudasw4_b200 import dbformat, synth
from tests import oracle_lib
orc = oracle_lib.load()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
def run(db, q, blosum, gop, gex, tag):
    with sw.CudaSW4(deviceIds=[0], numTop=5, blosumType=blosum) as eng:
        eng.setGapScores(gop, gex); eng.setDatabase(db)
        outs = []
        for rep in range(3):
            eng.scan(dbformat.decode(q)); s, i = eng.lastScanAllScores(); g = np.empty(len(s), np.int32); g[i] = s; outs.append(g)
        ref = orc.scan(blosum, q, db, gop, gex)
        bad = np.nonzero(outs[0] != ref)[0]
        stable = all((o == outs[0]).all() for o in outs)
        print(f"{tag}: n={db.num_sequences} q={len(q)} blosum{blosum} {gop}/{gex} mismatches={len(bad)} stable={stable}", flush=True)
        if len(bad):
            print("   ids", bad[:12], "len", db.lengths[bad[:12]], "got", outs[0][bad[:12]], "ref", ref[bad[:12]], flush=True)
golden = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden")
tiny = dbformat.read_db(os.path.join(golden, "tinydb/db"))
qs = [dbformat.encode(s) for _, s in synth.load_queries()]
run(tiny, qs[1], 45, -13, -2, "tiny-q1-b45")
run(tiny, qs[1], 62, -11, -1, "tiny-q1-b62")
for (lo, hi) in ((1, 32), (33, 64), (65, 128), (129, 192), (193, 256), (257, 384), (385, 512), (513, 768), (769, 1024)):
    seqs = [synth.random_residues(rng, int(x)) for x in rng.integers(lo, hi + 1, 300)]
    db = dbformat.from_sequences(seqs)
    for ql in (5, 40, 144, 189, 333):
        for blosum, gop, gex in ((62, -11, -1), (45, -13, -2)):
            run(db, synth.random_residues(rng, ql), blosum, gop, gex, f"L{lo}-{hi}")
