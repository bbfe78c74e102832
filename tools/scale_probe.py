"""Scale / long-sequence probes (development aid).
  python tools/scale_probe.py c5            long queries x long subjects vs oracle (subset) + GCUPS
  python tools/scale_probe.py c4 NSEQ       UniRef50-shaped shard with NSEQ sequences: upload time + GCUPS"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cudasw4_b200 as sw
from cudasw4_b200 import synth, dbformat
mode = sys.argv[1]
if mode == "c5":
    from tests import oracle_lib
    orc = oracle_lib.load()
    db, queries = synth.config_c5(n_subjects=400)
    print("C5:", db.num_sequences, "subjects", db.num_residues, "residues, max len", int(db.lengths.max()), flush=True)
    for blosum, gop, gex in ((45, -13, -2), (80, -14, -2)):
        with sw.CudaSW4(deviceIds=[0], numTop=10, blosumType=blosum) as eng:
            eng.setGapScores(gop, gex); eng.setDatabase(db); eng.prefetchDBToGpus()
            for qi, q in enumerate(queries):
                r = eng.scan(dbformat.decode(q))
                line = f"blosum{blosum} {gop}/{gex} q{qi} len {len(q)}: {r.stats.gcups:.1f} GCUPS (kernels {r.stats.cells/1e9/r.stats.kernelSeconds:.1f}) ovf {r.stats.numOverflows} top {r.scores[:3]}"
                if qi in (0, 4, 7):  # oracle check (CPU: ~1e10 cells each)
                    t0 = time.time(); ref = orc.scan(blosum, q, db, gop, gex); s, i = orc.topk(ref, 10)
                    sc, ids = eng.lastScanAllScores(); got = np.empty_like(ref); got[ids] = sc
                    line += f" | oracle {time.time()-t0:.0f}s all-equal {bool((got == ref).all())} topk-equal {r.scores == s.tolist() and r.referenceIds == i.tolist()}"
                print(line, flush=True)
else:
    n = int(sys.argv[2])
    t0 = time.time()
    rng = np.random.default_rng(4)
    L = np.sort(synth.lognormal_lengths(rng, n, 5.247, 0.80, 11, 45000, total=n * 261.5))
    db = synth._db_from_sorted_lengths(rng, L)
    print(f"generated {n} seqs, {db.num_residues} residues in {time.time()-t0:.1f}s", flush=True)
    queries = synth.load_queries()
    with sw.CudaSW4(deviceIds=[0], numTop=10, blosumType=62, verbose=True) as eng:
        t0 = time.time(); eng.setDatabase(db); t1 = time.time(); eng.prefetchDBToGpus(); t2 = time.time()
        print(f"setDatabase {t1-t0:.1f}s upload+layout {t2-t1:.1f}s", flush=True)
        tot_c = tot_s = 0
        for qi in (0, 5, 9, 14, 19):
            r = eng.scan(queries[qi][1]); tot_c += r.stats.cells; tot_s += r.stats.seconds
            print(f"q{qi} len {len(queries[qi][1])}: {r.stats.gcups:.1f} GCUPS ovf {r.stats.numOverflows} top {r.scores[:3]} {r.referenceIds[:3]}", flush=True)
        print(f"total {tot_c/1e9/tot_s:.1f} GCUPS", flush=True)
