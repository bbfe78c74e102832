"""Scale / long-sequence probes (development aid).
  python tools/scale_probe.py c5            long queries x long subjects vs oracle (subset) + GCUPS
  python tools/scale_probe.py c4 NSEQ       UniRef50-shaped shard with NSEQ sequences: upload time + GCUPS
  python tools/scale_probe.py c4full [NSEQ] the whole C4 shape (65 M sequences / 17 G residues) on the visible GPUs, planted
                                            queries + a random sample of subjects checked against the oracle"""
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
elif mode == "c4full":
    from tests import oracle_lib
    orc = oracle_lib.load()
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 65_000_000
    t0 = time.time()
    rng = np.random.default_rng(4)
    L = np.sort(synth.lognormal_lengths(rng, n, 5.247, 0.80, 11, 45000, total=n * 261.5))
    queries = synth.load_queries()
    planted = [dbformat.encode(queries[qi][1]) for qi in (0, 5, 9, 14, 19)]
    for j, pseq in enumerate(planted):  # make an equal-length slot for every planted query
        L[np.searchsorted(L, len(pseq))] = len(pseq)
    L = np.sort(L)
    db = synth._db_from_sorted_lengths(rng, L, planted, pool=1 << 26)
    print(f"generated {n} seqs, {db.num_residues} residues, max len {int(db.lengths.max())} in {time.time()-t0:.1f}s", flush=True)
    sample = np.sort(np.concatenate([rng.integers(0, n, 3000), np.arange(n - 20, n), np.arange(0, 20)]))
    with sw.CudaSW4(numTop=10, blosumType=62, verbose=True) as eng:
        t0 = time.time(); eng.setDatabase(db); t1 = time.time(); eng.prefetchDBToGpus(); t2 = time.time()
        print(f"setDatabase {t1-t0:.1f}s upload+layout+warm-up {t2-t1:.1f}s", flush=True)
        tot_c = tot_s = 0
        for qi in (0, 5, 9, 14, 19):
            q = queries[qi][1]
            r = eng.scan(q); tot_c += r.stats.cells; tot_s += r.stats.seconds
            qc = dbformat.encode(q)
            selfscore = orc.score(62, qc, qc, -11, -1)
            sc, ids = eng.lastScanAllScores()
            got = np.empty(n, np.int32); got[ids] = sc
            ref = orc.scan(62, qc, db, -11, -1, subset=sample)
            ok = bool((got[sample] == ref).all())
            s, i = orc.topk(got, 10)
            print(f"q{qi} len {len(q)}: {r.stats.gcups:.1f} GCUPS ovf {r.stats.numOverflows} top1 {r.scores[0]} (self {selfscore}) id {r.referenceIds[0]} "
                  f"len {eng.getReferenceLength(r.referenceIds[0])} | sample of {len(sample)} vs oracle equal {ok} | top-k vs host selection equal "
                  f"{r.scores == s.tolist() and r.referenceIds == i.tolist()}", flush=True)
        print(f"total {tot_c/1e9/tot_s:.1f} GCUPS", flush=True)
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
