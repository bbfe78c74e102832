"""Quick GCUPS probe (development aid): peak config (pseudo DB n x L) and its distinct-subject variant."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudasw4_b200 as sw
from cudasw4_b200 import dbformat, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 256
mode = sys.argv[3] if len(sys.argv) > 3 else "both"
queries = synth.load_queries()
def run(tag, setup):
    with sw.CudaSW4(deviceIds=[0], numTop=10, blosumType=62) as eng:
        t0 = time.time(); setup(eng); eng.prefetchDBToGpus(); t1 = time.time()
        eng.scan(queries[0][1])
        tot_cells = tot_s = tot_k = 0.0
        for h, q in queries:
            r = eng.scan(q)
            tot_cells += r.stats.cells; tot_s += r.stats.seconds; tot_k += r.stats.kernelSeconds
            print(f"  {tag} q={len(q):5d}  {r.stats.gcups:8.1f} GCUPS  kernel-only {r.stats.cells/1e9/r.stats.kernelSeconds:8.1f}  launches {r.stats.kernelLaunches}", flush=True)
        print(f"{tag}: upload {t1-t0:.1f}s  total {tot_cells/1e9/tot_s:.1f} GCUPS, kernel-only {tot_cells/1e9/tot_k:.1f} GCUPS", flush=True)
if mode in ("both", "pseudo"):
    run(f"pseudo {n}x{L}", lambda e: e.setPseudoDatabase(n, L))
if mode in ("both", "distinct"):
    db = synth.config_c2(n=n, length=L, distinct=True)
    run(f"distinct {n}x{L}", lambda e: e.setDatabase(db))
