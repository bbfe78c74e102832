"""Per-length-class throughput (development aid): one pseudo database per class capacity (n x L = 128 M residues, every
subject fills its class exactly), a few scans of one query -> GCUPS of that kernel instantiation alone.
usage: python tools/class_sweep.py [query index] [lengths...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudasw4_b200 as sw
from cudasw4_b200 import synth
qi = int(sys.argv[1]) if len(sys.argv) > 1 else 9
caps = [int(x) for x in sys.argv[2:]] or ([8 * r for r in range(4, 32, 2)] + [16 * r for r in range(16, 34, 2)] + [32 * r for r in range(18, 34, 2)] + [2048])
q = synth.load_queries()[qi][1]
for L in caps:
    n = max(2000, int(os.environ.get("SWEEP_RESIDUES", 128_000_000)) // L)
    with sw.CudaSW4(deviceIds=[0], numTop=10, blosumType=62) as eng:
        eng.setPseudoDatabase(n, L)
        eng.prefetchDBToGpus()
        best = 0
        for _ in range(3):
            r = eng.scan(q)
            best = max(best, r.stats.cells / 1e9 / r.stats.kernelSeconds)
        print(f"L {L:5d} n {n:8d} q {len(q)}: {best:8.1f} GCUPS (kernels)", flush=True)
