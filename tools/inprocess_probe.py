"""In-process multi-GPU probe (development aid): ONE handle over all visible GPUs (one host thread per GPU, host merge of
the per-GPU top-k lists), peak-benchmark database of 1 M x 256 subjects per GPU, the 20 queries through sw4_scan_many."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cudasw4_b200 as sw
from cudasw4_b200 import synth
n = torch.cuda.device_count()
qs = [q for _, q in synth.load_queries()]
with sw.CudaSW4(deviceIds=list(range(n)), numTop=10, blosumType=62) as eng:
    eng.setPseudoDatabase(1_000_000 * n, 256)
    eng.prefetchDBToGpus()
    eng.scanMany(qs)
    best = max(eng.scanMany(qs)[1].gcups for _ in range(3))
    res = eng.scan(qs[0])
    print(f"{n} GPUs in one process: scan_many {best:.1f} GCUPS, top ids {res.referenceIds[:5]}, scores {res.scores[:3]}", flush=True)
