"""Small mixed workload for compute-sanitizer (memcheck / racecheck): every kernel family runs at least once - narrow
two-row, multi-segment two-row (classes above 512 columns; the long class too with SW4_LONG_ARRAY_ITEMS_PER_GROUP=0),
CTA-wide s16 / s32 arrays, one-warp s32, both top-k paths - plus query batching, a two-shard handle and streaming."""
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
    for ql in (() if "streaming-only" in sys.argv else (40, 333, 5200)):
        r = eng.scan(dbformat.decode(q[:ql]))
        print(ql, r.scores[:3], r.stats.numOverflows, flush=True)
    if "streaming-only" not in sys.argv:
        many, tot = eng.scanMany([dbformat.decode(q[:n]) for n in (64, 200, 700, 90)])
        print("many", [m.scores[0] for m in many], flush=True)
        eng.setNumTop(5000)
        print("large k", len(eng.scan(dbformat.decode(q[:100])).scores), flush=True)
seqs += [synth.random_residues(rng, int(n)) for n in rng.integers(20, 600, 6000)]
db = dbformat.from_sequences(seqs)
for mem in (1 << 20, 3 << 19, 2 << 20, 3 << 20, 4 << 20):  # the smallest budget whose slots hold the longest block: streaming
    try:
        with sw.CudaSW4(deviceIds=[0, 0], numTop=10, blosumType=45, gop=-13, gex=-2, memoryConfig=sw.MemoryConfig(maxGpuMem=mem)) as eng:
            eng.setDatabase(db)
            many, tot = eng.scanMany([dbformat.decode(q[:n]) for n in (150, 333)])
            info = eng.dbInfo()
            print("two shards, max_gpu_mem", mem, "streaming", info.streaming, "batches", info.num_batches, [m.scores[0] for m in many], flush=True)
            break
    except sw.SW4Error as e:
        print("max_gpu_mem", mem, "->", str(e)[:90], flush=True)
