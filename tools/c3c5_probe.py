"""C3 (Swiss-Prot-shaped) and C5 (long sequences) throughput probe (development aid): total + per-query GCUPS.
usage: python tools/c3c5_probe.py [c3] [c5] [queries=0,5,9,14,19]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudasw4_b200 as sw
from cudasw4_b200 import synth
which = [a for a in sys.argv[1:] if a in ("c3", "c5")] or ["c3", "c5"]
qsel = [a for a in sys.argv[1:] if a.startswith("queries=")]
qidx = [int(x) for x in qsel[0].split("=")[1].split(",")] if qsel else list(range(20))
queries = synth.load_queries()
for cfg in which:
    db = synth.config_c3() if cfg == "c3" else synth.config_c5()[0]
    with sw.CudaSW4(deviceIds=[0], numTop=10, blosumType=62) as eng:
        eng.setDatabase(db)
        eng.prefetchDBToGpus()
        eng.scan(queries[0][1])
        cells = secs = 0.0
        per = []
        for qi in qidx:
            best = None
            for _ in range(2):
                r = eng.scan(queries[qi][1])
                if best is None or r.stats.seconds < best.stats.seconds:
                    best = r
            cells += best.stats.cells; secs += best.stats.seconds
            per.append(f"{best.stats.gcups:.0f}")
        many, tot = eng.scanMany([queries[qi][1] for qi in qidx])
        print(f"{cfg}: total {cells/1e9/secs:.1f} GCUPS (one scan at a time), scan_many {tot.gcups:.1f} GCUPS | per query: {' '.join(per)}", flush=True)
