#!/usr/bin/env python3
"""Run the reference's own `align` (oracle/_ref/align_gapfix, built for sm_100a from /root/reference: --dpx and the
default half2 kernels) and ours (build/align) back to back on the same GPU box, same database files, same queries, and
write a table to gpurun_out/compare_reference.md (copy into profiles/). Also diffs the TSV results.
usage: python tools/compare_reference.py [configs...]   configs: c2 c2d c3 c5 c4s (default: c2 c2d c3)
c4s = one eighth of the UniRef50-shaped database (8,125,000 subjects / 2.1 G residues: one rank's share of an 8-GPU run)."""
import os, re, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cudasw4_b200 import dbformat, synth

REF = os.path.join(ROOT, "oracle/_ref/align_gapfix")
OURS = os.path.join(ROOT, "build/align")
WORK = os.environ.get("SW4_WORK", "/tmp/sw4_compare")
os.makedirs(WORK, exist_ok=True)
configs = sys.argv[1:] or ["c2", "c2d", "c3"]
qfile = os.path.join(WORK, "allqueries.fasta")
dbformat.write_fasta(qfile, synth.load_queries())

def run(binary, args, tag):
    t0 = time.time()
    r = subprocess.run([binary] + args, cwd=WORK, capture_output=True, text=True)
    if r.returncode != 0:
        return None, None, f"rc={r.returncode} {r.stderr[-300:]}"
    m = re.findall(r"Total time: ([0-9.e+-]+) s, ([0-9.e+-]+) GCUPS", r.stdout)
    per = re.findall(r"Scan time: ([0-9.e+-]+) s, ([0-9.e+-]+) GCUPS", r.stdout)
    return (float(m[-1][1]) if m else None), [float(p[1]) for p in per], f"wall {time.time()-t0:.1f}s"

rows = []
for cfg in configs:
    if cfg == "c2":
        dbargs, name = ["--pseudodb", "1000000", "256"], "C2 PseudoDB 1M x 256 (identical subjects)"
    else:
        prefix = os.path.join(WORK, cfg)
        if not os.path.exists(prefix + "0chars"):
            t0 = time.time()
            if cfg == "c2d": db = synth.config_c2(distinct=True)
            elif cfg == "c3": db = synth.config_c3()
            elif cfg == "c5": db, _ = synth.config_c5()
            elif cfg == "c4s":
                rng = np.random.default_rng(4)
                L = synth.config_c4_lengths(seed=4, n=8_125_000, total=17.0e9 / 8)
                qs = [dbformat.encode(q) for _, q in synth.load_queries()]
                L[: len(qs)] = [len(q) for q in qs]   # equal-length slots for the planted copies of the queries
                db = synth._db_from_sorted_lengths(rng, L, [q.copy() for q in qs], pool=1 << 26)
            dbformat.write_db(prefix, db)
            print(f"generated {cfg}: {db.num_sequences} seqs, {db.num_residues} residues in {time.time()-t0:.1f}s", flush=True)
        dbargs = ["--db", prefix]
        name = {"c2d": "C2' 1M x 256 distinct random subjects", "c3": "C3 Swiss-Prot-shaped (570k seqs, 205M aa)",
                "c5": "C5 long sequences", "c4s": "C4/8 UniRef50-shaped shard (8.1M seqs, 2.1G aa)"}[cfg]
    common = ["--query", qfile, "--verbose", "--uploadFull", "--prefetchDBFile", "--mat", "blosum62", "--tsv"] + dbargs
    res = {}
    for tag, binary, extra in (("ref --dpx", REF, ["--dpx"]), ("ref half2", REF, []), ("ours", OURS, ["--dpx"])):
        top = ["--top", "10", "--of", f"{cfg}_{tag.replace(' ', '_').replace('-', '')}.tsv"]
        total, per, note = run(binary, common + extra + top, tag)
        res[tag] = (total, per, note)
        print(cfg, tag, total, note, flush=True)
    same = {}
    ours_tsv = open(os.path.join(WORK, f"{cfg}_ours.tsv")).read() if res["ours"][0] else ""

    def score_view(text):
        """(query number, result number, score) per line + the ids of the entries strictly above each query's last score:
        what must agree when only the order inside a tie class may differ (above 1e6 subjects the reference's order of
        equal scores is an artefact of its chunked sort)."""
        rows_, per_q = [], {}
        for line in text.splitlines()[1:]:
            c = line.split("\t")
            if len(c) >= 8:
                per_q.setdefault(c[0], []).append((int(c[4]), int(c[7])))
        for qn, lst in per_q.items():
            last = lst[-1][0]
            rows_.append((qn, [sc for sc, _ in lst], sorted(i for sc, i in lst if sc > last)))
        return rows_

    for tag in ("ref --dpx", "ref half2"):
        fn = os.path.join(WORK, f"{cfg}_{tag.replace(' ', '_').replace('-', '')}.tsv")
        ref_tsv = open(fn).read() if os.path.exists(fn) else None
        if ref_tsv is None:
            same[tag] = False
        elif ref_tsv == ours_tsv:
            same[tag] = True
        else:
            same[tag] = "scores + ids above the last score identical (tie order differs)" if score_view(ref_tsv) == score_view(ours_tsv) else False
    rows.append((name, res, same))

with open(os.path.join(ROOT, "gpurun_out", "compare_reference.md"), "w") as f:
    f.write("# ours vs the reference's own align on the same B200 box (total GCUPS over the 20 queries, top-10, TSV)\n\n")
    f.write("| config | reference --dpx | reference half2 | ours | ours / best ref | TSV identical (dpx / half2) |\n|---|---|---|---|---|---|\n")
    for name, res, same in rows:
        a, b, c = res["ref --dpx"][0], res["ref half2"][0], res["ours"][0]
        best = max(x for x in (a, b) if x) if (a or b) else None
        f.write(f"| {name} | {a} | {b} | {c} | {c/best:.2f}x | {same['ref --dpx']} / {same['ref half2']} |\n" if best and c else f"| {name} | {a} | {b} | {c} | n/a | {same} |\n")
    f.write("\nper-query GCUPS (query order of allqueries.fasta):\n\n")
    for name, res, _ in rows:
        for tag in res:
            if res[tag][1]:
                f.write(f"* {name} — {tag}: " + " ".join(f"{x:.0f}" for x in res[tag][1]) + "\n")
print(open(os.path.join(ROOT, "gpurun_out", "compare_reference.md")).read())
