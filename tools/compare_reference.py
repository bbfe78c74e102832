#!/usr/bin/env python3
"""Run the reference's own `align` (oracle/_ref/align_gapfix, built for sm_100a from /root/reference: --dpx and the
default half2 kernels) and ours (build/align) back to back on the same GPU box, same database files, same queries, and
write a table to gpurun_out/compare_reference.md (copy into profiles/). Also diffs the TSV results.
usage: python tools/compare_reference.py [configs...]   configs: c2 c2d c3 c5 (default: c2 c2d c3)"""
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
            dbformat.write_db(prefix, db)
            print(f"generated {cfg}: {db.num_sequences} seqs, {db.num_residues} residues in {time.time()-t0:.1f}s", flush=True)
        dbargs = ["--db", prefix]
        name = {"c2d": "C2' 1M x 256 distinct random subjects", "c3": "C3 Swiss-Prot-shaped (570k seqs, 205M aa)",
                "c5": "C5 long sequences"}[cfg]
    common = ["--query", qfile, "--verbose", "--uploadFull", "--prefetchDBFile", "--mat", "blosum62", "--tsv"] + dbargs
    res = {}
    for tag, binary, extra in (("ref --dpx", REF, ["--dpx"]), ("ref half2", REF, []), ("ours", OURS, ["--dpx"])):
        top = ["--top", "10", "--of", f"{cfg}_{tag.replace(' ', '_').replace('-', '')}.tsv"]
        total, per, note = run(binary, common + extra + top, tag)
        res[tag] = (total, per, note)
        print(cfg, tag, total, note, flush=True)
    same = {}
    ours_tsv = open(os.path.join(WORK, f"{cfg}_ours.tsv")).read() if res["ours"][0] else ""
    for tag in ("ref --dpx", "ref half2"):
        fn = os.path.join(WORK, f"{cfg}_{tag.replace(' ', '_').replace('-', '')}.tsv")
        same[tag] = os.path.exists(fn) and open(fn).read() == ours_tsv
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
