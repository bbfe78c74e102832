#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ from the REFERENCE's own code, run in the authoring
container (needs /root/reference and `make -C oracle ref`). Everything the GPU box needs is written here; nothing at
test/bench time reads /root/reference.

  allqueries.json        the 20 benchmark queries (reference allqueries.fasta; public UniProt entries) as a fixture
  ref_cpu_gotoh.json     scores from the reference's private scalar CPU Gotoh (src/cudasw4.cuh:2331-2392, BLOSUM62)
  ref_misc.json          letter map for all 256 byte values, partition boundaries, PseudoDB subjects
  tinydb/ , tiesdb/      output files of the reference `makedb` binary for two small FASTA inputs (+ the inputs)
  survey_kat.json        the known-answer tables of SURVEY.md §8c (two independent implementations, all 4 matrices)
"""
import ctypes, json, os, subprocess, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cudasw4_b200 import dbformat, synth  # noqa: E402

REF = "/root/reference"
G = os.path.join(ROOT, "tests/golden")
h = ctypes.CDLL(os.path.join(ROOT, "oracle/_ref/libref_harness.so"))
h.ref_cpu_gotoh_blosum62.restype = ctypes.c_int
h.ref_pseudodb.restype = ctypes.c_long


def ref_score(q, s, gop, gex):
    q = np.ascontiguousarray(q, dtype=np.uint8); s = np.ascontiguousarray(s, dtype=np.uint8)
    return int(h.ref_cpu_gotoh_blosum62(q.ctypes.data_as(ctypes.c_char_p), s.ctypes.data_as(ctypes.c_char_p),
                                        len(q), len(s), gop, gex))


def main():
    os.makedirs(G, exist_ok=True)
    recs = dbformat.read_fasta(os.path.join(REF, "allqueries.fasta"))
    assert len(recs) == 20 and sum(len(s) for _, s in recs) == 41752
    with open(os.path.join(ROOT, "cudasw4_b200", "data", "allqueries.json"), "w") as f:
        json.dump({"source": "reference allqueries.fasta (20 UniProt proteins used by run*benchmark.sh)",
                   "records": [{"header": hd, "sequence": s} for hd, s in recs]}, f, indent=0)

    # --- reference makedb on two small inputs -------------------------------------------------------------
    for name, records in (
        ("tinydb", None),  # first 40 lines of allqueries.fasta, as in SURVEY.md §3.1 / §10
        ("tiesdb", "ties"),
    ):
        d = os.path.join(G, name); os.makedirs(d, exist_ok=True)
        fa = os.path.join(d, "input.fasta")
        if records is None:
            with open(os.path.join(REF, "allqueries.fasta")) as src, open(fa, "w") as dst:
                for i, line in enumerate(src):
                    if i >= 40: break
                    dst.write(line)
        else:  # 300 sequences, few distinct lengths => exercises the (unstable) sort order + odd letters + padding
            rng = np.random.default_rng(11)
            rs = []
            for i in range(300):
                L = int(rng.choice([1, 3, 4, 5, 17, 48, 49, 64, 65, 100, 256, 257, 300]))
                s = list(dbformat.decode(synth.random_residues(rng, L)))
                if i % 7 == 0: s[int(rng.integers(0, L))] = "X"
                if i % 11 == 0: s[int(rng.integers(0, L))] = "b"
                rs.append((f"t{i} some description {i*i}", "".join(s)))
            dbformat.write_fasta(fa, rs, width=70)
        for fn in os.listdir(d):
            if fn.startswith("db"): os.remove(os.path.join(d, fn))
        subprocess.run([os.path.join(ROOT, "oracle/_ref/makedb"), fa, os.path.join(d, "db")], check=True,
                       stdout=subprocess.DEVNULL)

    # --- reference CPU Gotoh (BLOSUM62) -------------------------------------------------------------------
    tiny = dbformat.read_db(os.path.join(G, "tinydb/db"))
    cases = []
    qs = [dbformat.encode(s) for _, s in recs]
    for gop, gex in ((-11, -1), (-5, -3), (-13, -2), (-1, -1)):
        for qi in range(3):
            cases.append({"kind": "tinydb", "query": qi, "gop": gop, "gex": gex,
                          "scores": [ref_score(qs[qi], tiny.sequence(j), gop, gex) for j in range(tiny.num_sequences)]})
    rng = np.random.default_rng(5)
    rnd = []
    for k in range(200):
        lq, ls = int(rng.integers(0, 300)), int(rng.integers(0, 300))
        if k < 6: lq, ls = [(0, 5), (5, 0), (1, 1), (1, 300), (300, 1), (0, 0)][k]
        q = rng.integers(0, 21, lq).astype(np.uint8); s = rng.integers(0, 21, ls).astype(np.uint8)
        if k % 5 == 0 and lq > 20:  # plant similarity
            s = np.concatenate([s[: ls // 2], q[5:lq - 5], s[ls // 2:]]).astype(np.uint8)
        gop, gex = [(-11, -1), (-10, -1), (-13, -2), (-3, -3)][k % 4]
        rnd.append({"q": q.tolist(), "s": s.tolist(), "gop": gop, "gex": gex, "score": ref_score(q, s, gop, gex)})
    # a high scoring self alignment (above the half2 limit 2048, below the s16 limit)
    big = qs[19]
    rnd.append({"q": big.tolist(), "s": big.tolist(), "gop": -11, "gex": -1, "score": ref_score(big, big, -11, -1)})
    with open(os.path.join(G, "ref_cpu_gotoh.json"), "w") as f:
        json.dump({"source": "reference src/cudasw4.cuh:2331-2392 via oracle/ref_harness.cu (BLOSUM62_20)",
                   "tinydb_cases": cases, "random_cases": rnd}, f)

    # --- misc -------------------------------------------------------------------------------------------------
    allbytes = bytes(range(256))
    out = ctypes.create_string_buffer(256)
    h.ref_convert_aa(allbytes, out, 256)
    b = (ctypes.c_int * 64)(); nb = h.ref_length_partition_boundaries(b, 64)
    pseudo = {}
    for L, seed in ((128, 42), (256, 42), (512, 42), (1024, 42), (77, 7)):
        Lp = (L + 3) // 4 * 4
        buf = np.zeros(3 * Lp, np.uint8); lens = np.zeros(3, np.int32); offs = np.zeros(4, np.uint64)
        n = h.ref_pseudodb(ctypes.c_long(3), L, seed, buf.ctypes.data_as(ctypes.c_char_p), ctypes.c_long(buf.size),
                           lens.ctypes.data_as(ctypes.c_void_p), offs.ctypes.data_as(ctypes.c_void_p))
        assert n == 3 * Lp
        pseudo[f"{L}_{seed}"] = {"chars": buf.tolist(), "lengths": lens.tolist(), "offsets": offs.tolist()}
    with open(os.path.join(G, "ref_misc.json"), "w") as f:
        json.dump({"convert_256": list(out.raw), "boundaries": list(b)[:nb], "pseudodb": pseudo}, f)

    # --- SURVEY.md §8c known answers (generated during the survey by two independent implementations) -------------
    kat = {
        "tinydb": [
            {"blosum": 62, "gop": -11, "gex": -1, "query": 0, "scores": [24, 719, 26, 28, 26, 33, 29]},
            {"blosum": 62, "gop": -11, "gex": -1, "query": 1, "scores": [23, 26, 977, 35, 28, 30, 31]},
            {"blosum": 62, "gop": -11, "gex": -1, "query": 2, "scores": [19, 28, 35, 1135, 30, 33, 32]},
            {"blosum": 45, "gop": -13, "gex": -2, "query": 0, "scores": [33, 849, 32, 34, 30]},
            {"blosum": 45, "gop": -13, "gex": -2, "query": 1, "scores": [29, 32, 1182, 46, 51]},
            {"blosum": 50, "gop": -13, "gex": -2, "query": 0, "scores": [32, 910, 37, 36, 33]},
            {"blosum": 50, "gop": -13, "gex": -2, "query": 1, "scores": [30, 37, 1258, 45, 47]},
            {"blosum": 80, "gop": -10, "gex": -1, "query": 0, "scores": [24, 780, 26, 24, 26]},
            {"blosum": 80, "gop": -10, "gex": -1, "query": 1, "scores": [24, 26, 1065, 30, 24]},
            {"blosum": 62, "gop": -5, "gex": -3, "query": 0, "scores": [25, 719, 45, 38, 40]},
            {"blosum": 62, "gop": -5, "gex": -3, "query": 1, "scores": [23, 45, 977, 54, 49]},
        ],
        "pseudodb_blosum62_gop-11_gex-1": {
            "128": [24, 25, 38, 29, 34, 26, 33, 39, 28, 37, 34, 37, 41, 31, 37, 31, 34, 33, 40, 26],
            "256": [25, 26, 38, 29, 34, 35, 33, 39, 31, 37, 37, 38, 51, 40, 41, 38, 34, 33, 40, 36],
            "512": [27, 26, 38, 31, 38, 38, 33, 39, 33, 37, 37, 43, 51, 41, 50, 38, 44, 44, 40, 36],
            "1024": [35, 30, 38, 45, 38, 41, 37, 39, 33, 37, 37, 44, 51, 41, 50, 39, 47, 44, 40, 36],
        },
    }
    with open(os.path.join(G, "survey_kat.json"), "w") as f:
        json.dump(kat, f)
    print("golden fixtures written to", G)


if __name__ == "__main__":
    main()
