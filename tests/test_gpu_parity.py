"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes -> libsw4b200.so), against the CPU
oracle on the same seeded inputs, against the committed golden fixtures, and through size-independent properties at
the benchmark's full size. Bar: bit-exact scores, identical top-k lists with ties broken by ascending DB id."""
import json
import os

import numpy as np
import pytest

import cudasw4_b200 as sw
from cudasw4_b200 import dbformat, synth

pytestmark = pytest.mark.gpu

MATRICES = [(62, -11, -1), (45, -13, -2), (50, -13, -2), (80, -10, -1), (62, -5, -3), (45, -9, -3), (80, -14, -2)]


def _engine(**kw):
    return sw.CudaSW4(deviceIds=[0], **kw)


def _all_scores(eng, query):
    eng.scan(query)
    scores, ids = eng.lastScanAllScores()
    out = np.empty(len(scores), dtype=np.int32)
    out[ids] = scores
    return out


def _mixed_db(seed, n, lo, hi, extra_lengths=()):
    rng = np.random.default_rng(seed)
    L = np.concatenate([rng.integers(lo, hi + 1, n), np.array(extra_lengths, dtype=np.int64)])
    seqs = [synth.random_residues(rng, int(x)) for x in L]
    for s in seqs[::17]:  # sprinkle the 'other' code
        if len(s):
            s[rng.integers(0, len(s))] = 20
    return dbformat.from_sequences(seqs), rng


def test_golden_tinydb_files(oracle, golden_dir):
    with open(os.path.join(golden_dir, "survey_kat.json")) as f:
        kat = json.load(f)
    queries = synth.load_queries()
    tiny = dbformat.read_db(os.path.join(golden_dir, "tinydb", "db"))
    for c in kat["tinydb"]:
        with _engine(numTop=7, blosumType=c["blosum"]) as eng:
            eng.setGapScores(c["gop"], c["gex"])
            eng.setDatabase(os.path.join(golden_dir, "tinydb", "db"))
            got = _all_scores(eng, queries[c["query"]][1])
            assert got[: len(c["scores"])].tolist() == c["scores"], c
            ref = oracle.scan(c["blosum"], dbformat.encode(queries[c["query"]][1]), tiny, c["gop"], c["gex"])
            assert got.tolist() == ref.tolist()
            assert eng.getReferenceLength(3) == 222
            assert eng.getReferenceHeader(1).startswith("sp|") or len(eng.getReferenceHeader(1)) > 0
            assert eng.getReferenceSequence(1) == dbformat.decode(tiny.sequence(1))


def test_golden_reference_cpu_scores(golden_dir):
    """Scores the reference's own CPU Gotoh produced (tests/golden/ref_cpu_gotoh.json) for the tiny database."""
    with open(os.path.join(golden_dir, "ref_cpu_gotoh.json")) as f:
        g = json.load(f)
    queries = synth.load_queries()
    for c in g["tinydb_cases"]:
        with _engine(numTop=7, blosumType=62) as eng:
            eng.setGapScores(c["gop"], c["gex"])
            eng.setDatabase(os.path.join(golden_dir, "tinydb", "db"))
            assert _all_scores(eng, queries[c["query"]][1]).tolist() == c["scores"], c


def test_pseudodb_known_answers_and_tie_rule(golden_dir):
    with open(os.path.join(golden_dir, "survey_kat.json")) as f:
        kat = json.load(f)["pseudodb_blosum62_gop-11_gex-1"]
    queries = synth.load_queries()
    for L in (128, 256, 512, 1024):
        with _engine(numTop=10, blosumType=62) as eng:
            eng.setPseudoDatabase(3001, L)
            for qi in (0, 5, 12, 19):
                res = eng.scan(queries[qi][1])
                assert res.scores == [kat[str(L)][qi]] * 10
                assert res.referenceIds == list(range(10))  # all subjects tie: ascending DB id
                scores, _ = eng.lastScanAllScores()
                assert (scores == kat[str(L)][qi]).all()


@pytest.mark.parametrize("blosum,gop,gex", MATRICES)
def test_mixed_lengths_all_scores_match_oracle(oracle, blosum, gop, gex):
    # every length class of the packed kernel (1..1024), class boundaries, and long subjects on the exact 32-bit path
    edges = [1, 2, 31, 32, 33, 64, 65, 96, 97, 128, 129, 192, 193, 256, 257, 384, 385, 512, 513, 768, 769, 1023, 1024,
             1025, 1500, 2500]
    db, rng = _mixed_db(100 + blosum + gop, 1500, 1, 900, edges)
    queries = [synth.random_residues(rng, n) for n in (1, 3, 37, 144, 300, 517)]
    with _engine(numTop=25, blosumType=blosum) as eng:
        eng.setGapScores(gop, gex)
        eng.setDatabase(db)
        for q in queries:
            letters = dbformat.decode(q)
            res = eng.scan(letters)
            got_scores, got_ids = eng.lastScanAllScores()
            got = np.empty(db.num_sequences, np.int32)
            got[got_ids] = got_scores
            ref = oracle.scan(blosum, q, db, gop, gex)
            bad = np.nonzero(got != ref)[0]
            assert len(bad) == 0, (len(q), bad[:10], got[bad[:10]], ref[bad[:10]], db.lengths[bad[:10]])
            s, i = oracle.topk(ref, 25)
            assert res.scores == s.tolist() and res.referenceIds == i.tolist()


@pytest.mark.parametrize("lo,hi", [(1, 32), (33, 64), (65, 128), (129, 192), (193, 256), (257, 384), (385, 512),
                                   (513, 768), (769, 1024), (1025, 1700)])
def test_every_length_class_alone(oracle, lo, hi):
    """One database per length-class range (so that single classes get the whole GPU and every (G, R) kernel
    instantiation runs with many rounds), several query lengths incl. odd ones and ones shorter than a pipeline."""
    rng = np.random.default_rng(1000 + lo)
    seqs = [synth.random_residues(rng, int(x)) for x in rng.integers(lo, hi + 1, 700)]
    db = dbformat.from_sequences(seqs)
    with _engine(numTop=8, blosumType=62) as eng:
        eng.setDatabase(db)
        for ql, (blosum, gop, gex) in zip((5, 40, 145, 333), ((62, -11, -1), (45, -13, -2), (80, -10, -1), (50, -13, -2))):
            q = synth.random_residues(rng, ql)
            eng.setBlosum(blosum)
            eng.setGapScores(gop, gex)
            res = eng.scan(dbformat.decode(q))
            scores, ids = eng.lastScanAllScores()
            got = np.empty(db.num_sequences, np.int32)
            got[ids] = scores
            ref = oracle.scan(blosum, q, db, gop, gex)
            bad = np.nonzero(got != ref)[0]
            assert len(bad) == 0, (ql, blosum, bad[:8], got[bad[:8]], ref[bad[:8]], db.lengths[bad[:8]])
            s, i = oracle.topk(ref, 8)
            assert res.scores == s.tolist() and res.referenceIds == i.tolist()


def test_planted_homologs_and_ties(oracle):
    recs, queries = synth.config_c1(seed=1, n=1200)
    seqs = [dbformat.encode(s) for _, s in recs]
    db = dbformat.from_sequences(seqs, [h for h, _ in recs])
    with _engine(numTop=10, blosumType=62) as eng:
        eng.setDatabase(db)
        for qi in (0, 3, 7, 19):
            q = dbformat.encode(queries[qi][1])
            res = eng.scan(queries[qi][1])
            ref = oracle.scan(62, q, db, -11, -1)
            s, i = oracle.topk(ref, 10)
            assert res.scores == s.tolist() and res.referenceIds == i.tolist()
            assert res.stats.gcups > 0 and res.stats.cells == float(db.num_residues) * len(q)
            assert "hom_q%d_" % qi in eng.getReferenceHeader(res.referenceIds[0])


def test_c1_full_config_all_queries(oracle):
    """BASELINE.json configs[0] / SURVEY.md 8(d) C1 at full size: 10,130 sequences (lognormal lengths, planted homologs,
    exact duplicates => score ties, non-standard letters), all 20 queries, BLOSUM62 -11/-1, top-10: every score and the
    top-10 list (ties by DB id) against the oracle (1.4e11 cells: ~15 s of CPU on the box's 16 cores)."""
    recs, queries = synth.config_c1(seed=1, n=10_000)
    db = dbformat.from_sequences([dbformat.encode(s) for _, s in recs], [h for h, _ in recs])
    with _engine(numTop=10, blosumType=62) as eng:
        eng.setDatabase(db)
        for qi, (_, q) in enumerate(queries):
            qc = dbformat.encode(q)
            res = eng.scan(q)
            scores, ids = eng.lastScanAllScores()
            got = np.empty(db.num_sequences, np.int32)
            got[ids] = scores
            ref = oracle.scan(62, qc, db, -11, -1)
            bad = np.nonzero(got != ref)[0]
            assert len(bad) == 0, (qi, bad[:8], got[bad[:8]], ref[bad[:8]], db.lengths[bad[:8]])
            s, i = oracle.topk(ref, 10)
            assert res.scores == s.tolist() and res.referenceIds == i.tolist(), qi
            L = db.lengths
            assert res.stats.numOverflows == int(((ref >= 25000) & (L > 240) & (L <= 8000)).sum()), qi


def test_long_query_against_short_subjects(oracle):
    """Query much longer than the subjects (long periods, many ring refills per alignment), odd and even lengths."""
    rng = np.random.default_rng(77)
    seqs = [synth.random_residues(rng, int(n)) for n in rng.integers(20, 520, 600)]
    db = dbformat.from_sequences(seqs)
    with _engine(numTop=10, blosumType=62) as eng:
        eng.setDatabase(db)
        for ql in (9001, 12000):
            q = synth.random_residues(rng, ql)
            q[100:100 + len(seqs[5])] = seqs[5]  # plant one subject inside the query
            res = eng.scan(dbformat.decode(q))
            scores, ids = eng.lastScanAllScores()
            got = np.empty(db.num_sequences, np.int32)
            got[ids] = scores
            ref = oracle.scan(62, q, db, -11, -1)
            assert (got == ref).all(), np.nonzero(got != ref)[0][:10]
            s, i = oracle.topk(ref, 10)
            assert res.scores == s.tolist() and res.referenceIds == i.tolist()


def test_scores_beyond_16_bit_are_exact(oracle):
    rng = np.random.default_rng(9)
    q = synth.random_residues(rng, 7000)
    q[::3] = 17  # tryptophan-rich: self score far above 32767
    seqs = [synth.random_residues(rng, int(n)) for n in rng.integers(1100, 3000, 20)]
    seqs += [q.copy(), synth.mutate(rng, q, 0.1), np.concatenate([synth.random_residues(rng, 80), q[:5000]])]
    db = dbformat.from_sequences(seqs)
    for blosum, gop, gex in ((45, -13, -2), (80, -10, -1)):
        with _engine(numTop=5, blosumType=blosum) as eng:
            eng.setGapScores(gop, gex)
            eng.setDatabase(db)
            res = eng.scan(dbformat.decode(q))
            ref = oracle.scan(blosum, q, db, gop, gex)
            s, i = oracle.topk(ref, 5)
            assert s[0] > 32767
            assert res.scores == s.tolist() and res.referenceIds == i.tolist()


def test_many_saturating_subjects_and_overflow_count(oracle):
    """More 16-bit overflows than SMs (the exact array kernel streams several subjects per CTA), long and short
    subjects among them; stats.numOverflows follows the reference: subjects of 241..8000 residues whose exact score
    reaches 25000 (longer ones go straight to its 32-bit kernel, src/cudasw4.cuh:2152-2169)."""
    rng = np.random.default_rng(33)
    q = synth.random_residues(rng, 5600)
    q[::4] = 17  # W scores 11: lifts the self score well above 25000
    seqs = [synth.mutate(rng, q, 0.02 + 0.002 * (i % 40)) for i in range(330)]
    seqs += [np.concatenate([synth.random_residues(rng, 1500), q, synth.random_residues(rng, 3000)]) for _ in range(6)]  # > 8000
    seqs += [q[:4000].copy(), q[1000:4800].copy()]
    seqs += [synth.random_residues(rng, int(n)) for n in rng.integers(300, 6000, 150)]
    db = dbformat.from_sequences(seqs)
    with _engine(numTop=20, blosumType=62) as eng:
        eng.setDatabase(db)
        for rep in range(2):
            res = eng.scan(dbformat.decode(q))
            scores, ids = eng.lastScanAllScores()
            got = np.empty(db.num_sequences, np.int32)
            got[ids] = scores
            ref = oracle.scan(62, q, db, -11, -1)
            bad = np.nonzero(got != ref)[0]
            assert len(bad) == 0, (rep, bad[:8], got[bad[:8]], ref[bad[:8]], db.lengths[bad[:8]])
            s, i = oracle.topk(ref, 20)
            assert res.scores == s.tolist() and res.referenceIds == i.tolist()
            L = db.lengths
            expect = int(((ref >= 25000) & (L > 240) & (L <= 8000)).sum())
            assert expect > 300 and int((ref >= 25000).sum()) > expect
            assert res.stats.numOverflows == expect


def test_many_long_subjects_repeated_scans(oracle):
    """Enough multi-segment subjects to spread the long class over many (and an odd number of) SMs, scanned several
    times with different queries: every score must match the oracle every time (no stale or shared border state)."""
    rng = np.random.default_rng(21)
    L = np.concatenate([rng.integers(1025, 3200, 2300), rng.integers(60, 900, 800)])
    seqs = [synth.random_residues(rng, int(n)) for n in L]
    qs = [synth.random_residues(rng, n) for n in (97, 310, 150)]
    seqs += [np.concatenate([q, synth.random_residues(rng, 1500)]) for q in qs]
    db = dbformat.from_sequences(seqs)
    with _engine(numTop=10, blosumType=62) as eng:
        eng.setDatabase(db)
        for rep in range(2):
            for q in qs:
                res = eng.scan(dbformat.decode(q))
                scores, ids = eng.lastScanAllScores()
                got = np.empty(db.num_sequences, np.int32)
                got[ids] = scores
                ref = oracle.scan(62, q, db, -11, -1)
                bad = np.nonzero(got != ref)[0]
                assert len(bad) == 0, (rep, len(q), bad[:8], got[bad[:8]], ref[bad[:8]], db.lengths[bad[:8]])
                s, i = oracle.topk(ref, 10)
                assert res.scores == s.tolist() and res.referenceIds == i.tolist()


def test_reconfigure_between_scans_and_empty_database(oracle):
    """One handle, settings changed between scans (matrix, gaps, top-k, database) like the reference's setters allow."""
    db, rng = _mixed_db(91, 900, 5, 700, [1500])
    q = synth.random_residues(rng, 256)
    with _engine(numTop=5, blosumType=62) as eng:
        eng.setDatabase(db)
        for blosum, gop, gex, k in ((62, -11, -1, 5), (45, -13, -2, 12), (80, -10, -1, 3), (50, -20, -5, 7)):
            eng.setBlosum(blosum)
            eng.setGapScores(gop, gex)
            eng.setNumTop(k)
            res = eng.scan(dbformat.decode(q))
            s, i = oracle.topk(oracle.scan(blosum, q, db, gop, gex), k)
            assert res.scores == s.tolist() and res.referenceIds == i.tolist(), (blosum, gop, gex)
        db2, _ = _mixed_db(92, 300, 30, 90)
        eng.setDatabase(db2)  # replaces the resident database
        res = eng.scan(dbformat.decode(q))
        s, i = oracle.topk(oracle.scan(50, q, db2, -20, -5), 7)
        assert res.scores == s.tolist() and res.referenceIds == i.tolist()
        eng.setDatabase(dbformat.from_sequences([]))
        res = eng.scan(dbformat.decode(q))
        assert res.scores == [] and res.referenceIds == []


def test_large_top_k_uses_exact_host_selection(oracle):
    db, rng = _mixed_db(55, 7000, 20, 300)
    q = synth.random_residues(rng, 180)
    ref = oracle.scan(62, q, db, -11, -1)
    for k in (4096, 5000, 100000):
        with _engine(numTop=k, blosumType=62) as eng:
            eng.setDatabase(db)
            res = eng.scan(dbformat.decode(q))
            s, i = oracle.topk(ref, k)
            assert len(res.scores) == min(k, db.num_sequences)
            assert res.scores == s.tolist() and res.referenceIds == i.tolist()


def test_edge_cases(oracle):
    rng = np.random.default_rng(3)
    seqs = [np.zeros(0, np.uint8), np.zeros(0, np.uint8), synth.random_residues(rng, 1), synth.random_residues(rng, 5)]
    db = dbformat.from_sequences(seqs)
    with _engine(numTop=10, blosumType=62) as eng:
        eng.setDatabase(db)
        res = eng.scan("ARNDW")
        ref = oracle.scan(62, dbformat.encode("ARNDW"), db, -11, -1)
        s, i = oracle.topk(ref, 10)
        assert len(res.scores) == 4 and res.scores == s.tolist() and res.referenceIds == i.tolist()
        res = eng.scan("")  # empty query: every score is 0, ids ascending
        assert res.scores == [0, 0, 0, 0] and res.referenceIds == [0, 1, 2, 3]
        res = eng.scan("xx*B")  # only 'other' letters
        assert max(res.scores) == 0
        eng.setNumTop(2)
        assert len(eng.scan("ARNDW").scores) == 2
    with pytest.raises(sw.SW4Error):
        _engine(blosumType=63)
    with _engine() as eng:
        with pytest.raises(sw.SW4Error):
            eng.scan("ARND")  # no database
        with pytest.raises(sw.SW4Error):
            eng.setGapScores(-100000, -1)
        with pytest.raises(sw.SW4Error):
            eng.setDatabase("/nonexistent/prefix")


def test_shards_partition_the_database(oracle):
    db, rng = _mixed_db(77, 3000, 20, 600, [1030, 1400])
    q = synth.random_residues(rng, 222)
    ref = oracle.scan(62, q, db, -11, -1)
    seen = np.zeros(db.num_sequences, bool)
    merged = []
    for rank in range(3):
        with _engine(numTop=15, blosumType=62) as eng:
            eng.setShard(rank, 3)
            eng.setDatabase(db)
            res = eng.scan(dbformat.decode(q))
            scores, ids = eng.lastScanAllScores()
            assert not seen[ids].any()
            seen[ids] = True
            assert (scores == ref[ids]).all()
            merged += list(zip(res.scores, res.referenceIds))
    assert seen.all()
    merged.sort(key=lambda t: (-t[0], t[1]))
    s, i = oracle.topk(ref, 15)
    assert [m[0] for m in merged[:15]] == s.tolist() and [m[1] for m in merged[:15]] == i.tolist()


def test_in_process_multi_gpu(oracle):
    """All visible GPUs driven from one handle, like the reference's single process (skipped on a 1-GPU box)."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    db, rng = _mixed_db(31, 6000, 10, 1400, [2500, 5000])
    qs = [synth.random_residues(rng, n) for n in (200, 431)]
    with sw.CudaSW4(deviceIds=list(range(ngpu)), numTop=20, blosumType=62) as eng:
        eng.setDatabase(db)
        for q in qs:
            res = eng.scan(dbformat.decode(q))
            scores, ids = eng.lastScanAllScores()
            assert sorted(ids.tolist()) == list(range(db.num_sequences))
            ref = oracle.scan(62, q, db, -11, -1)
            assert (scores == ref[ids]).all()
            s, i = oracle.topk(ref, 20)
            assert res.scores == s.tolist() and res.referenceIds == i.tolist()


def test_full_size_peak_config_properties():
    """BASELINE config[1] at full size (1M x 256): every subject is identical, so every score must equal the known
    answer, the checksum of scores is n * answer, and top-10 ids are 0..9."""
    with open(os.path.join(os.path.dirname(__file__), "golden", "survey_kat.json")) as f:
        kat = json.load(f)["pseudodb_blosum62_gop-11_gex-1"]["256"]
    queries = synth.load_queries()
    with _engine(numTop=10, blosumType=62) as eng:
        eng.setPseudoDatabase(1_000_000, 256)
        for qi in (0, 19):
            res = eng.scan(queries[qi][1])
            scores, ids = eng.lastScanAllScores()
            assert int(scores.astype(np.int64).sum()) == 1_000_000 * kat[qi]
            assert res.scores == [kat[qi]] * 10 and res.referenceIds == list(range(10))
            assert res.stats.numOverflows == 0
