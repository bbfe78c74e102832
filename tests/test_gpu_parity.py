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
    """All visible GPUs driven from one handle, like the reference's single process (on a 1-GPU box: the same GPU
    twice, which runs the same host threads, shard assignment and merge)."""
    import torch
    ngpu = torch.cuda.device_count()
    devices = list(range(ngpu)) if ngpu >= 2 else [0, 0]
    db, rng = _mixed_db(31, 6000, 10, 1400, [2500, 5000])
    qs = [synth.random_residues(rng, n) for n in (200, 431)]
    with sw.CudaSW4(deviceIds=devices, numTop=20, blosumType=62) as eng:
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


# ---------------------------------------------------------------------------------------------------------------------
# round 2: multi-shard on one GPU, query batching, streaming mode, pre-sharded databases, device top-k of any length,
# and the BASELINE configurations C3 / C5 at (or near) their specified shapes
# ---------------------------------------------------------------------------------------------------------------------
def _check_all(eng, oracle, db, q, blosum, gop, gex, k, res=None):
    res = res or eng.scan(dbformat.decode(q))
    scores, ids = eng.lastScanAllScores()
    assert sorted(ids.tolist()) == list(range(db.num_sequences))
    ref = oracle.scan(blosum, q, db, gop, gex)
    got = np.empty(db.num_sequences, np.int32)
    got[ids] = scores
    bad = np.nonzero(got != ref)[0]
    assert len(bad) == 0, (len(q), bad[:8], got[bad[:8]], ref[bad[:8]], db.lengths[bad[:8]])
    s, i = oracle.topk(ref, k)
    assert res.scores == s.tolist() and res.referenceIds == i.tolist()
    return ref


def test_multi_shard_handle_on_one_gpu(oracle):
    """The in-process multi-GPU driver (one host thread per shard, host merge of the per-shard top-k lists; the
    analogue of src/cudasw4.cuh:1490-2262, 1415-1458) exercised on a single GPU: the same device id three times."""
    db, rng = _mixed_db(31, 6000, 10, 1400, [2500, 5000])
    qs = [synth.random_residues(rng, n) for n in (200, 431)]
    with sw.CudaSW4(deviceIds=[0, 0, 0], numTop=20, blosumType=62) as eng:
        eng.setDatabase(db)
        assert eng.dbInfo().shard_sequences == db.num_sequences
        for q in qs:
            _check_all(eng, oracle, db, q, 62, -11, -1, 20)
        res, total = eng.scanMany([dbformat.decode(q) for q in qs])
        for q, r in zip(qs, res):
            s, i = oracle.topk(oracle.scan(62, q, db, -11, -1), 20)
            assert r.scores == s.tolist() and r.referenceIds == i.tolist()


def test_scan_many_equals_single_scans(oracle):
    """sw4_scan_many: several queries in flight; results must equal one scan per query."""
    db, rng = _mixed_db(41, 5000, 5, 1200, [1500, 2600, 4000])
    lens = (1, 17, 144, 333, 600, 97, 1500, 256, 40)
    qs = [synth.random_residues(rng, n) for n in lens]
    qs[6][200:200 + 1200] = db.sequence(db.num_sequences - 3)[:1200]  # a strong hit on a long subject
    with _engine(numTop=12, blosumType=62) as eng:
        eng.setDatabase(db)
        singles = [eng.scan(dbformat.decode(q)) for q in qs]
        many, total = eng.scanMany([dbformat.decode(q) for q in qs])
        assert len(many) == len(qs)
        for q, a, b in zip(qs, singles, many):
            assert a.scores == b.scores and a.referenceIds == b.referenceIds
            assert a.stats.numOverflows == b.stats.numOverflows and b.stats.cells == a.stats.cells
            s, i = oracle.topk(oracle.scan(62, q, db, -11, -1), 12)
            assert b.scores == s.tolist() and b.referenceIds == i.tolist()
        assert total.cells == sum(r.stats.cells for r in many) and total.gcups > 0
        # the context of the last query holds its scores
        _check_all(eng, oracle, db, qs[-1], 62, -11, -1, 12, res=many[-1])
        assert eng.scanMany([])[0] == []


def test_streaming_mode_equals_resident(oracle):
    """a database larger than maxGpuMem is streamed in batches on every scan (two device slots, upload of
    the next batch overlapping the kernels); scores, top-k and overflow statistics must equal the resident run."""
    rng = np.random.default_rng(61)
    q = synth.random_residues(rng, 5600)
    q[::4] = 17
    L = np.concatenate([rng.integers(1, 1300, 30000), rng.integers(1300, 9000, 300)])
    seqs = [synth.random_residues(rng, int(n)) for n in L]
    seqs += [synth.mutate(rng, q, 0.03) for _ in range(12)] + [np.concatenate([synth.random_residues(rng, 2000), q, synth.random_residues(rng, 900)])]
    db = dbformat.from_sequences(seqs)
    qs = [q, synth.random_residues(rng, 150), synth.random_residues(rng, 700)]
    with _engine(numTop=15, blosumType=62) as eng:
        eng.setDatabase(db)
        resident = [eng.scan(dbformat.decode(x)) for x in qs]
        assert eng.dbInfo().streaming == 0 and eng.dbInfo().num_batches == 1
        refs = [oracle.scan(62, x, db, -11, -1) for x in qs]
        for x, r, ref in zip(qs, resident, refs):
            s, i = oracle.topk(ref, 15)
            assert r.scores == s.tolist() and r.referenceIds == i.tolist()
        assert resident[0].stats.numOverflows >= 12
    for max_mem in (24 << 20, 14 << 20):
        mc = sw.MemoryConfig(maxGpuMem=max_mem)
        with _engine(numTop=15, blosumType=62, memoryConfig=mc) as eng:
            eng.setDatabase(db)
            for rep in range(2):
                for x, r, ref in zip(qs, resident, refs):
                    res = eng.scan(dbformat.decode(x))
                    assert res.scores == r.scores and res.referenceIds == r.referenceIds
                    assert res.stats.numOverflows == r.stats.numOverflows
                    scores, ids = eng.lastScanAllScores()
                    got = np.empty(db.num_sequences, np.int32)
                    got[ids] = scores
                    assert (got == ref).all(), np.nonzero(got != ref)[0][:10]
            info = eng.dbInfo()
            assert info.streaming == 1 and info.num_batches >= 3, (info.streaming, info.num_batches)
            many, _ = eng.scanMany([dbformat.decode(x) for x in qs] * 2)  # one pass over the batches serves all six
            for r, m in zip(resident * 2, many):
                assert m.scores == r.scores and m.referenceIds == r.referenceIds and m.stats.numOverflows == r.stats.numOverflows
            # switching the memory configuration re-plans the layout
            eng.setMemoryConfig(sw.MemoryConfig())
            res = eng.scan(dbformat.decode(qs[1]))
            assert res.scores == resident[1].scores and eng.dbInfo().streaming == 0


def test_streaming_multi_shard(oracle):
    db, rng = _mixed_db(67, 20000, 8, 900, [1800, 3000])
    q = synth.random_residues(rng, 280)
    with sw.CudaSW4(deviceIds=[0, 0], numTop=10, blosumType=62, memoryConfig=sw.MemoryConfig(maxGpuMem=3 << 20)) as eng:
        eng.setDatabase(db)
        _check_all(eng, oracle, db, q, 62, -11, -1, 10)
        assert eng.dbInfo().streaming == 1 and eng.dbInfo().num_batches >= 2


def test_presharded_database(oracle):
    """Every rank holds only its own shard in host memory (sw4_set_database_shard_memory): global ids in the results,
    merged lists equal the oracle's over the whole database."""
    db, rng = _mixed_db(83, 5000, 10, 1300, [2200])
    q = synth.random_residues(rng, 310)
    ref = oracle.scan(62, q, db, -11, -1)
    world = 3
    merged, seen = [], np.zeros(db.num_sequences, bool)
    for rank in range(world):
        gids = np.array([i for i in range(db.num_sequences) if (i // 256) % world == rank], dtype=np.int32)
        local = dbformat.from_sequences([db.sequence(int(i)) for i in gids])
        with _engine(numTop=15, blosumType=62) as eng:
            eng.setDatabaseShard(local, gids, db.num_sequences)
            res = eng.scan(dbformat.decode(q))
            scores, ids = eng.lastScanAllScores()
            assert ids.tolist() == gids.tolist() and (scores == ref[gids]).all()
            seen[ids] = True
            merged += list(zip(res.scores, res.referenceIds))
            assert eng.getReferenceLength(int(gids[5])) == int(db.lengths[gids[5]])
            assert eng.dbInfo().num_sequences == db.num_sequences and eng.dbInfo().shard_sequences == len(gids)
    assert seen.all()
    merged.sort(key=lambda t: (-t[0], t[1]))
    s, i = oracle.topk(ref, 15)
    assert [m[0] for m in merged[:15]] == s.tolist() and [m[1] for m in merged[:15]] == i.tolist()


def test_device_top_k_of_any_length():
    """k above the shared-memory path (4096) is selected and sorted on the device too (the reference accepts any --top,
    src/cudasw4.cuh:1365-1401); heavy ties must come out in ascending id order."""
    queries = synth.load_queries()
    with _engine(numTop=6000, blosumType=62) as eng:
        eng.setPseudoDatabase(20001, 128)
        res = eng.scan(queries[3][1])
        assert res.scores == [29] * 6000 and res.referenceIds == list(range(6000))
        eng.setNumTop(20001)
        res = eng.scan(queries[3][1])
        assert res.scores == [29] * 20001 and res.referenceIds == list(range(20001))
    with sw.CudaSW4(deviceIds=[0, 0], numTop=9000, blosumType=62) as eng:
        eng.setPseudoDatabase(20001, 128)
        res = eng.scan(queries[3][1])
        assert res.scores == [29] * 9000 and res.referenceIds == list(range(9000))


def test_gap_setters_keep_the_other_score(oracle):
    db, rng = _mixed_db(97, 400, 30, 400)
    q = synth.random_residues(rng, 200)
    with _engine(numTop=5, blosumType=45, gop=-13, gex=-2) as eng:
        eng.setDatabase(db)
        eng.setGapOpenScore(-10)   # must keep gex = -2
        _check_all(eng, oracle, db, q, 45, -10, -2, 5)
        eng.setGapExtendScore(-3)  # must keep gop = -10
        _check_all(eng, oracle, db, q, 45, -10, -3, 5)
        with pytest.raises(sw.SW4Error):
            eng.setGapOpenScore(-100000)
        _check_all(eng, oracle, db, q, 45, -10, -3, 5)  # a rejected value changes nothing


def test_c3_shaped_sample(oracle):
    """BASELINE configs[2] (Swiss-Prot-shaped) as a 50,000-subject sample of the same length law incl. subjects above
    8000 residues and planted homologs: every score, top-10 and overflow count vs the oracle for a short and a long query."""
    rng = np.random.default_rng(3)
    L = synth.lognormal_lengths(rng, 50_000, 5.683, 0.636, 2, 35213)
    L = np.concatenate([L, rng.integers(8001, 20000, 6)])
    queries = [dbformat.encode(q) for _, q in synth.load_queries()]
    planted = [queries[0].copy(), synth.mutate(rng, queries[0], 0.1, 0), queries[12].copy(), synth.mutate(rng, queries[12], 0.1, 0)]
    L[: len(planted)] = [len(p) for p in planted]
    db = synth._db_from_sorted_lengths(rng, L, planted)
    with _engine(numTop=10, blosumType=62) as eng:
        eng.setDatabase(db)
        for qi in (0, 12):
            res = eng.scan(dbformat.decode(queries[qi]))
            ref = _check_all(eng, oracle, db, queries[qi], 62, -11, -1, 10, res=res)
            Ls = db.lengths
            assert res.stats.numOverflows == int(((ref >= 25000) & (Ls > 240) & (Ls <= 8000)).sum())
            assert res.scores[0] == int(oracle.scan(62, queries[qi], dbformat.from_sequences([queries[qi]]), -11, -1)[0])


@pytest.mark.parametrize("blosum,gop,gex", [(45, -9, -3), (80, -14, -2)])
def test_c5_at_spec_long_query_custom_gaps(oracle, blosum, gop, gex):
    """BASELINE configs[4] at its largest shape: a 35,000-residue query against 20-35 k subjects with planted copies
    (exact, 10 % mutated, embedded in flanks), custom gap scores on the saturating path: scores far above 32767."""
    rng = np.random.default_rng(5)
    q = synth.random_residues(rng, 35000)
    seqs = [synth.random_residues(rng, int(n)) for n in (20000, 23500, 27111, 31000, 35000)]
    fl, fr = 211, 97
    seqs += [q.copy(), synth.mutate(rng, q, 0.10, indels=2), np.concatenate([synth.random_residues(rng, fl), q[:30000], synth.random_residues(rng, fr)])]
    seqs += [synth.random_residues(rng, int(n)) for n in (300, 1024, 1025, 5000)]
    db = dbformat.from_sequences(seqs)
    with _engine(numTop=6, blosumType=blosum) as eng:
        eng.setGapScores(gop, gex)
        eng.setDatabase(db)
        ref = _check_all(eng, oracle, db, q, blosum, gop, gex, 6)
        assert int(ref.max()) > 100000 and int((ref > 32767).sum()) >= 3


def test_pseudo_database_with_lengths_is_reproducible(oracle):
    """sw4_set_pseudo_database_lengths: every rank generates its own shard of the same database; the checker-side
    restatement (synth.pseudo_lengths_sequence) re-creates any sequence; planted sequences replace their slot."""
    rng = np.random.default_rng(4)
    lengths = np.sort(synth.lognormal_lengths(rng, 3000, 5.247, 0.80, 11, 45000)).astype(np.int32)
    q = synth.random_residues(rng, 180)
    gid = int(np.searchsorted(lengths, 180))
    lengths[gid] = 180
    lengths = np.sort(lengths)
    gid = int(np.searchsorted(lengths, 180))
    seqs = [synth.pseudo_lengths_sequence(77, g, int(lengths[g])) for g in range(len(lengths))]
    seqs[gid] = q
    db = dbformat.from_sequences(seqs, presorted=True)
    ref = oracle.scan(62, q, db, -11, -1)
    merged = []
    for rank in range(2):
        with _engine(numTop=10, blosumType=62) as eng:
            eng.setShard(rank, 2)
            eng.setPseudoDatabaseLengths(lengths, 77, {gid: q})
            res = eng.scan(dbformat.decode(q))
            scores, ids = eng.lastScanAllScores()
            assert ((ids // 256) % 2 == rank).all() and (scores == ref[ids]).all()
            assert eng.getReferenceSequence(int(ids[7])) == dbformat.decode(seqs[int(ids[7])])
            merged += list(zip(res.scores, res.referenceIds))
    merged.sort(key=lambda t: (-t[0], t[1]))
    s, i = oracle.topk(ref, 10)
    assert [m[0] for m in merged[:10]] == s.tolist() and [m[1] for m in merged[:10]] == i.tolist()
    assert merged[0][1] == gid


def test_border_rows_never_leak_stale_entries(oracle):
    """The border scratch of the long-subject kernels is reused across items, segments, queries and scans. With the
    buffer poisoned (SW4_DEBUG_POISON_BORDER: every entry a large positive score) any read of an entry the current
    alignment did not write itself would raise a score. Run in a subprocess (the switch is read by the library)."""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import cudasw4_b200 as sw
from cudasw4_b200 import dbformat, synth
from tests import oracle_lib
oracle = oracle_lib.load()
rng = np.random.default_rng(12)
L = np.concatenate([rng.integers(513, 1025, 900), rng.integers(1025, 6000, 500), rng.integers(20, 512, 300)])
db = dbformat.from_sequences([synth.random_residues(rng, int(n)) for n in L])
with sw.CudaSW4(deviceIds=[0], numTop=10, blosumType=62) as eng:
    eng.setDatabase(db)
    for ql in (3001, 97, 640, 1530, 33):
        q = synth.random_residues(rng, ql)
        res = eng.scan(dbformat.decode(q))
        scores, ids = eng.lastScanAllScores()
        got = np.empty(db.num_sequences, np.int32); got[ids] = scores
        ref = oracle.scan(62, q, db, -11, -1)
        bad = np.nonzero(got != ref)[0]
        assert len(bad) == 0, (ql, bad[:8], got[bad[:8]], ref[bad[:8]], db.lengths[bad[:8]])
print("ok")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in ({}, {"SW4_LONG_ARRAY_ITEMS_PER_GROUP": "0"}, {"SW4_NO_TWO_ROW_MULTI": "1"}):
        env = dict(os.environ, SW4_DEBUG_POISON_BORDER="1", **extra)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "ok" in r.stdout, (extra, r.stdout[-500:], r.stderr[-1500:])


def test_benchmark_mode_without_result_lists(oracle):
    """`--top 0` (the reference's benchmark mode, runpeakbenchmark.sh:27): scans run, nothing is selected or returned."""
    db, rng = _mixed_db(5, 800, 20, 700)
    qs = [dbformat.decode(synth.random_residues(rng, n)) for n in (50, 300)]
    with _engine(numTop=0, blosumType=62) as eng:
        eng.setDatabase(db)
        res = eng.scan(qs[1])
        assert res.scores == [] and res.referenceIds == [] and res.stats.gcups > 0
        scores, ids = eng.lastScanAllScores()
        ref = oracle.scan(62, dbformat.encode(qs[1]), db, -11, -1)
        assert (scores == ref[ids]).all()
        many, total = eng.scanMany(qs + [""])
        assert [m.scores for m in many] == [[], [], []] and total.cells == sum(m.stats.cells for m in many)


def test_huge_query_and_tight_scratch(oracle):
    """A 70,000-residue query (profile of 140 MB, periods of 35,000 steps, border rows of 0.5 MB each) and the same
    database under a border-scratch budget so small that the multi-segment kernels get fewer slots than CTAs (they then
    wait for one another) or none at all (fallback to the one-warp-per-pair kernel)."""
    rng = np.random.default_rng(70)
    L = np.concatenate([rng.integers(40, 1100, 260), rng.integers(1100, 4000, 40), [20000]])
    seqs = [synth.random_residues(rng, int(n)) for n in L]
    q = synth.random_residues(rng, 70000)
    q[30000:30000 + len(seqs[270])] = seqs[270]  # one planted subject
    db = dbformat.from_sequences(seqs)
    ref = oracle.scan(62, q, db, -11, -1)
    s, i = oracle.topk(ref, 10)
    with _engine(numTop=10, blosumType=62) as eng:
        eng.setDatabase(db)
        res = eng.scan(dbformat.decode(q))
        scores, ids = eng.lastScanAllScores()
        assert (scores == ref[ids]).all(), np.nonzero(scores != ref[ids])[0][:8]
        assert res.scores == s.tolist() and res.referenceIds == i.tolist()
    q2 = q[29000:33000]
    ref2 = oracle.scan(62, q2, db, -11, -1)
    for temp in (24 << 20, 3 << 20):   # 4000-residue query: 33 KB per border row array
        with _engine(numTop=10, blosumType=62, memoryConfig=sw.MemoryConfig(maxTempBytes=temp)) as eng:
            eng.setDatabase(db)
            for rep in range(2):
                eng.scan(dbformat.decode(q2))
                scores, ids = eng.lastScanAllScores()
                assert (scores == ref2[ids]).all(), (temp, np.nonzero(scores != ref2[ids])[0][:8])
    with _engine(numTop=10, blosumType=62, memoryConfig=sw.MemoryConfig(maxTempBytes=256 << 10)) as eng:
        eng.setDatabase(db)
        with pytest.raises(sw.SW4Error):   # not even 16 row arrays fit: refused up front, not mid-scan
            eng.scan(dbformat.decode(q2))
