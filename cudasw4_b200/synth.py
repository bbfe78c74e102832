"""Deterministic synthetic protein databases of the shapes BASELINE.json names (no network => no UniProt).

Shapes follow SURVEY.md §8(d): C1 (10k mixed, planted homologs/ties/odd letters), C2 (reference PseudoDB: 1M identical
subjects of length 256) and C2' (1M distinct), C3 (Swiss-Prot-shaped), C4 (UniRef50-shaped, generated already
length-sorted), C5 (long sequences). Residues are drawn from UniProt background frequencies.
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import dbformat

_FREQ = np.array([8.25, 5.53, 4.06, 5.45, 1.37, 3.93, 6.75, 7.07, 2.27, 5.96, 9.66, 5.84, 2.42, 3.86, 4.70, 6.56, 5.34,
                  1.08, 2.92, 6.87])
_FREQ = _FREQ / _FREQ.sum()
_CDF = np.cumsum(_FREQ)

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_queries() -> list[tuple[str, str]]:
    """The 20 benchmark queries of the reference (allqueries.fasta), carried as a fixture (header, letters)."""
    with open(os.path.join(_DATA, "allqueries.json")) as f:
        return [(r["header"], r["sequence"]) for r in json.load(f)["records"]]


def random_residues(rng: np.random.Generator, n: int) -> np.ndarray:
    """n residue codes (uint8, 0..19) from the background distribution."""
    u = rng.random(n, dtype=np.float32)
    return np.minimum(np.searchsorted(_CDF, u, side="right"), 19).astype(np.uint8)


def lognormal_lengths(rng, n, mu, sigma, lo, hi, total=None):
    L = np.clip(np.rint(rng.lognormal(mu, sigma, n)), lo, hi).astype(np.int64)
    if total is not None:  # rescale so that sum(len) ~= total
        L = np.clip(np.rint(L * (total / L.sum())), lo, hi).astype(np.int64)
    return L


def mutate(rng, codes: np.ndarray, sub_rate: float, indels: int = 2) -> np.ndarray:
    c = codes.copy()
    m = rng.random(len(c)) < sub_rate
    c[m] = random_residues(rng, int(m.sum()))
    for _ in range(indels):
        p = int(rng.integers(1, max(2, len(c) - 1)))
        if rng.random() < 0.5:
            c = np.concatenate([c[:p], random_residues(rng, int(rng.integers(1, 6))), c[p:]])
        else:
            c = np.concatenate([c[:p], c[p + int(rng.integers(1, 6)):]])
    return c


def pseudo_subject(length: int, seed: int = 42) -> np.ndarray:
    """The one subject the reference's `--pseudodb n len` replicates (src/dbdata.hpp:219-246): libstdc++
    mt19937(seed) + uniform_int_distribution<>(0,19) over "ARNDCQEGHILKMFPSTWYV". Re-implemented here in numpy:
    libstdc++ (GCC >= 11) maps a 32-bit engine draw x to a 20-value range with Lemire's multiply-shift method:
    value = (x * 20) >> 32, redrawing when the low 32 bits of the product are < 2^32 mod 20 = 16."""
    state = np.zeros(624, dtype=np.uint32)  # std::mt19937(seed) seeding (init_genrand)
    state[0] = seed & 0xFFFFFFFF
    for i in range(1, 624):
        prev = int(state[i - 1])
        state[i] = (1812433253 * (prev ^ (prev >> 30)) + i) & 0xFFFFFFFF
    mt = np.random.MT19937()
    mt.state = {"bit_generator": "MT19937", "state": {"key": state, "pos": 624}}
    out = np.empty(length, dtype=np.uint8)
    k = 0
    while k < length:
        prod = mt.random_raw(1).astype(np.uint64)[0] * np.uint64(20)
        if int(prod) & 0xFFFFFFFF < 16:
            continue
        out[k] = int(prod) >> 32
        k += 1
    return out


_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_PSEUDO_CUM = np.array([21, 35, 46, 60, 63, 73, 91, 109, 115, 130, 155, 170, 176, 186, 198, 214, 228, 231, 238, 256])
_PSEUDO_TABLE = np.searchsorted(_PSEUDO_CUM, np.arange(256), side="right").astype(np.uint8)


def _mix64(z: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic), as sw4_mix64 in csrc/engine.cu."""
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def pseudo_lengths_sequence(seed: int, gid: int, length: int) -> np.ndarray:
    """Residue codes of sequence `gid` of sw4_set_pseudo_database_lengths(seed): the checker-side restatement of the
    generator in cudasw4_b200/csrc/engine.cu."""
    key = _mix64(np.array([(seed + gid) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]
    words = (length + 7) // 8
    with np.errstate(over="ignore"):
        w = _mix64(key + np.arange(words, dtype=np.uint64))
    b = (w[:, None] >> (np.arange(8, dtype=np.uint64) * np.uint64(8))[None, :]) & np.uint64(0xFF)
    return _PSEUDO_TABLE[b.reshape(-1)[:length].astype(np.int64)]


def config_c1(seed: int = 1, n: int = 10_000):
    """C1: returns (records[(header, letters)] in FASTA order, queries) - mixed lengths with planted structure."""
    rng = np.random.default_rng(seed)
    queries = load_queries()
    lengths = lognormal_lengths(rng, n, 5.68, 0.64, 10, 6000)
    recs: list[tuple[str, str]] = []
    for i, L in enumerate(lengths):
        recs.append((f"syn|C1_{i:05d}| random len={L}", dbformat.decode(random_residues(rng, int(L)))))
    for qi, (_, q) in enumerate(queries):  # 3 mutated copies per query
        qc = dbformat.encode(q)
        for rate in (0.05, 0.20, 0.40):
            recs.append((f"syn|C1_hom_q{qi}_{int(rate*100)}| planted", dbformat.decode(mutate(rng, qc, rate))))
    for j in range(50):  # exact duplicates => score ties
        src = recs[int(rng.integers(0, n))]
        recs.append((f"syn|C1_dup_{j}| copy of {src[0].split('|')[1]}", src[1]))
    odd = "XBZU*acdx "
    for j in range(20):  # non-standard letters
        s = list(dbformat.decode(random_residues(rng, int(rng.integers(30, 400)))))
        for _ in range(8):
            s[int(rng.integers(0, len(s)))] = odd[int(rng.integers(0, len(odd)))]
        recs.append((f"syn|C1_odd_{j}| odd letters", "".join(s).replace(" ", "x")))
    perm = rng.permutation(len(recs))
    return [recs[i] for i in perm], queries


def config_c2(n: int = 1_000_000, length: int = 256, distinct: bool = False, seed: int = 2) -> dbformat.SequenceDB:
    """C2 = reference PseudoDB(n, length, 42) (all subjects identical); distinct=True gives C2' (honest variant)."""
    if distinct:
        rng = np.random.default_rng(seed)
        codes = random_residues(rng, n * length).reshape(n, length)
    else:
        codes = np.broadcast_to(pseudo_subject(length, 42), (n, length))
    return dbformat.from_equal_length_matrix(np.ascontiguousarray(codes))


def _db_from_sorted_lengths(rng, lengths: np.ndarray, planted: list[np.ndarray] | None = None,
                            pool: int | None = None) -> dbformat.SequenceDB:
    """Vectorised builder for big shapes: lengths must be ascending; residues random; `planted` sequences are
    written over the subjects whose length matches best."""
    lengths = np.sort(lengths).astype(np.int64)
    n = len(lengths)
    padded = (lengths + 3) // 4 * 4
    offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(padded, out=offsets[1:])
    total = int(offsets[-1])
    if pool is None:
        chars = random_residues(rng, total)
    else:  # very large shapes: tile a random pool at random rotations (subjects are distinct substrings of the pool)
        chars = np.empty(total, dtype=np.uint8)
        base = random_residues(rng, pool)
        for o in range(0, total, pool):
            m = min(pool, total - o)
            r = int(rng.integers(0, pool))
            chars[o:o + m] = np.roll(base, r)[:m]
    # padding bytes -> 20 (at most 3 per sequence)
    ends = offsets[:-1].astype(np.int64) + lengths
    for k in range(3):
        sel = (padded - lengths) > k
        chars[ends[sel] + k] = dbformat.PAD_CODE
    del ends
    if planted:
        used = set()
        for p in planted:
            i = int(np.searchsorted(lengths, len(p), side="left"))
            while i in used:
                i += 1
            if i >= n or lengths[i] != len(p):
                continue  # only plant when an equal-length slot exists
            used.add(i)
            o = int(offsets[i])
            chars[o:o + len(p)] = p
    return dbformat.SequenceDB(chars, offsets, lengths.astype(np.int32), np.full(n, ord("H"), dtype=np.uint8),
                               np.arange(n + 1, dtype=np.uint64))


def config_c3(seed: int = 3, n: int = 570_000, total: float = 205e6) -> dbformat.SequenceDB:
    """C3: Swiss-Prot-shaped (lognormal(5.683, 0.636), clipped [2, 35213], ~205 M residues, 25 very long subjects)."""
    rng = np.random.default_rng(seed)
    L = lognormal_lengths(rng, n - 25, 5.683, 0.636, 2, 35213, total=total)
    L = np.concatenate([L, rng.integers(8001, 35214, 25)])
    queries = [dbformat.encode(q) for _, q in load_queries()]
    planted = []
    for qc in queries:
        planted += [qc.copy(), mutate(rng, qc, 0.1, indels=0)]
    # make sure equal-length slots exist for the planted sequences
    L[: len(planted)] = [len(p) for p in planted]
    return _db_from_sorted_lengths(rng, L, planted)


def config_c4_lengths(seed: int = 4, n: int = 65_000_000, total: float = 17.0e9) -> np.ndarray:
    """C4: UniRef50-shaped length multiset (lognormal(5.247, 0.80), clipped [11, 45000]); sorted ascending."""
    rng = np.random.default_rng(seed)
    return np.sort(lognormal_lengths(rng, n, 5.247, 0.80, 11, 45000, total=total))


def config_c5(seed: int = 5, n_subjects: int = 2000, lo: int = 2000, hi: int = 35000,
              query_lengths=(2048, 3000, 5000, 8000, 8001, 12000, 20000, 35000)):
    """C5: long queries x long subjects with planted copies so exact scores exceed the s16 envelope."""
    rng = np.random.default_rng(seed)
    queries = [random_residues(rng, int(L)) for L in query_lengths]
    L = np.exp(rng.uniform(np.log(lo), np.log(hi), n_subjects)).astype(np.int64)
    seqs = [random_residues(rng, int(x)) for x in L]
    for q in queries:
        seqs.append(q.copy())                              # exact copy
        seqs.append(mutate(rng, q, 0.10, indels=2))        # 10 % mutated
        fl, fr = int(rng.integers(50, 400)), int(rng.integers(50, 400))
        seqs.append(np.concatenate([random_residues(rng, fl), q, random_residues(rng, fr)]))  # embedded
    return dbformat.from_sequences(seqs), queries
