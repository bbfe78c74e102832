"""ctypes binding of the CPU oracle (oracle/liboracle.so). TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs - never by the product."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = os.path.join(ROOT, "oracle", "liboracle.so")


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.sw4o_gotoh_score.restype = ctypes.c_int
        lib.sw4o_scan.restype = ctypes.c_int
        lib.sw4o_topk.restype = ctypes.c_long
        lib.sw4o_gcups.restype = ctypes.c_double
        lib.sw4o_gcups.argtypes = [ctypes.c_double, ctypes.c_double]

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(ctypes.c_void_p)

    def convert(self, letters: bytes) -> np.ndarray:
        out = np.empty(len(letters), dtype=np.uint8)
        self.lib.sw4o_convert_letters(letters, self._p(out), ctypes.c_long(len(letters)))
        return out

    def matrix(self, blosum: int) -> np.ndarray:
        out = np.empty(441, dtype=np.int8)
        assert self.lib.sw4o_substitution_matrix(blosum, self._p(out)) == 0
        return out.reshape(21, 21)

    def score(self, blosum, q, s, gop, gex) -> int:
        q = np.ascontiguousarray(q, dtype=np.uint8)
        s = np.ascontiguousarray(s, dtype=np.uint8)
        return int(self.lib.sw4o_gotoh_score(blosum, self._p(q), len(q), self._p(s), len(s), gop, gex))

    def scan(self, blosum, q, db, gop, gex, threads=0, subset=None):
        """All subject scores of a dbformat.SequenceDB (or of the index subset)."""
        q = np.ascontiguousarray(q, dtype=np.uint8)
        chars = np.ascontiguousarray(db.chars, dtype=np.uint8)
        offsets = np.ascontiguousarray(db.offsets, dtype=np.uint64)
        lengths = np.ascontiguousarray(db.lengths, dtype=np.int32)
        if subset is not None:
            subset = np.asarray(subset)
            offsets = np.ascontiguousarray(offsets[subset])
            lengths = np.ascontiguousarray(lengths[subset])
        n = len(lengths)
        out = np.empty(n, dtype=np.int32)
        used = self.lib.sw4o_scan(blosum, self._p(q), len(q), self._p(chars), self._p(offsets), self._p(lengths),
                                  ctypes.c_long(n), gop, gex, self._p(out), threads)
        assert used > 0
        self.last_threads = used
        return out

    def topk(self, scores: np.ndarray, k: int):
        scores = np.ascontiguousarray(scores, dtype=np.int32)
        m = min(k, len(scores))
        so = np.empty(m, dtype=np.int32)
        io = np.empty(m, dtype=np.int32)
        got = self.lib.sw4o_topk(self._p(scores), ctypes.c_long(len(scores)), ctypes.c_long(k), self._p(so), self._p(io))
        assert got == m
        return so, io

    def pseudo_subject(self, length, seed=42):
        out = np.empty(length, dtype=np.uint8)
        self.lib.sw4o_pseudo_subject(length, seed, self._p(out))
        return out

    def partition(self, length):
        return int(self.lib.sw4o_length_partition(length))


def load() -> Oracle:
    src = os.path.join(ROOT, "oracle", "sw_oracle.cpp")
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, stdout=subprocess.DEVNULL)
    return Oracle(ctypes.CDLL(_LIB))
