"""Reader / writer for the reference's on-disk database format (numpy only, no GPU).

Format (reference src/makedb.cpp:182-276, src/dbdata.hpp:21-28, src/dbdata.cpp:48-128):
  <prefix>0chars          residue CODES 0..20, every sequence padded with 20 to a multiple of 4 bytes
  <prefix>0offsets        uint64[n+1], byte offsets into chars (including padding)
  <prefix>0lengths        int32[n], true lengths, ascending (sequences are sorted by length)
  <prefix>0headers        concatenated header bytes
  <prefix>0headeroffsets  uint64[n+1]
  <prefix>0metadata       int32 numPartitions(36), int32[36] boundaries, uint64[36] sequences per partition
  <prefix>metadata        empty file (src/dbdata.cpp:192-197)
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

LETTERS = "ARNDCQEGHILKMFPSTWYV"
PAD_CODE = 20

# length k is in partition i iff BOUNDARIES[i-1] < k <= BOUNDARIES[i]   (src/length_partitions.hpp:75-113)
BOUNDARIES = np.array(
    [48, 64, 80, 96, 112, 128, 144, 160, 176, 192, 208, 224, 240, 256, 288, 320, 352, 384, 416, 448, 480, 512, 576,
     640, 704, 768, 832, 896, 960, 1024, 1088, 1152, 1216, 1280, 8000, 2**31 - 2], dtype=np.int64)

_ENCODE = np.full(256, PAD_CODE, dtype=np.uint8)
for _i, _c in enumerate(LETTERS):
    _ENCODE[ord(_c)] = _i
_DECODE = np.frombuffer((LETTERS + "-").encode(), dtype=np.uint8)


def encode(seq: str | bytes | np.ndarray) -> np.ndarray:
    """letters -> codes; anything that is not one of the 20 upper-case letters becomes 20 (src/convert.cuh:6-34)."""
    if isinstance(seq, str):
        seq = seq.encode()
    if isinstance(seq, (bytes, bytearray)):
        seq = np.frombuffer(seq, dtype=np.uint8)
    return _ENCODE[seq]


def decode(codes: np.ndarray) -> str:
    """codes -> letters, 20 -> '-' (src/convert.cuh:36-64)."""
    return _DECODE[np.minimum(np.asarray(codes, dtype=np.uint8), 20)].tobytes().decode()


@dataclass
class SequenceDB:
    chars: np.ndarray          # uint8 codes, padded per sequence to a multiple of 4
    offsets: np.ndarray        # uint64[n+1]
    lengths: np.ndarray        # int32[n], ascending
    headers: np.ndarray        # uint8
    header_offsets: np.ndarray  # uint64[n+1]

    @property
    def num_sequences(self) -> int:
        return int(self.lengths.shape[0])

    @property
    def num_residues(self) -> int:
        return int(self.lengths.astype(np.int64).sum())

    def sequence(self, i: int) -> np.ndarray:
        o = int(self.offsets[i])
        return self.chars[o:o + int(self.lengths[i])]

    def header(self, i: int) -> str:
        return self.headers[int(self.header_offsets[i]):int(self.header_offsets[i + 1])].tobytes().decode(errors="replace")

    def partition_counts(self) -> np.ndarray:
        part = np.searchsorted(BOUNDARIES, self.lengths.astype(np.int64), side="left")
        return np.bincount(part, minlength=len(BOUNDARIES)).astype(np.uint64)


def from_sequences(seqs: list[np.ndarray], headers: list[str] | None = None, presorted: bool = False) -> SequenceDB:
    """Build the in-memory DB from code arrays. Sorted by length with a *stable* sort (the reference uses an unstable
    std::sort by length only, makedb.cpp:191-195, so the order inside one length class is implementation defined)."""
    n = len(seqs)
    lengths = np.array([len(s) for s in seqs], dtype=np.int32)
    order = np.arange(n) if presorted else np.argsort(lengths, kind="stable")
    if headers is None:
        headers = [f"seq{i}" for i in range(n)]
    lengths = lengths[order]
    padded = (lengths.astype(np.int64) + 3) // 4 * 4
    offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(padded, out=offsets[1:])
    chars = np.full(int(offsets[-1]), PAD_CODE, dtype=np.uint8)
    hb = [headers[i].encode() for i in order]
    header_offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum([len(h) for h in hb], out=header_offsets[1:])
    for k, i in enumerate(order):
        o = int(offsets[k])
        chars[o:o + len(seqs[i])] = seqs[i]
    return SequenceDB(chars, offsets, lengths, np.frombuffer(b"".join(hb), dtype=np.uint8).copy(), header_offsets)


def from_equal_length_matrix(codes: np.ndarray) -> SequenceDB:
    """codes: uint8 [n, L] -> DB of n subjects of length L (vectorised; used for the big synthetic shapes)."""
    n, L = codes.shape
    Lp = (L + 3) // 4 * 4
    chars = np.full((n, Lp), PAD_CODE, dtype=np.uint8)
    chars[:, :L] = codes
    offsets = (np.arange(n + 1, dtype=np.uint64) * np.uint64(Lp))
    return SequenceDB(chars.reshape(-1), offsets, np.full(n, L, dtype=np.int32),
                      np.full(n, ord("H"), dtype=np.uint8), np.arange(n + 1, dtype=np.uint64))


def write_db(prefix: str, db: SequenceDB) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    p0 = prefix + "0"
    db.chars.astype(np.uint8).tofile(p0 + "chars")
    db.offsets.astype("<u8").tofile(p0 + "offsets")
    db.lengths.astype("<i4").tofile(p0 + "lengths")
    db.headers.astype(np.uint8).tofile(p0 + "headers")
    db.header_offsets.astype("<u8").tofile(p0 + "headeroffsets")
    with open(p0 + "metadata", "wb") as f:
        np.array([len(BOUNDARIES)], dtype="<i4").tofile(f)
        BOUNDARIES.astype("<i4").tofile(f)
        db.partition_counts().astype("<u8").tofile(f)
    open(prefix + "metadata", "wb").close()


def read_db(prefix: str) -> SequenceDB:
    p0 = prefix + "0"
    return SequenceDB(
        np.fromfile(p0 + "chars", dtype=np.uint8),
        np.fromfile(p0 + "offsets", dtype="<u8"),
        np.fromfile(p0 + "lengths", dtype="<i4"),
        np.fromfile(p0 + "headers", dtype=np.uint8),
        np.fromfile(p0 + "headeroffsets", dtype="<u8"),
    )


def read_fasta(path: str) -> list[tuple[str, str]]:
    """Minimal FASTA reader: header = the whole line after '>' (kseqpp keeps the full header line)."""
    import gzip
    opener = gzip.open if path.endswith(".gz") else open
    out: list[tuple[str, str]] = []
    name, chunks = None, []
    with opener(path, "rt") as f:
        for line in f:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if name is not None:
                    out.append((name, "".join(chunks)))
                name, chunks = line[1:], []
            elif name is not None:
                chunks.append(line.strip())
    if name is not None:
        out.append((name, "".join(chunks)))
    return out


def write_fasta(path: str, records: list[tuple[str, str]], width: int = 60) -> None:
    with open(path, "w") as f:
        for h, s in records:
            f.write(">" + h + "\n")
            for i in range(0, len(s), width):
                f.write(s[i:i + width] + "\n")
