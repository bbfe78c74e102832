"""CPU tests of the synthetic benchmark shapes (cudasw4_b200/synth.py): the generators bench.py and the parity tests rely
on must stay reproducible - bench.py's C4 leg re-creates sampled sequences on every rank from (seed, global id) alone."""
import numpy as np

from cudasw4_b200 import dbformat, synth


def test_pseudo_lengths_sequence_is_frozen():
    # the checker-side restatement of sw4_set_pseudo_database_lengths' generator (csrc/engine.cu); the GPU suite compares
    # the library's sequences with it (test_pseudo_database_with_lengths_is_reproducible)
    assert dbformat.decode(synth.pseudo_lengths_sequence(4, 12345, 50)) == "KADNQPNLSDSDAIAFVALIIIRTSSPFLYSPLDFPNSVIGASKNPARPG"
    assert dbformat.decode(synth.pseudo_lengths_sequence(77, 0, 20)) == "SVGGGRATGSVGQIPIIRKV"
    # a prefix property: the first residues do not depend on the length
    a, b = synth.pseudo_lengths_sequence(9, 7, 13), synth.pseudo_lengths_sequence(9, 7, 200)
    assert (a == b[:13]).all() and a.max() <= 19 and len(synth.pseudo_lengths_sequence(9, 7, 0)) == 0


def test_pseudo_residue_table_follows_background_frequencies():
    counts = np.bincount(synth._PSEUDO_TABLE, minlength=20)
    assert counts.sum() == 256 and counts.min() >= 3
    assert np.abs(counts / 256 - synth._FREQ).max() < 0.004   # 1/256 quantisation


def test_c4_length_law_is_reproducible():
    L = synth.config_c4_lengths(seed=4, n=100_000, total=17.0e9 * 100_000 / 65e6)
    assert (len(L), int(L.sum()), int(L[0]), int(L[-1]), int(L[50_000])) == (100_000, 26_166_247, 11, 5039, 191)
    assert (np.diff(L) >= 0).all()


def test_reference_pseudo_subject_matches_survey_vector():
    # the one subject `--pseudodb n 256` replicates (mt19937(42), libstdc++ uniform_int_distribution)
    s = dbformat.decode(synth.pseudo_subject(256, 42))
    assert s.startswith("GSVDPSKKDHDRRIWEMNPFARVPTYCADVDMEMLAHAQLMGNAQVGCIRSMDGLVKIAWMFDIRAYYVKTGEARCFCHF") and len(s) == 256
