"""Pins the CPU oracle against every known-answer vector the reference's own
unit tests hold for the k-mer counting path (caseywdunn/sharkmer v3.1.0):

  src/kmer/mod.rs:26-305      encoding / packing / N-splitting / histogram KATs
  src/kmer/counting.rs:365-510  KmerCounts behaviour
  src/pcr/mod.rs:1236-1342    18S x10 end-to-end table (string read from the
                              reference checkout when present; digest pinned
                              in tests/golden/)

plus an independent brute-force Python statement (sort | uniq -c).
"""
import json
import os
import random
import re

import numpy as np
import pytest

U32_MAX = 0xFFFFFFFF


# ---- src/kmer/mod.rs -------------------------------------------------------

def test_seq_to_reads_packing(oracle):  # mod.rs:61-111 (test_seq_to_reads, test_from_str)
    assert oracle.read_pack("CGTAATGCGGCGA") == ([0b01101100, 0b00111001, 0b10100110, 0b00000000], 13)
    assert oracle.read_pack("C") == ([0b01000000], 1)
    assert oracle.read_pack("CGTAATGCGGCG") == ([0b01101100, 0b00111001, 0b10100110], 12)
    assert oracle.read_pack("") == ([], 0)


def _subreads(seq):
    return [oracle_pack for oracle_pack in seq.split("N") if oracle_pack]


def test_seq_to_reads_n(oracle):  # mod.rs:113-156
    a = ([0b01101100], 4)
    b = ([0b00111001, 0b10100110, 0b00000000], 9)
    assert [oracle.read_pack(s) for s in _subreads("NCGTAATGCGGCG")] == [([0b01101100, 0b00111001, 0b10100110], 12)]
    for seq in ("CGTANATGCGGCGA", "NCGTANATGCGGCGA", "NCGTANATGCGGCGANN", "NNCGTANATGCGGCGA"):
        assert [oracle.read_pack(s) for s in _subreads(seq)] == [a, b]


def test_revcomp_kmer(oracle):  # mod.rs:158-177
    kmer = 0b0010_0110
    rc = oracle.revcomp_kmer(kmer, 3)
    assert oracle.revcomp_kmer(rc, 3) == kmer
    assert rc == 0b0001_1001
    kmer = 0b0110_1100_0011_1001_1010_0110
    rc = oracle.revcomp_kmer(kmer, 12)
    assert oracle.revcomp_kmer(rc, 12) == kmer
    assert rc == 0b0110_0101_1001_0011_1100_0110


def test_revcomp_matches_bit_parallel_and_python(oracle):
    """The byte-LUT walk (encoding.rs:235-262), skm_common.h's bit-parallel
    version (used on the device) and a naive loop agree for all k."""
    from sharkmer_b200 import common
    rng = random.Random(7)
    for k in range(1, 32):
        for _ in range(50):
            x = rng.getrandbits(2 * k)
            want = oracle.py_revcomp(x, k)
            assert oracle.revcomp_kmer(x, k) == want
            assert common.revcomp_kmer(x, k) == want


def test_get_kmers(oracle):  # mod.rs:179-226
    ints = [0b01101100, 0b00111001, 0b10100110]
    e = [0b01_1001_0011_1100_0110, 0b01_0110_0100_1111_0001, 0b10_0101_1001_0011_1100,
         0b00_0011_1001_1010_0110]
    assert oracle.read_get_kmers(ints, 12, 9) == e
    assert oracle.read_get_kmers(ints, 11, 9) == e[:3]
    assert oracle.read_get_kmers(ints, 10, 9) == e[:2]
    assert oracle.read_get_kmers(ints, 9, 9) == e[:1]
    assert oracle.read_get_kmers([0b01101100, 0b00111001], 8, 9) == []
    # the hot path gives the same values straight from ASCII
    assert oracle.kmers_from_ascii("CGTAATGCGGCG", 9) == e


def test_kmer_to_seq(oracle):  # mod.rs:228-237
    assert oracle.kmer_to_seq(0b1001_1000, 4) == "GCGA"
    assert oracle.kmer_to_seq(0b1001_1000_1001_1000, 8) == "GCGAGCGA"
    assert oracle.seq_to_kmer("GCGAGCGA") == 0b1001_1000_1001_1000


CASES = ["CGTAATGCGGCGA", "CGTANATGCGGCGA", "NCGTANATGCGGCGA", "NCGTANATGCGGCGANN",
         "NNCGTANATGCGGCGA", "TANCACN", "NTANCACNAGAAAATC", "AAAA", "ACGTACGTACGT"]


def test_kmers_from_ascii_matches_read_pipeline(oracle):  # mod.rs:249-270
    for k in (3, 5, 9, 11):
        for seq in CASES:
            got = oracle.kmers_from_ascii(seq, k)
            assert got == oracle.kmers_via_reads(seq, k), (seq, k)
            assert got == oracle.py_kmers(seq, k), (seq, k)


def test_kmers_from_ascii_short_sequences(oracle):  # mod.rs:272-278
    assert oracle.kmers_from_ascii("ACGT", 9) == []
    assert len(oracle.kmers_from_ascii("ACGTACGTA", 9)) == 1


def test_kmers_from_ascii_errors(oracle):  # encoding.rs:333, 353-356
    for bad in ("ACGTaCGT", "ACGTRACGT", "ACG TACG", "ACGT\n"):
        with pytest.raises(oracle.OracleError) as e:
            oracle.kmers_from_ascii(bad, 3)
        assert e.value.code == -1
    for k in (0, 32, 33):
        with pytest.raises(oracle.OracleError) as e:
            oracle.kmers_from_ascii("ACGT", k)
        assert e.value.code == -2


def test_count_valid_bases(oracle):  # mod.rs:280-286
    assert oracle.count_valid_bases("ACGTACGT") == 8
    assert oracle.count_valid_bases("ACNGT") == 4
    assert oracle.count_valid_bases("NNN") == 0
    assert oracle.count_valid_bases("") == 0


def test_histogram(oracle):  # mod.rs:288-305
    kc = oracle.KmerCounts(11)
    for kmer, c in ((1, 5), (20, 5), (2, 7), (11, 11), (12, 12)):
        kc.insert(kmer, c)
    v = oracle.Histogram.from_kmer_counts(kc, 10).get_vector()
    assert len(v) == 12
    assert v.tolist() == [0, 0, 0, 0, 0, 2, 0, 1, 0, 0, 0, 2]


# ---- src/kmer/counting.rs ----------------------------------------------------

def test_new_and_basic_ops(oracle):  # counting.rs:365-372
    kc = oracle.KmerCounts(5)
    assert kc.get_k() == 5 and kc.is_empty() and kc.len() == 0 and kc.get_n_kmers() == 0


def test_insert_get_accumulate(oracle):  # counting.rs:374-391
    kc = oracle.KmerCounts(5)
    kc.insert(42, 3)
    assert not kc.is_empty() and kc.get_count(42) == 3 and kc.contains(42) and not kc.contains(99)
    kc.insert(42, 7)
    assert kc.get_count(42) == 10 and kc.len() == 1


def test_saturating_add(oracle):  # counting.rs:393-399
    kc = oracle.KmerCounts(5)
    kc.insert(1, U32_MAX)
    kc.insert(1, 1)
    assert kc.get_count(1) == U32_MAX


def test_extend(oracle):  # counting.rs:401-422
    a, b = oracle.KmerCounts(5), oracle.KmerCounts(5)
    a.insert(1, 10); a.insert(2, 20); b.insert(2, 5); b.insert(3, 15)
    a.extend(b)
    assert (a.get_count(1), a.get_count(2), a.get_count(3)) == (10, 25, 15)
    with pytest.raises(oracle.OracleError):
        a.extend(oracle.KmerCounts(7))


def test_median_max(oracle):  # counting.rs:424-456
    kc = oracle.KmerCounts(5)
    assert kc.get_median_count() == 0
    kc.insert(1, 10); kc.insert(2, 20)
    assert kc.get_median_count() == 15
    kc.insert(3, 30)
    assert kc.get_median_count() == 20
    kc2 = oracle.KmerCounts(5)
    kc2.insert(1, 5); kc2.insert(2, 100); kc2.insert(3, 50)
    assert kc2.get_max_count() == 100


def test_remove_low_and_filtered_view(oracle):  # counting.rs:458-480
    kc = oracle.KmerCounts(5)
    kc.insert(1, 1); kc.insert(2, 5); kc.insert(3, 10)
    kc.remove_low_count_kmers(5)
    assert not kc.contains(1) and kc.contains(2) and kc.contains(3)
    kc = oracle.KmerCounts(5)
    kc.insert(1, 2); kc.insert(2, 10)
    fv = kc.filtered_view(5)
    assert fv.get_canonical(1) is None and fv.get_canonical(2) == 10
    assert fv.get_canonical_count(1) == 0 and fv.get_canonical_count(2) == 10


def test_ingest_seq(oracle):  # counting.rs:482-490
    kc = oracle.KmerCounts(3)
    kc.ingest_seq("ACGT")
    assert kc.get_n_unique_kmers() == 1 and kc.get_n_kmers() == 2


def test_extend_with_histogram(oracle):  # counting.rs:492-509
    a, b = oracle.KmerCounts(5), oracle.KmerCounts(5)
    a.insert(1, 3); b.insert(1, 2); b.insert(2, 5)
    h = oracle.Histogram(100)
    h.move_count(0, 3)
    a.extend_with_histogram(b, h)
    assert a.get_count(1) == 5 and a.get_count(2) == 5
    v = h.get_vector()
    assert v[5] == 2 and v.sum() == 2


def test_extend_with_histogram_saturation(oracle):  # counting.rs:183-189 (stored count is the new bin)
    a, b = oracle.KmerCounts(5), oracle.KmerCounts(5)
    a.insert(1, U32_MAX - 1); b.insert(1, 5)
    h = oracle.Histogram(10)
    h.move_count(0, U32_MAX - 1)
    assert a.extend_with_histogram(b, h) is True
    assert a.get_count(1) == U32_MAX
    assert h.get_vector()[11] == 1 and h.get_n_kmers() == U32_MAX


# ---- src/pcr/mod.rs:1236-1342: 18S x 10 replicates ---------------------------

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")


def test_18s_integration_table(oracle):
    with open(os.path.join(GOLDEN_DIR, "pcr_18s_k21.json")) as f:
        gold = json.load(f)
    s = open(os.path.join(GOLDEN_DIR, "pcr_18s_read.txt")).read().strip()
    ref = "/root/reference/src/pcr/mod.rs"
    if os.path.exists(ref):  # the fixture is the reference's test string, verbatim
        assert re.search(r'let read_string = "([ACGT]+)"', open(ref).read()).group(1) == s
    kc = oracle.KmerCounts(21)
    for _ in range(10):
        kc.ingest_seq(s)
    # pcr/mod.rs:1336-1342
    assert kc.len() == len(s) - 21 + 1 == gold["n_distinct"] == 1812
    assert kc.get_n_kmers() == (len(s) - 21 + 1) * 10 == gold["n_kmers"] == 18120
    keys, counts = kc.export_sorted()
    assert (counts == 10).all()
    assert len(s) == gold["n_bases"]
    assert kc.digest() == gold["digest"]
    # independent brute force
    want = oracle.py_count([s] * 10, 21)
    assert dict(zip(keys.tolist(), counts.tolist())) == dict(want)


def test_synth_goldens(oracle):
    """The committed synthetic-case goldens are what the oracle produces today."""
    cases = json.load(open(os.path.join(GOLDEN_DIR, "synth_cases.json")))
    for name, g in cases.items():
        reads = oracle.synth_reads(g["seed"], g["genome_len"], g["read_len"], g["sub_rate"], g["n_rate"], 0, g["n_reads"])
        run = oracle.Run(g["k"], g["chunks"], g["histo_max"])
        run.push_lines(reads)
        run.finish()
        t = run.table()
        assert (t.len(), t.digest(), run.n_kmers_ingested) == (g["n_unique"], g["digest"], g["n_kmers"]), name
        for c in range(g["chunks"]):
            assert run.histogram(c).tolist() == g["histograms"][c], (name, c)


# ---- brute force cross-check ---------------------------------------------------

@pytest.mark.parametrize("k", [1, 3, 15, 21, 31])
def test_run_matches_bruteforce(oracle, k):
    rng = random.Random(100 + k)
    seqs = []
    for i in range(2500):
        L = rng.choice([0, 1, k - 1, k, k + 1, 40, 75, 150]) if i % 7 == 0 else rng.randint(20, 60)
        s = "".join(rng.choice("ACGT" if rng.random() > 0.02 else "N") for _ in range(max(L, 0)))
        seqs.append(s)
    # a small "genome" so counts exceed 1
    genome = "".join(rng.choice("ACGT") for _ in range(400))
    for i in range(1500):
        a = rng.randint(0, 340)
        seqs.append(genome[a:a + 60])
    want = oracle.py_count(seqs, k)
    for chunks in (0, 1, 3):
        run = oracle.Run(k, chunks, histo_max=5)
        for s in seqs:
            run.push_seq(s)
        run.finish()
        keys, counts = run.table().export_sorted()
        assert dict(zip(keys.tolist(), counts.tolist())) == dict(want)
        assert run.n_reads_read == len(seqs)
        assert run.n_bases_read == sum(map(len, seqs))
        assert run.n_bases_ingested == sum(len(s) - s.count("N") for s in seqs)
        assert run.n_kmers_ingested == sum(want.values())
        if chunks:
            assert run.histogram(chunks - 1).tolist() == oracle.py_histogram(want, 5)
            # column i = histogram of the reads of chunks 0..i (batches of 1000, round-robin)
            for ci in range(chunks):
                sub = [s for j, s in enumerate(seqs) if (j // 1000) % chunks <= ci]
                assert run.histogram(ci).tolist() == oracle.py_histogram(oracle.py_count(sub, k), 5)


def test_final_histogram_independent_of_chunks(oracle):  # tests/spcr_18s.rs:437-528 (property)
    reads = oracle.synth_reads(seed=11, genome_len=20000, read_len=100, sub_rate=0.01, n_rate=0.002,
                               first=0, n=12345)
    finals = []
    for chunks in (1, 20):
        run = oracle.Run(21, chunks, histo_max=1000)
        run.push_lines(reads)
        run.finish()
        finals.append(run.histogram(chunks - 1))
    assert (finals[0] == finals[1]).all()
