"""Primer preprocessing + primer k-mer discovery (the sPCR stage that scans the count table):
the oracle restatement (oracle/primers_oracle.py) against the reference's own unit-test vectors
(src/pcr/primers.rs:484-833, src/pcr/mod.rs:1236-1311), and the product code
(sharkmer_b200/primers.py, integer based) against the oracle.  CPU only: the device scan is stood
in for by the oracle's find_oligos here; tests/test_gpu_parity.py runs the same discovery through
skm_scan_oligos."""
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import primers_oracle as po
from sharkmer_b200 import primers as pp

HERE = os.path.dirname(os.path.abspath(__file__))
READ_18S = open(os.path.join(HERE, "golden", "pcr_18s_read.txt")).read().strip()


def params_18s(cls):  # src/pcr/mod.rs:1250-1281
    return cls(forward_seq="AACCTGGTTGATCCTGCCAGT", reverse_seq="TGATCCTTCTGCAGGTTCACCTAC", gene_name="18s",
               min_count=3, mismatches=2, trim=15, max_primer_kmers=40)


class OracleScan:
    """Engine stand-in: scan_oligos answered by the oracle table (tests only)."""
    def __init__(self, table):
        self.table = table
        self.calls = 0
    def scan_oligos(self, oligos, length, min_count):
        self.calls += 1
        return self.table.find_oligos(oligos, length, min_count)


def table_from(oracle, seq, k, replicates=1):
    t = oracle.KmerCounts(k)
    for _ in range(replicates):
        t.ingest_seq(seq)
    return t


# ---- reference unit tests, restated (primers.rs:484-600) ----------------------------------------

def test_string_to_oligo_kats():
    for mod in (po, pp):
        assert mod.string_to_oligo("A") == (1, 0b00) and mod.string_to_oligo("T") == (1, 0b11)
        assert mod.string_to_oligo("ACGT") == (4, 0b00011011)
        assert mod.string_to_oligo("") == (0, 0)
        for bad in ("ACNGT", "X", "a"):
            with pytest.raises(Exception) as e:
                mod.string_to_oligo(bad)
            assert "Invalid nucleotide" in str(e.value)
        with pytest.raises(Exception) as e:
            mod.string_to_oligo("A" * 33)
        assert "exceeds maximum of 32 bases" in str(e.value)


def unpack(arr, length):
    return {pp.oligo_to_string(v, length) for v in arr}


def test_resolve_primer_kats():
    assert po.resolve_primer("ACGT") == {"ACGT"}
    assert po.resolve_primer("AR") == {"AA", "AG"}
    assert po.resolve_primer("RY") == {"AC", "AT", "GC", "GT"}
    assert po.resolve_primer("N") == {"A", "C", "G", "T"}
    for s in ("ACGT", "AR", "RY", "N", "BDHVKMSWACGT", ""):
        assert unpack(pp.resolve_primer(s), len(s)) == po.resolve_primer(s)
    assert all(po.is_valid_nucleotide(c) for c in "ACGTRYWSKMBDHVN")
    assert not po.is_valid_nucleotide("X") and not po.is_valid_nucleotide("a")


def test_combinations_and_permute_kats():
    assert len(po.combinations(4, 2)) == 6 and len(po.combinations(5, 0)) == 1 and len(po.combinations(3, 3)) == 1
    assert po.combinations(2, 5) == []
    assert po.permute_sequences({"ACG"}, 0) == {"ACG"}
    r = po.permute_sequences({"AC"}, 1)
    assert len(r) == 7 and {"AC", "TC", "AG"} <= r
    assert unpack(pp.hamming1(np.array([pp.string_to_oligo("AC")[1]], dtype=np.uint64), 2), 2) == r


def test_preprocess_levels_kat():  # primers.rs:759-821
    for cls, mod in ((po.PCRParams, po), (pp.PCRParams, pp)):
        prm = cls(forward_seq="ACGTACGT", reverse_seq="TGCATGCA", min_count=2, mismatches=2, trim=7)
        if mod is po:
            levels = po.preprocess_primer_by_mismatch(prm, False, 8)
        else:
            length, lv = pp.preprocess_primer_by_mismatch(prm, False, 8)
            assert length == 7
            levels = [unpack(l, length) for l in lv]
        assert len(levels) == 3 and "CGTACGT" in levels[0]
        for i in range(3):
            for j in range(i + 1, 3):
                assert not (levels[i] & levels[j])
        assert set().union(*levels) == po.preprocess_primer(po.PCRParams("ACGTACGT", "TGCATGCA", mismatches=2, trim=7), False, 8)
        assert [len(l) for l in levels] == [1, 21, 189]  # 7*3 and C(7,2)*9


def test_18s_reverse_primer_variants_kat():  # pcr/mod.rs:1285-1302
    flat = po.preprocess_primer(params_18s(po.PCRParams), True, 21)
    assert len(flat) == 991 and "TGCAGGTTCACCTAC" in flat and "GGCAGGTTCACCTAC" in flat
    length, lv = pp.preprocess_primer_by_mismatch(params_18s(pp.PCRParams), True, 21)
    assert length == 15 and [l.size for l in lv] == [1, 45, 945]
    assert set().union(*[unpack(l, 15) for l in lv]) == flat


def test_trim_clamp_and_short_primers():
    prm = po.PCRParams("ACGTACGTAC", "TTGCA", trim=30, mismatches=1)
    assert po.trimmed_primer(prm, False, 9) == "GTACGTAC"       # trim >= k -> k-1 (primers.rs:245-254)
    assert po.trimmed_primer(prm, True, 9) == "TTGCA"           # shorter than trim: kept whole
    q = pp.PCRParams("ACGTACGTAC", "TTGCA", trim=30, mismatches=1)
    assert pp.trimmed_primer(q, False, 9) == "GTACGTAC" and pp.trimmed_primer(q, True, 9) == "TTGCA"
    # mismatches are clamped to the primer length (primers.rs:279-280)
    lv = po.preprocess_primer_by_mismatch(po.PCRParams("AC", "AC", mismatches=5, trim=15), False, 21)
    assert len(lv) == 3 and [len(l) for l in lv] == [1, 6, 9]
    length, lv2 = pp.preprocess_primer_by_mismatch(pp.PCRParams("AC", "AC", mismatches=5, trim=15), False, 21)
    assert [l.size for l in lv2] == [1, 6, 9]


def test_too_many_variants_is_an_error():
    for mod in (po, pp):
        prm = mod.PCRParams("N" * 7 + "ACGT", "ACGT")
        with pytest.raises(Exception) as e:
            mod.preprocess_primer_by_mismatch(prm, False, 21)
        assert "has too many ambiguous bases: 16384 resolved variants exceeds limit of 10000" in str(e.value)


# ---- find_oligos / discovery on tables (primers.rs:603-700, pcr/mod.rs:1303-1350) ---------------

def test_find_oligos_kats(oracle):
    def find(seq, oligo, min_count, k=5):
        t = table_from(oracle, seq, k)
        n, v = po.string_to_oligo(oligo)
        return t.find_oligos(np.array([v], dtype=np.uint64), n, min_count)
    assert find("ACGTACGT", "ACG", 1)[0].size >= 1
    assert find("AAAAAAAAAA", "GGG", 1)[0].size == 0
    assert find("AACCCAACC", "AAC", 2)[0].size == 0
    keys, _ = find("TTTTTTT", "AAA", 1)
    assert keys.size == 1 and pp.oligo_to_string(keys[0], 5) == "AAAAA"   # stored in the oligo's orientation
    assert find("ACGTACGT", "ACGT", 1)[0].size >= 1                       # oligo length k-1


def test_filter_primer_kmers_kats(oracle):
    seqs = ["AAAAA", "AAAAC", "AAACG", "AACGT", "ACGTA"]
    km = [po.string_to_oligo(s)[1] for s in seqs]
    assert po.filter_primer_kmers({}, 10) == {}
    assert len(po.filter_primer_kmers({k: i + 1 for i, k in enumerate(km)}, 3)) == 3
    capped = po.filter_primer_kmers({k: 2 for k in km}, 3)
    assert sorted(capped) == sorted(km)[:3]                                 # ties broken by k-mer value
    assert len(po.filter_primer_kmers({km[0]: 3, km[1]: 3}, 5)) == 2


def test_18s_primer_kmers_kat(oracle):  # pcr/mod.rs:1303-1311, 1344-1350
    t = table_from(oracle, READ_18S, 21, replicates=10)
    prm = params_18s(po.PCRParams)
    rev_variants = po.preprocess_primer(prm, True, 21)
    assert len(po.get_kmers_from_primers(rev_variants, t, prm.min_count)) == 1
    fwd, rev = po.get_primer_kmers(prm, t)
    assert len(fwd) == 1 and len(rev) == 1
    (fk, fc), (rk, rc) = pp.get_primer_kmers(params_18s(pp.PCRParams), OracleScan(t), 21)
    assert dict(zip(fk.tolist(), fc.tolist())) == fwd and dict(zip(rk.tolist(), rc.tolist())) == rev
    assert list(fc) == [10] and list(rc) == [10]
    # the forward primer k-mer starts with the trimmed primer; the reverse primer binds with one
    # mismatch (site CGCAGGTTCACCTAC..., the reverse complement of the read's 3' end, mod.rs:1253-1255)
    assert pp.oligo_to_string(fk[0], 21).startswith("GTTGATCCTGCCAGT")
    assert pp.oligo_to_string(rk[0], 21) == "CGCAGGTTCACCTACGGAAAC"


@pytest.mark.parametrize("seed", range(6))
def test_discovery_matches_oracle_on_random_tables(oracle, seed):
    """Random genome with planted near-matches of both primers, random caps / mismatch budgets /
    ambiguity codes: the integer implementation must give exactly the oracle's primer k-mers."""
    rng = random.Random(seed)
    k = rng.choice([15, 21, 25, 31])
    genome = "".join(rng.choice("ACGT") for _ in range(6000))
    fwd = "".join(rng.choice("ACGT") for _ in range(rng.randint(8, 24)))
    rev = "".join(rng.choice("ACGT") for _ in range(rng.randint(8, 24)))
    def mutate(s, n):
        s = list(s)
        for p in rng.sample(range(len(s)), min(n, len(s))):
            s[p] = rng.choice("ACGT")
        return "".join(s)
    def rc(s):
        return s[::-1].translate(str.maketrans("ACGT", "TGCA"))
    t = oracle.KmerCounts(k)
    for i in range(60):   # plant variants (0-3 substitutions) with different copy numbers, both strands
        site = mutate(fwd if i % 2 == 0 else rev, rng.randint(0, 3)) + "".join(rng.choice("ACGT") for _ in range(40))
        site = site if rng.random() < 0.5 else rc(site)
        for _ in range(rng.randint(1, 6)):
            t.ingest_seq("".join(rng.choice("ACGT") for _ in range(5)) + site)
    t.ingest_seq(genome)
    if seed % 2:   # ambiguity codes in the primers
        fwd = fwd[:-3] + "R" + fwd[-2:]
        rev = rev[:-5] + "N" + rev[-4:-1] + "Y"
    args = dict(forward_seq=fwd, reverse_seq=rev, min_count=rng.randint(1, 3), mismatches=rng.randint(0, 2),
                trim=rng.choice([10, 15, 40]), max_primer_kmers=rng.choice([1, 3, 40]))
    want_f, want_r = po.get_primer_kmers(po.PCRParams(**args), t)
    eng = OracleScan(t)
    (fk, fc), (rk, rc_) = pp.get_primer_kmers(pp.PCRParams(**args), eng, k)
    assert dict(zip(fk.tolist(), fc.tolist())) == want_f
    assert dict(zip(rk.tolist(), rc_.tolist())) == want_r
    assert list(fk) == sorted(fk) and list(rk) == sorted(rk)


# ---- the C++ twin (sharkmer_b200/host/primers.hpp) through the test-only mock ABI ----------------

@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("primers") / "host_harness"
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-o", str(out), os.path.join(HERE, "host", "mock_abi.cpp"),
                    "-lz"], check=True)
    return str(out)


def run_cpp(harness, tmp_path, table, k, args):
    keys, counts = table.export_sorted()
    tf = tmp_path / "table.txt"
    with open(tf, "w") as f:
        for a, b in zip(keys.tolist(), counts.tolist()):
            f.write(f"{a} {b}\n")
    cmd = [harness, "-k", str(k), "--table", str(tf), "--forward", args["forward_seq"], "--reverse", args["reverse_seq"],
           "--mismatches", str(args["mismatches"]), "--trim", str(args["trim"]), "--min-count", str(args["min_count"]),
           "--cap", str(args["max_primer_kmers"])]
    r = subprocess.run(cmd, capture_output=True, text=True)
    fwd, rev = {}, {}
    for line in r.stdout.split("\n"):
        if line:
            d, km, ct = line.split()
            (fwd if d == "F" else rev)[int(km)] = int(ct)
    return r, fwd, rev


def test_cpp_18s_primer_kmers(harness, oracle, tmp_path):
    t = table_from(oracle, READ_18S, 21, replicates=10)
    args = dict(forward_seq="AACCTGGTTGATCCTGCCAGT", reverse_seq="TGATCCTTCTGCAGGTTCACCTAC", min_count=3, mismatches=2,
                trim=15, max_primer_kmers=40)
    r, fwd, rev = run_cpp(harness, tmp_path, t, 21, args)
    assert r.returncode == 0, r.stderr
    want_f, want_r = po.get_primer_kmers(po.PCRParams(**args), t)
    assert (fwd, rev) == (want_f, want_r) and len(fwd) == 1 and len(rev) == 1


@pytest.mark.parametrize("seed", range(4))
def test_cpp_discovery_matches_oracle(harness, oracle, tmp_path, seed):
    rng = random.Random(100 + seed)
    k = rng.choice([15, 21, 31])
    fwd = "".join(rng.choice("ACGT") for _ in range(rng.randint(8, 22)))
    rev = "".join(rng.choice("ACGT") for _ in range(rng.randint(8, 22)))
    t = oracle.KmerCounts(k)
    for i in range(50):
        s = list(fwd if i % 2 else rev)
        for p in rng.sample(range(len(s)), rng.randint(0, 3)):
            s[p] = rng.choice("ACGT")
        site = "".join(s) + "".join(rng.choice("ACGT") for _ in range(40))
        if rng.random() < 0.5:
            site = site[::-1].translate(str.maketrans("ACGT", "TGCA"))
        for _ in range(rng.randint(1, 5)):
            t.ingest_seq("".join(rng.choice("ACGT") for _ in range(4)) + site)
    t.ingest_seq("".join(rng.choice("ACGT") for _ in range(3000)))
    if seed % 2:
        fwd = fwd[:-4] + "W" + fwd[-3:]
        rev = "B" + rev[1:]
    args = dict(forward_seq=fwd, reverse_seq=rev, min_count=rng.randint(1, 3), mismatches=rng.randint(0, 2),
                trim=rng.choice([9, 15, 40]), max_primer_kmers=rng.choice([2, 5, 40]))
    r, got_f, got_r = run_cpp(harness, tmp_path, t, k, args)
    assert r.returncode == 0, r.stderr
    want_f, want_r = po.get_primer_kmers(po.PCRParams(**args), t)
    assert (got_f, got_r) == (want_f, want_r)


def test_cpp_primer_errors(harness, oracle, tmp_path):
    t = table_from(oracle, "ACGTACGTACGTTTGACCA", 5)
    base = dict(reverse_seq="ACG", min_count=1, mismatches=1, trim=15, max_primer_kmers=40)
    r, _, _ = run_cpp(harness, tmp_path, t, 5, dict(forward_seq="NNNNNNNACG", **base) | {"trim": 15})
    assert r.returncode == 0  # trimmed to k-1 = 4 bases: NACG -> 4 variants, fine
    t21 = table_from(oracle, READ_18S[:200], 21)
    r, _, _ = run_cpp(harness, tmp_path, t21, 21, dict(forward_seq="N" * 8 + "ACGTACG", **base))
    assert r.returncode == 1 and "has too many ambiguous bases: 65536 resolved variants exceeds limit of 10000" in r.stderr
    r, _, _ = run_cpp(harness, tmp_path, t21, 21, dict(forward_seq="ACGTXACGT", **base))
    assert r.returncode == 1 and "Invalid nucleotide X" in r.stderr
