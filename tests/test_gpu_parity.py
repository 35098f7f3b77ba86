"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle (and the committed goldens) on the same seeded inputs.  Bit-exact: this is
integer work, so every comparison is equality."""
import json
import os
import random
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMPTY = 0xFFFFFFFFFFFFFFFF
U32_MAX = 0xFFFFFFFF


@pytest.fixture(scope="module")
def skm():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from sharkmer_b200 import kmer
    return kmer


def lines(seqs):
    return ("\n".join(seqs) + "\n").encode() if seqs else b""


# ---- (1) pack kernel: the reference's packed-layout KATs (src/kmer/mod.rs:61-111) ----------

def test_pack_layout_kats(skm):
    e = skm.Engine(9)
    codes, breaks = e.pack(b"CGTAATGCGGCGA\n")
    top = int(codes[0])
    assert [(top >> s) & 0xFF for s in (56, 48, 40)] == [0b01101100, 0b00111001, 0b10100110]
    assert (top >> 38) & 3 == 0  # 13th base 'A'
    assert int(breaks[0]) == (0xFFFFFFFF >> 13)  # 13 bases, then newline + padding are breaks
    codes, breaks = e.pack(b"C\n")
    assert int(codes[0]) >> 56 == 0b01000000 and int(breaks[0]) == 0x7FFFFFFF
    # N is a break; bases after it keep their positions
    codes, breaks = e.pack(b"CGTANATGCGGCGA\n")
    assert int(breaks[0]) >> 17 == 0b000010000000001  # first 15 positions: N at 4, newline at 14
    # long input: every 32-base unit, unaligned tail
    rng = random.Random(1)
    s = "".join(rng.choice("ACGT") for _ in range(1000))
    codes, breaks = e.pack(lines([s]))
    for u in range(len(codes)):
        want = 0
        for j in range(32):
            p = u * 32 + j
            want = (want << 2) | ("ACGT".index(s[p]) if p < len(s) else 0)
        got = int(codes[u])
        nb = min(32, max(0, len(s) - u * 32))
        assert got == want, u  # breaks and padding pack as 00
        valid_mask = ((1 << nb) - 1) << (32 - nb)
        assert int(breaks[u]) == (~valid_mask) & 0xFFFFFFFF, u


# ---- (2) extract kernel: order-exact against kmers_from_ascii -----------------------------------

CASES = ["CGTAATGCGGCGA", "CGTANATGCGGCGA", "NCGTANATGCGGCGA", "NCGTANATGCGGCGANN",
         "NNCGTANATGCGGCGA", "TANCACN", "NTANCACNAGAAAATC", "AAAA", "ACGTACGTACGT", "", "N", "A"]


def check_extract(skm, oracle, seqs, k):
    e = skm.Engine(k)
    buf = lines(seqs)
    out = e.extract_kmers(buf)
    pos = 0
    for s in seqs:
        got = [int(v) for v in out[pos:pos + len(s) + 1] if int(v) != EMPTY]
        assert got == oracle.kmers_from_ascii(s, k), (s[:60], k)
        assert int(out[pos + len(s)]) == EMPTY  # nothing ends on a separator
        pos += len(s) + 1


def test_extract_reference_kats(skm, oracle):  # src/kmer/mod.rs:179-278
    e = skm.Engine(9)
    out = e.extract_kmers(b"CGTAATGCGGCG\n")
    got = [int(v) for v in out if int(v) != EMPTY]
    assert got == [0b01_1001_0011_1100_0110, 0b01_0110_0100_1111_0001, 0b10_0101_1001_0011_1100,
                   0b00_0011_1001_1010_0110]
    for k in (3, 5, 9, 11):
        check_extract(skm, oracle, CASES, k)
    assert skm.kmers_from_ascii("ACGT", 9) == []
    assert len(skm.kmers_from_ascii("ACGTACGTA", 9)) == 1


@pytest.mark.parametrize("k", [1, 3, 15, 21, 25, 31])
def test_extract_random_reads(skm, oracle, k):
    rng = random.Random(k)
    seqs = []
    for i in range(3000):
        L = rng.choice([0, 1, 2, k - 1, k, k + 1, 31, 32, 33, 63, 64, 65, 100, 150, 151, 250])
        s = "".join(rng.choice("ACGT") if rng.random() > 0.03 else "N" for _ in range(max(0, L)))
        seqs.append(s)
    check_extract(skm, oracle, seqs, k)


def test_extract_unaligned_tail_sizes(skm, oracle):
    rng = random.Random(5)
    for n in list(range(1, 70)) + [127, 128, 129, 255, 256, 257, 8191, 8192, 8193]:
        s = "".join(rng.choice("ACGT") for _ in range(n - 1))
        check_extract(skm, oracle, [s], 5)


# ---- (3) counting: sorted table, histograms, totals ---------------------------------------------

def run_oracle(oracle, reads, k, chunks, hmax):
    run = oracle.Run(k, chunks, hmax)
    run.push_lines(reads)
    run.finish()
    return run


def run_gpu(skm, reads, k, chunks, hmax, read_len, mode=0, capacity_hint=0, batches_per_call=3):
    """Feeds the engine the way src/io.rs does: 1000-read batches, round-robin over chunks."""
    e = skm.Engine(k, chunks, hmax, capacity_hint=capacity_hint, insert_mode=mode)
    n_chunks = max(1, chunks)
    line = read_len + 1
    n_reads = len(reads) // line
    per_chunk = [[] for _ in range(n_chunks)]
    for b in range((n_reads + 999) // 1000):
        per_chunk[b % n_chunks].append(reads[b * 1000 * line:min((b + 1) * 1000, n_reads) * line])
    # several batches of one chunk may share a call, in order
    for c in range(n_chunks):
        bl = per_chunk[c]
        for i in range(0, len(bl), batches_per_call):
            e.ingest_batch(c, np.concatenate(bl[i:i + batches_per_call]))
    e.finalize()
    return e


def compare(e, run, chunks):
    keys, counts = e.export(sorted=True)
    okeys, ocounts = run.table().export_sorted()
    assert keys.size == okeys.size
    assert (keys == okeys).all() and (counts == ocounts).all()
    assert e.digest() == run.table().digest()
    t = e.totals()
    assert (t.n_reads, t.n_bases, t.n_bases_read, t.n_kmers, t.n_unique) == (
        run.n_reads_ingested, run.n_bases_ingested, run.n_bases_read, run.n_kmers_ingested, okeys.size)
    for c in range(run.n_chunks):
        ct = e.chunk_totals(c)
        assert (ct.n_reads, ct.n_bases, ct.n_kmers) == run.chunk_totals(c), c
    for c in range(chunks):
        assert (e.histogram(c) == run.histogram(c)).all(), c
    if chunks:
        assert t.n_singletons == run.n_singletons()


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("k,chunks", [(21, 10), (31, 0), (31, 1), (25, 3), (1, 2), (15, 4)])
def test_count_parity_vs_oracle(skm, oracle, k, chunks, mode):
    L = 150
    reads = oracle.synth_reads(seed=100 + k, genome_len=60_000, read_len=L, sub_rate=0.01, n_rate=0.001,
                               first=0, n=23_456)
    run = run_oracle(oracle, reads, k, chunks, 200)
    e = run_gpu(skm, reads, k, chunks, 200, L, mode=mode)
    compare(e, run, chunks)


@pytest.mark.parametrize("tile_log2,max_buckets,k,chunks,g2", [(14, 16, 21, 3, None), (15, 16, 31, 0, None), (16, 16, 21, 10, None),
                                                                (16, 1024, 25, 2, None), (15, 4, 15, 4, None),
                                                                (13, 64, 21, 3, 11), (16, 16, 31, 2, 11), (14, 8, 25, 0, 9)])
def test_cluster_tile_sort_vs_oracle(skm, oracle, monkeypatch, tile_log2, max_buckets, k, chunks, g2):
    """Pass B by a thread-block cluster (tile_sort_cluster_kernel): 2 / 4 / 8 CTAs sort 2^14 / 2^15 / 2^16 k-mers as one
    tile, counts exchanged through distributed shared memory.  Few buckets, so that a bucket spans several full
    tiles and every CTA of a cluster has cells; 1024 buckets: every tile is a partial one."""
    monkeypatch.setenv("SKM_TILE_LOG2", str(tile_log2))
    monkeypatch.setenv("SKM_MAX_BUCKETS", str(max_buckets))
    if g2 is not None:   # forced sub-bucket bits (2^11 = the most a tile is sorted by; also the one-CTA sort at 2^13)
        monkeypatch.setenv("SKM_G2", str(g2))
    L = 150
    reads = oracle.synth_reads(seed=300 + tile_log2, genome_len=80_000, read_len=L, sub_rate=0.01, n_rate=0.001,
                               first=0, n=25_000)
    run = run_oracle(oracle, reads, k, chunks, 200)
    for hint in (0, 600_000):   # without / with a capacity hint (sub-bucket bits follow the table's partitions)
        e = run_gpu(skm, reads, k, chunks, 200, L, mode=2, capacity_hint=hint)
        compare(e, run, chunks)
        assert e.stage_times().tiled_launches >= 1
        e.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_count_parity_goldens(skm, oracle, mode):
    """Committed goldens (tests/golden/synth_cases.json, made by make_golden.py)."""
    cases = json.load(open(os.path.join(GOLDEN_DIR, "synth_cases.json")))
    for name, g in cases.items():
        reads = oracle.synth_reads(g["seed"], g["genome_len"], g["read_len"], g["sub_rate"], g["n_rate"], 0, g["n_reads"])
        e = run_gpu(skm, reads, g["k"], g["chunks"], g["histo_max"], g["read_len"], mode=mode)
        t = e.totals()
        assert (t.n_unique, t.n_kmers, t.n_bases) == (g["n_unique"], g["n_kmers"], g["n_bases_ingested"]), name
        assert e.digest() == g["digest"], name
        for c in range(g["chunks"]):
            assert e.histogram(c).tolist() == g["histograms"][c], (name, c)
        for c, want in enumerate(g["chunk_totals"]):
            ct = e.chunk_totals(c)
            assert [ct.n_reads, ct.n_bases, ct.n_kmers] == want, (name, c)


def test_18s_golden(skm):  # src/pcr/mod.rs:1236-1247,1336-1342
    gold = json.load(open(os.path.join(GOLDEN_DIR, "pcr_18s_k21.json")))
    s = open(os.path.join(GOLDEN_DIR, "pcr_18s_read.txt")).read().strip()
    kc = skm.KmerCounts(21)
    for _ in range(10):
        kc.ingest_seq(s)
    assert kc.len() == len(s) - 21 + 1 == gold["n_distinct"]
    assert kc.get_n_kmers() == gold["n_kmers"]
    keys, counts = kc.export_sorted()
    assert (counts == 10).all()
    assert kc.digest() == gold["digest"]
    # the same through the batch path
    e = skm.Engine(21, chunks=1)
    e.ingest_batch(0, lines([s] * 10))
    e.finalize()
    assert e.digest() == gold["digest"] and e.histogram(0)[10] == 1812


@pytest.mark.parametrize("poly_frac", [0.04, 0.10, 0.7, 1.0])
def test_skewed_buckets_overflow_and_fallback(skm, oracle, poly_frac):
    """Single-GPU bucketing is one pass into fixed-capacity regions (CapLayout): a k-mer repeated
    many times overfills its region.  4 % poly-A reads spill into the overflow run, 10 % nearly
    fill it, 70 % / 100 % overflow that too and the engine re-buckets the batch exactly."""
    L = 150
    n = 30_000
    reads = oracle.synth_reads(seed=11, genome_len=200_000, read_len=L, sub_rate=0.01, n_rate=0.001, first=0, n=n).copy()
    lines = reads.reshape(n, L + 1)
    rng = np.random.default_rng(5)
    poly = rng.random(n) < poly_frac
    lines[poly, :L] = ord("A")
    reads = lines.reshape(-1)
    for chunks in (0, 3):
        run = run_oracle(oracle, reads, 21, chunks, 300)
        e = run_gpu(skm, reads, 21, chunks, 300, L, mode=2)
        compare(e, run, chunks)


@pytest.mark.parametrize("skew", [False, True])
def test_speculative_insert_launch_and_its_guard(skm, oracle, skew):
    """Batches large enough for the one-pass capped layout on ONE GPU (>= 4096 k-mers per bucket), all on the device by
    the time skm_finalize runs: the insert is launched speculatively, behind the lists' events, with the overflow check
    on the device (spec_guard_kernel).  skew: a tenth of the reads are poly-A, a capped bucket overflows, the
    speculative launch must do NOTHING and the batch is rebuilt with the exact layout.  Either way: the oracle's table."""
    L, chunks = 100, 2
    per_batch = 4_600_000 // (L + 1)
    n = per_batch * chunks
    reads = oracle.synth_reads(61, 1_000_000, L, 0.005, 0.001, 0, n).copy()
    if skew:
        reads.reshape(n, L + 1)[::10, :L] = ord("A")
    run = run_oracle(oracle, reads, 21, chunks, 500)
    e = skm.Engine(21, chunks, 500, insert_mode=2)
    line = L + 1
    by_chunk = [[] for _ in range(chunks)]
    for b in range((n + 999) // 1000):
        by_chunk[b % chunks].append(reads[b * 1000 * line:min((b + 1) * 1000, n) * line])
    for c in range(chunks):
        e.ingest_batch(c, np.concatenate(by_chunk[c]))   # one large batch per chunk
    e.finalize()
    compare(e, run, chunks)
    assert e.stage_times().tiled_launches >= 1
    e.close()


def test_table_growth_from_tiny(skm, oracle):
    """capacity_hint = 0: the table starts at 2^16 slots and must grow several times."""
    L = 100
    reads = oracle.synth_reads(seed=77, genome_len=3_000_000, read_len=L, sub_rate=0.02, n_rate=0.0, first=0, n=40_000)
    for mode in (1, 2):
        e = run_gpu(skm, reads, 31, 2, 100, L, mode=mode, capacity_hint=0)
        run = run_oracle(oracle, reads, 31, 2, 100)
        assert e.stage_times().n_grows >= 1
        compare(e, run, 2)


def test_ingest_reads_offsets_form(skm, oracle):
    rng = random.Random(3)
    seqs = ["".join(rng.choice("ACGTN" if rng.random() < 0.05 else "ACGT") for _ in range(rng.choice([0, 5, 31, 32, 90, 150])))
            for _ in range(5000)]
    bases = "".join(seqs).encode()
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(s) for s in seqs])
    e = skm.Engine(21, chunks=1, histo_max=50)
    e.ingest_reads(0, bases, offs)
    e.finalize()
    run = oracle.Run(21, 1, 50)
    for s in seqs:
        run.push_seq(s)
    run.finish()
    compare(e, run, 1)


def test_invalid_base_is_an_error(skm):  # src/kmer/encoding.rs:353-356
    for bad, ch in ((b"ACGTACGTaCGTACGT\n", "a"), (b"ACGT\nACGRT\n", "R"), (b"AC GT\n", " ")):
        e = skm.Engine(3, chunks=1)
        e.ingest_batch(0, bad)
        with pytest.raises(skm.SkmError) as err:
            e.finalize()
        assert err.value.code == 2
        assert f"Invalid character '{ch}' in sequence. Only ACGTN allowed." in str(err.value)
    with pytest.raises(skm.SkmError):
        skm.KmerCounts(3).ingest_seq("ACGU")


def test_no_reads_is_an_error(skm):  # src/io.rs:578-580
    e = skm.Engine(21, chunks=1)
    with pytest.raises(skm.SkmError) as err:
        e.finalize()
    assert err.value.code == 8 and "No reads were ingested" in str(err.value)


def test_empty_reads_are_counted(skm, oracle):
    e = skm.Engine(5, chunks=1)
    e.ingest_batch(0, b"\n\nACGTACG\n\nNNNNNNN\nAC\n")
    e.finalize()
    t = e.totals()
    assert (t.n_reads, t.n_bases, t.n_bases_read, t.n_kmers) == (6, 9, 16, 3)


# ---- (4) KmerCounts API behaviour (src/kmer/counting.rs:365-510) ----------------------------------

def test_counts_api_kats(skm):
    kc = skm.KmerCounts(5)
    assert kc.get_k() == 5 and kc.is_empty() and kc.len() == 0 and kc.get_n_kmers() == 0
    kc.insert(42, 3)
    assert kc.get_count(42) == 3 and kc.contains(42) and not kc.contains(99)
    kc.insert(42, 7)
    assert kc.get_count(42) == 10 and kc.len() == 1
    s = skm.KmerCounts(5)
    s.insert(1, U32_MAX)
    s.insert(1, 1)
    assert s.get_count(1) == U32_MAX  # saturating
    a, b = skm.KmerCounts(5), skm.KmerCounts(5)
    a.insert(1, 10); a.insert(2, 20); b.insert(2, 5); b.insert(3, 15)
    a.extend(b)
    assert (a.get_count(1), a.get_count(2), a.get_count(3)) == (10, 25, 15)
    with pytest.raises(skm.SkmError):
        a.extend(skm.KmerCounts(7))
    m = skm.KmerCounts(5)
    assert m.get_median_count() == 0
    m.insert(1, 10); m.insert(2, 20)
    assert m.get_median_count() == 15
    m.insert(3, 30)
    assert m.get_median_count() == 20 and m.get_max_count() == 30
    f = skm.KmerCounts(5)
    f.insert(1, 2); f.insert(2, 10)
    fv = f.filtered_view(5)
    assert fv.get_canonical(1) is None and fv.get_canonical(2) == 10
    assert fv.get_canonical_count(1) == 0 and fv.get_canonical_count(2) == 10
    g = skm.KmerCounts(3)
    g.ingest_seq("ACGT")
    assert g.get_n_unique_kmers() == 1 and g.get_n_kmers() == 2


def test_histogram_kat(skm):  # src/kmer/mod.rs:288-305
    kc = skm.KmerCounts(11)
    for kmer, c in ((1, 5), (20, 5), (2, 7), (11, 11), (12, 12)):
        kc.insert(kmer, c)
    v = skm.Histogram.from_kmer_counts(kc, 10).get_vector()
    assert v.tolist() == [0, 0, 0, 0, 0, 2, 0, 1, 0, 0, 0, 2]


def test_saturation_in_histogram(skm):
    e = skm.Engine(5, chunks=1, histo_max=10)
    e.insert_counts([1, 2, 3], [U32_MAX, U32_MAX - 1, 7])
    e.insert_counts([1, 2], [5, 5])
    e.snapshot_histogram(0)
    h = e.histogram(0)
    assert h[7] == 1 and h[11] == 2
    t = e.totals()
    assert t.n_saturated == 2 and t.n_kmers == 2 * U32_MAX + 7
    c, f = e.lookup([1, 2, 3, 4], 0, 1)
    assert c.tolist() == [U32_MAX, U32_MAX, 7, 0] and f.tolist() == [True, True, True, False]


def test_lookup_batch_vs_oracle(skm, oracle):
    L = 120
    reads = oracle.synth_reads(seed=5, genome_len=30_000, read_len=L, sub_rate=0.01, n_rate=0.0, first=0, n=8_000)
    run = run_oracle(oracle, reads, 25, 0, 100)
    e = run_gpu(skm, reads, 25, 0, 100, L)
    keys, _ = run.table().export_sorted()
    rng = np.random.default_rng(0)
    q = np.concatenate([keys[::7], rng.integers(0, 1 << 50, 5000, dtype=np.uint64)])
    # query in a random orientation
    flip = rng.random(q.size) < 0.5
    qq = np.array([skm.revcomp_kmer(int(x), 25) if f else int(x) for x, f in zip(q, flip)], dtype=np.uint64)
    for min_count in (0, 2, 5):
        got, found = e.lookup(qq, min_count)
        fv = run.table().filtered_view(min_count)
        want = np.array([fv.get_canonical_count(int(x)) for x in qq], dtype=np.uint32)
        assert (got == want).all()
        wantf = np.array([fv.get_canonical(int(x)) is not None for x in qq])
        _, found2 = e.lookup(qq, min_count, 2)
        assert (found2 == wantf).all()


def test_scan_oligos_vs_oracle(skm, oracle):  # src/pcr/primers.rs:163-226
    L, k = 120, 25
    reads = oracle.synth_reads(seed=8, genome_len=30_000, read_len=L, sub_rate=0.01, n_rate=0.0, first=0, n=8_000)
    run = run_oracle(oracle, reads, k, 0, 100)
    e = run_gpu(skm, reads, k, 0, 100, L)
    keys, _ = run.table().export_sorted()
    rng = np.random.default_rng(1)
    for olen, min_count in ((12, 2), (8, 1), (20, 5), (24, 0)):
        # oligos taken from real k-mers (prefixes and reverse-complemented suffixes) plus random ones
        pick = keys[rng.integers(0, keys.size, 40)]
        oligos = [int(x) >> (2 * (k - olen)) for x in pick[:20]]
        oligos += [skm.revcomp_kmer(int(x) & ((1 << (2 * olen)) - 1), olen) for x in pick[20:]]
        oligos += [int(x) for x in rng.integers(0, 1 << (2 * olen), 20, dtype=np.uint64)]
        gk, gc = e.scan_oligos(oligos, olen, min_count)
        ok, oc = run.table().find_oligos(oligos, olen, min_count)
        assert gk.size == ok.size and gk.size > 0
        assert (gk == ok).all() and (gc == oc).all()
    with pytest.raises(skm.SkmError):
        e.scan_oligos([1], k, 0)  # oligo length must be < k


# ---- (5) synthetic generator: device == host, bit for bit ------------------------------------------

def test_get_primer_kmers_vs_oracle(skm, oracle):  # src/pcr/primers.rs:376-478, pcr/mod.rs:1344-1350
    """sPCR's primer k-mer discovery (one device scan per primer direction and mismatch level)
    against the oracle restatement: the 18S case of the reference's own test, then a synthetic
    genome with planted primer sites."""
    from oracle import primers_oracle as po
    from sharkmer_b200 import primers as pp
    read = open(os.path.join(os.path.dirname(__file__), "golden", "pcr_18s_read.txt")).read().strip()
    e = skm.Engine(21, chunks=0)
    t = oracle.KmerCounts(21)
    for _ in range(10):
        e.ingest_batch(0, np.frombuffer((read + "\n").encode(), dtype=np.uint8))
        t.ingest_seq(read)
    e.finalize()
    args = dict(forward_seq="AACCTGGTTGATCCTGCCAGT", reverse_seq="TGATCCTTCTGCAGGTTCACCTAC", min_count=3, mismatches=2,
                trim=15, max_primer_kmers=40)
    (fk, fc), (rk, rc) = pp.get_primer_kmers(pp.PCRParams(**args), e, 21)
    want_f, want_r = po.get_primer_kmers(po.PCRParams(**args), t)
    assert dict(zip(fk.tolist(), fc.tolist())) == want_f and dict(zip(rk.tolist(), rc.tolist())) == want_r
    assert fk.size == 1 and rk.size == 1 and fc[0] == 10 and rc[0] == 10

    rng = random.Random(9)
    L = 150
    fwd, rev = "ACGGTCATTGCAGGTCAAGT", "TTGACCGTAGGCATCCAGTA"
    def rcs(s):
        return s[::-1].translate(str.maketrans("ACGT", "TGCA"))
    lines = []
    for i in range(3000):
        s = [rng.choice("ACGT") for _ in range(L)]
        if i % 3 == 0:   # plant a (possibly mutated, possibly reverse-complemented) primer site
            site = list(fwd if i % 2 else rev)
            for p in rng.sample(range(len(site)), rng.randint(0, 3)):
                site[p] = rng.choice("ACGT")
            site = "".join(site)
            if rng.random() < 0.5:
                site = rcs(site)
            at = rng.randint(0, L - len(site))
            s[at:at + len(site)] = site
        lines.append("".join(s))
    lines = lines + lines[:1500]   # some k-mers twice
    buf = np.frombuffer(("\n".join(lines) + "\n").encode(), dtype=np.uint8)
    for k, cap, mm in ((21, 40, 2), (31, 3, 1), (15, 5, 2)):
        e = skm.Engine(k, chunks=0)
        e.ingest_batch(0, buf)
        e.finalize()
        t = oracle.KmerCounts(k)
        for s in lines:
            t.ingest_seq(s)
        args = dict(forward_seq=fwd[:-2] + "R" + fwd[-1], reverse_seq=rev, min_count=1, mismatches=mm, trim=15,
                    max_primer_kmers=cap)
        (fk, fc), (rk, rc) = pp.get_primer_kmers(pp.PCRParams(**args), e, k)
        want_f, want_r = po.get_primer_kmers(po.PCRParams(**args), t)
        assert dict(zip(fk.tolist(), fc.tolist())) == want_f, k
        assert dict(zip(rk.tolist(), rc.tolist())) == want_r, k
        assert len(want_f) > 0 and len(want_r) > 0


def test_device_synth_matches_host(skm, oracle):
    e = skm.Engine(21)
    L, n, n_chunks = 150, 2345, 3
    for c in range(n_chunks):
        d = e.device_alloc(n * (L + 1))
        e.synth_device(9, 1_000_000, L, oracle.rate_to_thresh(0.01), oracle.rate_to_thresh(0.001), c, n_chunks, 0, n, d)
        got = np.empty(n * (L + 1), dtype=np.uint8)
        e.memcpy_d2h(got, d, got.size)
        e.device_free(d)
        # chunk-local read i is global read ((i//1000)*n_chunks + c)*1000 + i%1000
        want = np.empty_like(got)
        for b in range((n + 999) // 1000):
            cnt = min(1000, n - b * 1000)
            first = (b * n_chunks + c) * 1000
            want[b * 1000 * (L + 1):(b * 1000 + cnt) * (L + 1)] = oracle.synth_reads(9, 1_000_000, L, 0.01, 0.001, first, cnt)
        assert (got == want).all(), c


# ---- (6) properties at larger size --------------------------------------------------------------------

def test_chunk_invariance_and_digest_300k_reads(skm, oracle):
    """tests/spcr_18s.rs:437-528 (final histogram independent of --chunks), on
    device-generated reads, plus the oracle's digest on the same input."""
    L, n = 150, 300_000
    G, seed = 2_000_000, 21
    st, nt = oracle.rate_to_thresh(0.01), oracle.rate_to_thresh(0.001)
    finals, digests = [], []
    for chunks, mode in ((1, 1), (20, 1), (7, 2)):
        e = skm.Engine(21, chunks, 1000, capacity_hint=30_000_000, insert_mode=mode)
        per = [0] * chunks
        for b in range(n // 1000):
            per[b % chunks] += 1000
        for c in range(chunks):
            d = e.device_alloc(per[c] * (L + 1))
            e.synth_device(seed, G, L, st, nt, c, chunks, 0, per[c], d)
            e.ingest_device(c, d, per[c] * (L + 1))
            e.sync()
            e.device_free(d)
        e.finalize()
        finals.append(e.histogram(chunks - 1))
        digests.append(e.digest())
        t = e.totals()
        assert t.n_reads == n
    assert (finals[0] == finals[1]).all() and (finals[0] == finals[2]).all()
    assert digests[0] == digests[1] == digests[2]
    run = oracle.Run(21, 1, 1000)
    run.push_lines(oracle.synth_reads(seed, G, L, 0.01, 0.001, 0, n))
    run.finish()
    assert run.table().digest() == digests[0]
    assert (run.histogram(0) == finals[0]).all()


def test_reset_reuses_ctx(skm, oracle):
    L = 100
    e = skm.Engine(21, 2, 100)
    for seed in (1, 2):
        reads = oracle.synth_reads(seed, 50_000, L, 0.01, 0.0, 0, 4000)
        e.reset()
        for b in range(4):
            e.ingest_batch(b % 2, reads[b * 1000 * (L + 1):(b + 1) * 1000 * (L + 1)])
        e.finalize()
        run = run_oracle(oracle, reads, 21, 2, 100)
        compare(e, run, 2)


# ---- (7) hash-sharded path: several ranks on one GPU ---------------------------------------------------

def _feed_ranks(engines, reads, n_reads, line, chunks, world):
    """1000-read batches round-robin over chunks (src/io.rs:355-361); the batches of a chunk are
    dealt to the ranks in turn."""
    n_chunks = max(1, chunks)
    for b in range((n_reads + 999) // 1000):
        c = b % n_chunks
        r = (b // n_chunks) % world
        engines[r].ingest_batch(c, reads[b * 1000 * line:min((b + 1) * 1000, n_reads) * line])


def _check_sharded(group_like, engines, run, chunks, world):
    from sharkmer_b200 import common
    all_keys, all_counts = [], []
    for r, e in enumerate(engines):
        keys, cnts = e.export(sorted=True)
        assert all(common.owner_rank(common.hash_kmer(int(x)), world) == r for x in keys[:3000])
        all_keys.append(keys)
        all_counts.append(cnts)
    okeys, ocounts = run.table().export_sorted()
    all_keys, all_counts = np.concatenate(all_keys), np.concatenate(all_counts)
    order = np.argsort(all_keys, kind="stable")
    # the oracle's keys are sorted and distinct: equality also says the partitions are disjoint
    assert all_keys.size == okeys.size
    assert np.array_equal(all_keys[order], okeys)
    assert np.array_equal(all_counts[order].astype(np.uint64), ocounts.astype(np.uint64))
    for c in range(chunks):   # every rank holds the columns summed over all ranks
        for r, e in enumerate(engines):
            got, want = e.histogram(c), run.histogram(c)
            bad = np.nonzero(got != want)[0]
            assert bad.size == 0, (c, r, bad[:8].tolist(), got[bad[:8]].tolist(), want[bad[:8]].tolist())
    assert sum(int(e.totals().n_kmers) for e in engines) == run.n_kmers_ingested
    assert sum(int(e.totals().n_unique) for e in engines) == okeys.size
    assert sum(int(e.totals().n_reads) for e in engines) == run.n_reads_ingested


@pytest.mark.parametrize("slices", ["1", "0"])
@pytest.mark.parametrize("world,k,chunks,mode", [(2, 25, 3, 0), (3, 21, 5, 0), (2, 31, 0, 0), (4, 21, 2, 2), (8, 21, 2, 0)])
def test_sharded_group_vs_oracle(skm, oracle, monkeypatch, world, k, chunks, mode, slices):
    """skm_group_*: `world` ranks in one process on ONE GPU — bucketing by (owner, region), tile sort,
    copy-engine exchange into the peers' arenas, collective finalize, column sum.  Sorted table
    (partition by partition, each key on its owner) and every histogram column vs the oracle.
    Both exchange layouts: slices of one cluster-sorted list (what a capacity_hint selects), and one
    re-bucketed list per owner (no hint, or tables too fine for the slices)."""
    from sharkmer_b200.multigpu import Group
    monkeypatch.setenv("SKM_MG_SLICES", slices)
    L, n, hmax = 120, 23_000, 100
    reads = oracle.synth_reads(31 + world, 50_000, L, 0.01, 0.001, 0, n)
    run = run_oracle(oracle, reads, k, chunks, hmax)
    g = Group(k, chunks, hmax, [0] * world, arena_bytes_per_rank=96 << 20, insert_mode=mode)
    _feed_ranks(g.engines, reads, n, L + 1, chunks, world)
    g.finalize()
    _check_sharded(g, g.engines, run, chunks, world)
    assert all(e.mg_bytes_sent() > 0 for e in g.engines)
    # sharded table services answer like one table
    keys, counts = run.table().export_sorted()
    probe = np.concatenate([keys[::97][:400], np.array([12345, 99999999], dtype=np.uint64)])
    got, found = g.lookup(probe, 0, 1)
    want = dict(zip(keys.tolist(), counts.tolist()))
    assert got.tolist() == [want.get(int(x), 0) for x in probe]
    # a second sample through the same group (reset is collective)
    g.reset()
    reads2 = oracle.synth_reads(77, 30_000, L, 0.01, 0.0, 0, 8000)
    run2 = run_oracle(oracle, reads2, k, chunks, hmax)
    _feed_ranks(g.engines, reads2, 8000, L + 1, chunks, world)
    g.finalize()
    _check_sharded(g, g.engines, run2, chunks, world)
    g.close()


@pytest.mark.parametrize("world,chunks,g2", [(2, 3, None), (4, 0, None), (3, 2, None), (4, 2, 11)])
def test_sharded_group_few_buckets_full_cluster_tiles(skm, oracle, monkeypatch, world, chunks, g2):
    """Sender lists with few, large buckets (SKM_MAX_BUCKETS=16): an owner's coarse bucket spans several full
    cluster tiles, so the slices shipped hold whole 2^16-cell tiles sorted by all eight CTAs."""
    from sharkmer_b200.multigpu import Group
    monkeypatch.setenv("SKM_MAX_BUCKETS", "16")
    monkeypatch.setenv("SKM_MG_SLICES", "1")
    if g2 is not None:
        monkeypatch.setenv("SKM_G2", str(g2))
    L, n, hmax, k = 150, 24_000, 100, 21
    reads = oracle.synth_reads(171 + world, 60_000, L, 0.01, 0.001, 0, n)
    run = run_oracle(oracle, reads, k, chunks, hmax)
    g = Group(k, chunks, hmax, [0] * world, arena_bytes_per_rank=128 << 20, insert_mode=0)
    # a rank ingests all of a chunk's share at once: one list per (rank, chunk) with ~100 k k-mers per bucket
    line, n_chunks = L + 1, max(1, chunks)
    per = [[[] for _ in range(n_chunks)] for _ in range(world)]
    for b in range((n + 999) // 1000):
        per[(b // n_chunks) % world][b % n_chunks].append(reads[b * 1000 * line:min((b + 1) * 1000, n) * line])
    for r in range(world):
        for c in range(n_chunks):
            if per[r][c]:
                g.engines[r].ingest_batch(c, np.concatenate(per[r][c]))
    g.finalize()
    _check_sharded(g, g.engines, run, chunks, world)
    g.close()


def test_sharded_group_with_capacity_hint(skm, oracle):
    """No override: with a capacity_hint the ctx picks the exchange layout itself (slices of one cluster-sorted list
    when the hinted table's partitions are coarse enough for them), and the sub-bucket bits follow the hinted table."""
    from sharkmer_b200.multigpu import Group
    L, n, hmax, k = 120, 23_000, 100, 21
    reads = oracle.synth_reads(201, 50_000, L, 0.01, 0.001, 0, n)
    for world, chunks, hint in ((2, 3, 400_000), (4, 0, 150_000), (3, 2, 40_000_000)):
        run = run_oracle(oracle, reads, k, chunks, hmax)
        g = Group(k, chunks, hmax, [0] * world, arena_bytes_per_rank=96 << 20, capacity_hint=hint, insert_mode=2)
        _feed_ranks(g.engines, reads, n, L + 1, chunks, world)
        g.finalize()
        _check_sharded(g, g.engines, run, chunks, world)
        g.close()


@pytest.mark.parametrize("slices", ["1", "0"])
def test_sharded_group_skewed_and_unbalanced(skm, oracle, monkeypatch, slices):
    """Overflowing capped buckets on a sender (poly-A reads: one k-mer millions of times) take the
    exact re-bucketing path and are shipped again; one rank gets no reads at all."""
    from sharkmer_b200.multigpu import Group
    monkeypatch.setenv("SKM_MG_SLICES", slices)
    L, n, k, chunks, hmax, world = 100, 16_000, 21, 2, 1000, 3
    reads = oracle.synth_reads(5, 40_000, L, 0.01, 0.001, 0, n).copy()
    lines = reads.reshape(n, L + 1)
    lines[::3, :L] = ord("A")          # a third of the reads are poly-A
    run = run_oracle(oracle, reads, k, chunks, hmax)
    g = Group(k, chunks, hmax, [0] * world, arena_bytes_per_rank=96 << 20, insert_mode=2)
    # rank 2 ingests nothing
    n_chunks = max(1, chunks)
    for b in range(n // 1000):
        g.engines[(b // n_chunks) % 2].ingest_batch(b % n_chunks, reads[b * 1000 * (L + 1):(b + 1) * 1000 * (L + 1)])
    g.finalize()
    _check_sharded(g, g.engines, run, chunks, world)
    g.close()


@pytest.mark.parametrize("slices", ["1", "0"])
@pytest.mark.parametrize("world,k,chunks,skew", [(2, 21, 2, False), (4, 31, 1, True)])
def test_sharded_group_capped_layouts(skm, oracle, monkeypatch, world, k, chunks, skew, slices):
    """Batches large enough for the one-pass capped layouts on the senders (coarse list by (owner, region); then
    either slices of it, cluster-sorted, or every owner's list re-bucketed into its 1024 regions by ONE launch for
    all owners).  skew: a tenth of the reads are poly-A, so a capped bucket overflows, the batch is rebuilt exactly
    and shipped again."""
    from sharkmer_b200.multigpu import Group
    monkeypatch.setenv("SKM_MG_SLICES", slices)
    L, hmax = 100, 1000
    per_batch = (4_200_000 * world) // (L + 1)          # reads per (rank, chunk) batch: past the capped threshold
    n = per_batch * world * chunks
    reads = oracle.synth_reads(90 + world, 2_000_000, L, 0.002, 0.001, 0, n).copy()
    if skew:
        reads.reshape(n, L + 1)[::10, :L] = ord("A")
    run = run_oracle(oracle, reads, k, chunks, hmax)
    line = L + 1
    g = Group(k, chunks, hmax, [0] * world, arena_bytes_per_rank=per_batch * line * 24 * chunks + (64 << 20), insert_mode=0)
    # the reference deals 1000-read batches round-robin over the chunks (src/io.rs:355-361); here every rank
    # ingests its share of a chunk's batches as ONE large batch
    by_chunk = [[] for _ in range(chunks)]
    for b in range((n + 999) // 1000):
        by_chunk[b % chunks].append(reads[b * 1000 * line:min((b + 1) * 1000, n) * line])
    for c in range(chunks):
        whole = np.concatenate(by_chunk[c])
        n_c = whole.size // line
        cuts = [n_c * r // world for r in range(world + 1)]
        for r in range(world):
            g.engines[r].ingest_batch(c, whole[cuts[r] * line:cuts[r + 1] * line])
    g.finalize()
    _check_sharded(g, g.engines, run, chunks, world)
    st = [e.stage_times() for e in g.engines]
    assert all(t.tiled_launches >= 1 for t in st)
    g.close()


def _mp_worker(rank, world, port, q, k, chunks, hmax, L, n, seed):
    import os
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    from oracle import oracle as o
    from sharkmer_b200 import kmer
    from sharkmer_b200.multigpu import ShardedCounter
    torch.cuda.set_device(0)
    reads = o.synth_reads(seed, 50_000, L, 0.01, 0.001, 0, n)
    e = kmer.Engine(k, chunks, hmax, device=0, n_ranks=world, rank=rank)
    sc = ShardedCounter(e, torch.device("cuda", 0), arena_bytes=64 << 20)
    line = L + 1
    for b in range(n // 1000):
        if (b // chunks) % world == rank:
            e.ingest_batch(b % chunks, reads[b * 1000 * line:(b + 1) * 1000 * line])
    cols = sc.finalize()
    keys, cnts = e.export(sorted=True)
    t = e.totals()
    q.put((rank, keys, cnts, cols, int(t.n_kmers)))
    dist.barrier()
    e.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("slices", ["1", "0"])
def test_sharded_two_processes_one_gpu(oracle, monkeypatch, slices):
    """The torchrun shape of the multi-GPU path on a one-GPU box: two PROCESSES (two ranks) share
    cuda:0, arenas mapped through CUDA IPC, the library's all-gather carried by gloo
    (sharkmer_b200.multigpu.ShardedCounter, what bench.py uses at --gpus N).  No NCCL and no
    kernel that waits for another rank: peer copies + host-side collectives only."""
    import socket
    import torch.multiprocessing as mp
    monkeypatch.setenv("SKM_MG_SLICES", slices)   # (inherited by the spawned ranks)
    world, k, chunks, hmax, L, n, seed = 2, 21, 4, 100, 100, 16_000, 44
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_mp_worker, args=(r, world, port, q, k, chunks, hmax, L, n, seed)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    run = run_oracle(oracle, oracle.synth_reads(seed, 50_000, L, 0.01, 0.001, 0, n), k, chunks, hmax)
    merged = {}
    for _, keys, cnts, cols, _ in got:
        assert not (set(keys[:2000].tolist()) & set(merged))
        merged.update(zip(keys.tolist(), cnts.tolist()))
        for c in range(chunks):
            assert (cols[c] == run.histogram(c)).all(), c
    okeys, ocounts = run.table().export_sorted()
    assert merged == dict(zip(okeys.tolist(), ocounts.tolist()))
    assert sum(t[4] for t in got) == run.n_kmers_ingested


# ---- (8) BASELINE.json configurations at full size ----------------------------------------------------

def _golden_full(name):
    path = os.path.join(GOLDEN_DIR, "full_cases.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/full_cases.json missing (python tests/golden/make_golden_full.py)")
    g = json.load(open(path))
    if name not in g:
        pytest.skip(f"{name} not in full_cases.json")
    return g[name]


def _run_config_on_device(skm, oracle, g, mode=0, hint=0):
    """Device-generated reads of a full_cases.json configuration, fed like src/io.rs does."""
    k, chunks, hmax, L, n = g["k"], g["chunks"], g["histo_max"], g["read_len"], g["n_reads"]
    n_chunks = max(1, chunks)
    st, nt = oracle.rate_to_thresh(g["sub_rate"]), oracle.rate_to_thresh(g["n_rate"])
    e = skm.Engine(k, chunks, hmax, capacity_hint=hint, insert_mode=mode)
    n_batches = (n + 999) // 1000
    for c in range(n_chunks):
        per = len(range(c, n_batches, n_chunks)) * 1000     # (n is a multiple of 1000 * n_chunks here)
        d = e.device_alloc(per * (L + 1))
        e.synth_device(g["seed"], g["genome_len"], L, st, nt, c, n_chunks, 0, per, d)
        e.ingest_device(c, d, per * (L + 1))
        e.sync()
        e.device_free(d)
    e.finalize()
    return e


def _check_against_golden(e, g):
    t = e.totals()
    assert (t.n_reads, t.n_bases, t.n_kmers, t.n_unique) == (g["n_reads_ingested"], g["n_bases_ingested"], g["n_kmers"], g["n_unique"])
    assert e.digest() == g["digest"]
    for c in range(g["chunks"]):
        want = np.zeros(g["histo_max"] + 2, dtype=np.uint64)
        for b, v in g["histograms"][c].items():
            want[int(b)] = v
        assert (e.histogram(c) == want).all(), c


@pytest.mark.parametrize("name", ["C1_chunks0", "C1_chunks1"])
def test_c1_full_size_vs_oracle_and_golden(skm, oracle, name):
    """BASELINE config 1 (sharkmer -k 31 --max-reads 1000000, 1 M synthetic 150 bp reads), chunks 0 and 1:
    the full sorted (k-mer, count) table and the histogram against the oracle run here, and against the
    committed golden (digest, totals, histogram) the oracle produced in the build container."""
    g = _golden_full(name)
    e = _run_config_on_device(skm, oracle, g, hint=45_000_000)
    _check_against_golden(e, g)
    run = oracle.Run(g["k"], g["chunks"], g["histo_max"])
    for first in range(0, g["n_reads"], 250_000):
        run.push_lines(oracle.synth_reads(g["seed"], g["genome_len"], g["read_len"], g["sub_rate"], g["n_rate"], first, 250_000))
    run.finish()
    keys, counts = e.export(sorted=True)
    okeys, ocounts = run.table().export_sorted()
    assert keys.size == okeys.size == g["n_unique"]
    assert (keys == okeys).all() and (counts == ocounts).all()
    if g["chunks"]:
        assert (e.histogram(0) == run.histogram(0)).all()
    e.close()


def test_c2_full_size_vs_golden(skm, oracle):
    """BASELINE config 2 — the bench workload (k=21, 10 chunks, 10 M reads) at full size through the tiled
    insert, against the oracle's committed result: table digest, totals and all ten histogram columns."""
    g = _golden_full("C2")
    e = _run_config_on_device(skm, oracle, g, hint=320_000_000)
    assert e.stage_times().tiled_launches >= 1
    _check_against_golden(e, g)
    e.close()


def test_memory_bounded_rounds(skm, oracle, monkeypatch):
    """Inputs whose k-mer lists do not fit beside the table (BASELINE config 5 in miniature): with a small
    memory budget the batches stay packed, and finalize builds and counts their lists a few at a time —
    in the middle of a chunk if need be.  Same table, same histogram columns."""
    L, n, k, chunks, hmax = 150, 400_000, 25, 2, 200
    reads = oracle.synth_reads(61, 3_000_000, L, 0.01, 0.001, 0, n)
    run = run_oracle(oracle, reads, k, chunks, hmax)
    # budget = table (2^26 slots = 1 GiB) + 4 GiB slack + ~1.5 lists of 100 k reads (15 M cells, 130 MB each)
    monkeypatch.setenv("SKM_MEM_BUDGET", str(1024 + 4096 + 200))
    e = skm.Engine(k, chunks, hmax, capacity_hint=38_000_000, insert_mode=2)
    line = L + 1
    for b in range(0, n // 1000, 100):        # 100 batches (100 k reads) per call, alternating chunks
        c = (b // 100) % chunks
        e.ingest_batch(c, reads[b * 1000 * line:(b + 100) * 1000 * line])
    e.finalize()
    assert e.stage_times().tiled_launches >= 3
    keys, counts = e.export(sorted=True)
    okeys, ocounts = run.table().export_sorted()
    # (the batches were dealt to the chunks 100 at a time, not round-robin: compare what does not depend on that)
    assert (keys == okeys).all() and (counts == ocounts).all()
    assert (e.histogram(chunks - 1) == run.histogram(chunks - 1)).all()
    e.close()


@pytest.mark.parametrize("slices", ["1", "0"])
def test_sharded_group_flush_rounds(skm, oracle, monkeypatch, slices):
    """skm_group_flush: a sharded count in several rounds (chunks == 0), lists and arenas freed in between."""
    from sharkmer_b200.multigpu import Group
    monkeypatch.setenv("SKM_MG_SLICES", slices)
    L, n, k, world = 100, 24_000, 31, 3
    reads = oracle.synth_reads(8, 60_000, L, 0.01, 0.001, 0, n)
    run = run_oracle(oracle, reads, k, 0, 100)
    g = Group(k, 0, 100, [0] * world, arena_bytes_per_rank=64 << 20, insert_mode=2)
    line = L + 1
    for b in range(n // 1000):
        g.engines[b % world].ingest_batch(0, reads[b * 1000 * line:(b + 1) * 1000 * line])
        if b % 8 == 7:
            g.flush()
    g.finalize()
    _check_sharded(g, g.engines, run, 0, world)
    g.close()
