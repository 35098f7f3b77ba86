"""sPCR over the count table (sharkmer_b200/pcr.py), CPU side: the reference's unit-test vectors for
every stage (src/pcr/graph.rs:653-765, pruning.rs:237-342, paths.rs:488-620, mod.rs:1236-1395), the
18S integration case, and end-to-end recovery of planted amplicons from synthetic reads.  The table
is the oracle's (tests only); tests/test_gpu_zz_pcr.py runs the same pipeline on the device table."""
import os
import random

import numpy as np
import pytest

from sharkmer_b200 import _lib, pcr
from sharkmer_b200.primers import PCRParams

HERE = os.path.dirname(os.path.abspath(__file__))
READ_18S = open(os.path.join(HERE, "golden", "pcr_18s_read.txt")).read().strip()


def revcomp_np(x, k):
    """Vectorised reverse complement of 2-bit packed k-mers (same bit tricks as skm_revcomp_kmer)."""
    x = ~np.asarray(x, dtype=np.uint64)
    for sh, m in ((2, 0x3333333333333333), (4, 0x0F0F0F0F0F0F0F0F), (8, 0x00FF00FF00FF00FF), (16, 0x0000FFFF0000FFFF)):
        m = np.uint64(m)
        x = ((x >> np.uint64(sh)) & m) | ((x & m) << np.uint64(sh))
    x = (x >> np.uint64(32)) | (x << np.uint64(32))
    return x >> np.uint64(64 - 2 * k)


class OracleTable:
    """Engine stand-in over an oracle KmerCounts (tests only): lookup (EITHER mode) + scan_oligos.
    The oracle's tables hold canonical k-mers only, so "the k-mer as given, else its reverse
    complement" is a search for min(kmer, revcomp) in the sorted export."""
    def __init__(self, table):
        self.t = table
        self.k = table.get_k()
        self.keys, self.counts = table.export_sorted()
        self.lookup_calls = 0
        self.lookups = 0
    def scan_oligos(self, oligos, length, min_count):
        return self.t.find_oligos(oligos, length, min_count)
    def lookup(self, kmers, min_count, mode):
        assert mode == _lib.LOOKUP_EITHER
        self.lookup_calls += 1
        self.lookups += len(kmers)
        q = np.asarray(kmers, dtype=np.uint64)
        canon = np.minimum(q, revcomp_np(q, self.k))
        pos = np.minimum(np.searchsorted(self.keys, canon), max(self.keys.size - 1, 0))
        hit = (self.keys[pos] == canon) if self.keys.size else np.zeros(q.size, dtype=bool)
        counts = np.where(hit, self.counts[pos] if self.keys.size else 0, 0).astype(np.uint32)
        found = hit & (counts >= min_count)
        return np.where(found, counts, 0).astype(np.uint32), found


def rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def mk_graph(nodes, edges):
    g = pcr.DiGraph()
    for sub, s, e in nodes:
        g.add_node(sub, s, e)
    for a, b, c in edges:
        g.add_edge(a, b, c, 1.0)
    return g


# ---- graph.rs unit tests -----------------------------------------------------------------------

def test_node_budget_suffix_mask_medians():
    assert pcr.compute_node_budget(0) == 100_000 and pcr.compute_node_budget(150_000_000) == 100_000
    assert pcr.compute_node_budget(750_000_000) == 500_000 and pcr.compute_node_budget(2**64 - 1) == 500_000
    assert pcr.compute_node_budget(450_000_000) == 300_000
    assert pcr.get_suffix_mask(3) == 0b1111 and pcr.get_suffix_mask(2) == 0b11
    assert pcr.median_f64([]) is None and pcr.median_f64([42]) == 42.0
    assert pcr.median_f64([9, 1, 5]) == 5.0 and pcr.median_f64([11, 1, 9, 5]) == 7.0 and pcr.median_f64([7, 3]) == 5.0
    m = pcr.median_f64([2**32 - 2, 2**32 - 1])
    assert abs(m - ((2**32 - 2) + (2**32 - 1)) / 2.0) < 1.0
    assert pcr.compute_median([3, 1, 4, 1, 5, 9, 2, 6]) == 3.5 and pcr.compute_median([]) == 0.0
    g = mk_graph([(0, True, False), (1, False, False), (2, False, True)], [(0, 1, 5), (1, 2, 15), (0, 2, 10)])
    assert pcr.median_f64(g.edge_counts()) == 10.0
    assert pcr.primer_counts_max_median([10, 20]) == (20, 15) and pcr.primer_counts_max_median([]) == (0, 0)
    assert pcr.primer_counts_max_median([5, 100, 50]) == (100, 50)


def test_coverage_thresholds():  # pcr/mod.rs:405-432
    assert pcr.compute_coverage_thresholds(10, 3) == [5, 3]             # high 5, step 0: [5,5,5,3] deduplicated
    assert pcr.compute_coverage_thresholds(4, 3) == [3]                 # high 2 <= min
    assert pcr.compute_coverage_thresholds(100, 2) == [50, 34, 18, 2]
    assert pcr.compute_coverage_thresholds(16, 2) == [8, 6, 4, 2]
    assert pcr.compute_coverage_thresholds(12, 2) == [6, 5, 4, 2]
    assert pcr.compute_coverage_thresholds(9, 2) == [4, 2]              # step 0: [4,4,4,2] deduplicated


# ---- pruning.rs unit tests ---------------------------------------------------------------------

def test_pruning_kats():
    g = mk_graph([(0, True, False), (0, False, False), (0, False, False), (0, False, True), (0, False, False)],
                 [(0, 1, 100), (1, 2, 100), (2, 3, 100), (2, 4, 1)])
    pcr.remove_low_coverage_tips(g, 3, 0.1)
    assert g.n_nodes == 4 and g.nodes[4] is None
    g = mk_graph([(0, True, False), (0, False, False), (0, False, True), (0, False, False)],
                 [(0, 1, 10), (1, 2, 10), (1, 3, 10)])
    pcr.remove_low_coverage_tips(g, 3, 0.1)
    assert g.n_nodes == 4
    g = mk_graph([(0, True, False), (0, False, False), (0, False, True), (0, False, False)], [(0, 1, 10), (1, 2, 10)])
    pcr.reachability_pruning(g)
    assert g.n_nodes == 3 and g.nodes[0] is not None and g.nodes[3] is None
    g = mk_graph([(0, True, False), (0, False, False), (0, False, True), (0, False, False)],
                 [(0, 1, 10), (1, 2, 10), (0, 3, 10)])
    pcr.reachability_pruning(g)
    assert g.n_nodes == 3 and g.nodes[3] is None
    g = pcr.DiGraph()
    pcr.reachability_pruning(g)
    assert g.n_nodes == 0
    g = mk_graph([(0, True, False), (0, False, False), (0, False, True)], [(0, 1, 10), (1, 2, 20), (0, 2, 30)])
    assert pcr.median_f64(g.edge_counts()) == 20.0


# ---- paths.rs unit tests -----------------------------------------------------------------------

def P(min_length, max_length, **kw):
    return PCRParams("ACGT", "TGCA", gene_name="test", min_count=2, mismatches=0, trim=0, min_length=min_length,
                     max_length=max_length, **kw)


def test_path_search_kats():
    g = mk_graph([(0, True, False), (1, False, False), (2, False, False), (3, False, True)], [(0, 1, 10), (1, 2, 10), (2, 3, 10)])
    paths = pcr.get_assembly_paths(g, 3, P(0, 100))
    assert len(paths) == 1 and [n for n, _ in paths[0]] == [0, 1, 2, 3]
    assert paths[0][0][1] is None and all(e is not None for _, e in paths[0][1:])
    g = mk_graph([(0, True, False), (1, False, False), (2, False, False), (3, False, True)],
                 [(0, 1, 10), (0, 2, 5), (1, 3, 10), (2, 3, 5)])
    paths = pcr.get_assembly_paths(g, 3, P(0, 100))
    assert len(paths) == 2 and [n for n, _ in paths[0]] == [0, 1, 3]     # highest coverage first
    assert pcr.get_assembly_paths(pcr.DiGraph(), 3, P(0, 100)) == []
    g = mk_graph([(0, True, False), (1, False, False), (2, False, False), (3, False, False), (4, False, True)],
                 [(0, 1, 10), (1, 2, 10), (2, 3, 10), (3, 4, 10)])
    assert pcr.get_assembly_paths(g, 3, P(0, 5)) == []                    # needs 5 nodes, cap is 4
    assert len(pcr.get_assembly_paths(g, 3, P(0, 6))) == 1
    g = mk_graph([(0, True, False), (1, False, False), (2, False, True)], [(0, 1, 10), (1, 2, 10)])
    assert pcr.get_assembly_paths(g, 3, P(0, 100, max_dfs_states=0)) == []
    g = mk_graph([(0, True, False), (1, False, False), (2, False, False)], [(0, 1, 1), (0, 2, 100)])
    ch = pcr.sorted_children(g, 0)
    assert [(c[0], c[1]) for c in ch] == [(1, 0), (2, 1)]                 # ascending: pop gives the high one
    # ties: petgraph walks a node's edges newest first and the sort is stable, so of two equal-count
    # children the OLDER edge ends up last and is explored first
    g = mk_graph([(0, True, False), (1, False, False), (2, False, False)], [(0, 1, 7), (0, 2, 7)])
    assert [c[0] for c in pcr.sorted_children(g, 0)] == [2, 1]


def test_sequences_scores_and_dedup():
    # start sub_kmer AC, then C, G, T appended: ACCGT; k = 3
    sub = lambda s: sum("ACGT".index(c) << (2 * (len(s) - 1 - i)) for i, c in enumerate(s))
    g = mk_graph([(sub("AC"), True, False), (sub("CC"), False, False), (sub("CG"), False, False), (sub("GT"), False, True)],
                 [(0, 1, 10), (1, 2, 20), (2, 3, 30)])
    pcr.annotate_coverage_ratios(g)
    assert [round(g.edges[e][3], 6) for e in g.edge_indices()] == [0.5, 1.0, 1.5]
    paths = pcr.get_assembly_paths(g, 3, P(0, 100))
    recs, nxt = pcr.generate_sequences_from_paths(g, paths, 3, "smp", P(0, 100))
    assert nxt == 1 and recs[0].seq == "ACCGT" and recs[0].id == "smp_test_0"
    assert recs[0].desc == ("sample=smp gene=test product=0 length=5 kmer_count_mean=20.00 kmer_count_median=20 "
                            "kmer_count_min=10 kmer_count_max=30 score=20.00")
    assert pcr._f64_display(10.5) == "10.5" and pcr._f64_display(7.0) == "7"
    sc = pcr.PathScore(1, 100.0, 2.0, 10.0)
    assert sc.composite() == 100.0 * 0.5 * 0.5
    assert pcr.bounded_levenshtein("ACGTACGT", "ACGTACGT", 0) == 0
    assert pcr.bounded_levenshtein("ACGTACGT", "ACGAACGT", 1) == 1 and pcr.bounded_levenshtein("ACGTACGT", "ACGAACGA", 1) is None
    assert pcr.bounded_levenshtein("ACGT", "ACGTTT", 2) == 2 and pcr.bounded_levenshtein("ACGT", "ACGTTTT", 2) is None
    assert pcr.bounded_levenshtein("", "AC", 2) == 2 and pcr.bounded_levenshtein("KITTEN", "SITTING", 3) == 3
    rng = random.Random(0)
    for _ in range(200):   # against the plain quadratic recurrence
        a = "".join(rng.choice("AC") for _ in range(rng.randint(0, 12)))
        b = "".join(rng.choice("AC") for _ in range(rng.randint(0, 12)))
        d = [[i + j if i * j == 0 else 0 for j in range(len(b) + 1)] for i in range(len(a) + 1)]
        for i in range(1, len(a) + 1):
            for j in range(1, len(b) + 1):
                d[i][j] = min(d[i - 1][j] + 1, d[i][j - 1] + 1, d[i - 1][j - 1] + (a[i - 1] != b[j - 1]))
        for k in (0, 1, 3, 20):
            want = d[-1][-1] if d[-1][-1] <= k else None
            assert pcr.bounded_levenshtein(a, b, k) == want, (a, b, k)
    mk = lambda seq, med: pcr.Record("x", "d", seq, pcr.PathScore(1, med, 0.0, 1.0))
    base = "ACGT" * 30
    near = base[:50] + "T" + base[51:]
    far = "".join(random.Random(1).choice("ACGT") for _ in range(120))
    kept = pcr.sort_and_deduplicate([mk(near, 5.0), mk(base, 9.0), mk(far, 7.0)], P(0, 1000))
    assert [r.seq for r in kept] == [base, far]                            # near is within 10 edits of base


def test_validate_pcr_params():  # pcr/mod.rs:296-401
    ok = PCRParams("ACGTAC", "TTGCAA", gene_name="g")
    assert pcr.validate_pcr_params(ok) == []
    errs = dict(pcr.validate_pcr_params(PCRParams("A", "ACXT", gene_name="", min_count=1, min_length=10, max_length=0)))
    assert "Forward primer sequence is too short: 'A'" in errs
    assert "Invalid nucleotide(s) X in reverse primer ACXT" in errs
    assert "min-length (10) is greater than max-length (0)" in errs and "max-length is 0" in errs
    assert "min-count is 1, must be at least 2" in errs and "Gene name is empty" in errs
    assert "Forward and reverse primers are identical: ACGT" in dict(pcr.validate_pcr_params(PCRParams("ACGT", "ACGT")))


# ---- the 18S case of the reference (pcr/mod.rs:1236-1395) --------------------------------------

def params_18s(**kw):
    return PCRParams(forward_seq="AACCTGGTTGATCCTGCCAGT", reverse_seq="TGATCCTTCTGCAGGTTCACCTAC", gene_name="18s",
                     min_count=3, mismatches=2, trim=15, min_length=0, max_length=2500, **kw)


def table_18s(oracle):
    t = oracle.KmerCounts(21)
    for _ in range(10):
        t.ingest_seq(READ_18S)
    return t


def test_18s_integration(oracle):
    from sharkmer_b200.primers import get_primer_kmers
    t = table_18s(oracle)
    tab = OracleTable(t)
    assert len(t) == len(READ_18S) - 21 + 1 and t.get_n_kmers() == (len(READ_18S) - 21 + 1) * 10
    prm = params_18s()
    (fk, fc), (rk, rc_) = get_primer_kmers(prm, tab, 21)
    assert fk.size == 1 and rk.size == 1
    seed, lookup = pcr.create_seed_graph(fk, rk, 21)
    assert seed.n_nodes == 2
    assert sum(1 for n in seed.node_indices() if seed.nodes[n][1]) == 1
    assert sum(1 for n in seed.node_indices() if seed.nodes[n][2]) == 1
    g, lookup, found, calls = pcr.extend_graph(seed.copy(), dict(lookup), tab, 1, 5, prm, pcr.DEFAULT_MAX_NUM_NODES, 21)
    assert found
    assert len(pcr.get_assembly_paths(g, 21, prm)) >= 1                   # "Expected paths after reverse extension"
    pcr.remove_low_coverage_tips(g, 21, 0.1)
    pcr.reachability_pruning(g)
    assert len(pcr.get_assembly_paths(g, 21, prm)) >= 1                   # "Expected paths after pruning"
    # bidirectional extension meets in the middle: far fewer device calls than nodes
    assert calls < g.n_nodes // 2

    out = pcr.do_pcr(tab, 21, "smp", prm, view_min_count=1)
    assert out.failure_reason is None and len(out.records) == 1
    # the product runs from the forward primer's binding site to the reverse primer's, both included
    a = READ_18S.index("GTTGATCCTGCCAGT")
    b = READ_18S.index(rc("CGCAGGTTCACCTAC")) + 15
    assert out.records[0].seq == READ_18S[a:b]
    assert out.records[0].id == "smp_18s_0" and f"length={b - a} " in out.records[0].desc
    assert "kmer_count_median=10 " in out.records[0].desc and out.records[0].desc.endswith("score=10.00")


def test_18s_failures(oracle):
    tab = OracleTable(table_18s(oracle))
    out = pcr.do_pcr(tab, 21, "s", PCRParams("ACGTTTGACCATGACCA", "TGATCCTTCTGCAGGTTCACCTAC", gene_name="x", min_count=3), view_min_count=1)
    assert out.records == [] and out.failure_reason == "forward primer not found"
    out = pcr.do_pcr(tab, 21, "s", PCRParams("ACGTTTGACCATGACCA", "GGGTTTGACCATGACAA", gene_name="x", min_count=3), view_min_count=1)
    assert out.failure_reason == "forward and reverse primers not found"
    prm = params_18s(max_dfs_states=100_000)
    prm.max_length = 300                                                   # product is ~1.8 kb
    out = pcr.do_pcr(tab, 21, "s", prm, view_min_count=1)
    assert out.records == [] and out.failure_reason == "no path found"
    out = pcr.do_pcr(tab, 21, "s", params_18s(), max_num_nodes=50, view_min_count=1)
    assert out.records == [] and out.failure_reason == "node budget exceeded"


# ---- synthetic reads with planted amplicons ------------------------------------------------------

def make_reads(rng, genome, n_reads, L, err=0.0):
    out = []
    for _ in range(n_reads):
        at = rng.randint(0, len(genome) - L)
        s = genome[at:at + L]
        if err:
            s = "".join(c if rng.random() > err else rng.choice("ACGT") for c in s)
        out.append(s if rng.random() < 0.5 else rc(s))
    return out


@pytest.mark.parametrize("seed,k", [(1, 21), (2, 31), (3, 25)])
def test_recovers_planted_amplicon(oracle, tmp_path, seed, k):
    rng = random.Random(seed)
    rnd = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    fwd, rev = rnd(22), rnd(22)
    insert = rnd(rng.randint(300, 600))
    amplicon = fwd + insert + rc(rev)
    genome = rnd(1500) + amplicon + rnd(1500)
    t = oracle.KmerCounts(k)
    for s in make_reads(rng, genome, 1200, 120, err=0.003):
        t.ingest_seq(s)
    tab = OracleTable(t)
    prm = PCRParams(fwd, rev, gene_name="locus", min_count=2, max_length=2000)
    out = pcr.do_pcr(tab, k, "syn", prm)
    assert out.failure_reason is None and len(out.records) >= 1, out
    trim = min(15, k - 1)
    want = amplicon[len(fwd) - trim:len(amplicon) - (len(rev) - trim)]
    assert out.records[0].seq == want
    assert tab.lookup_calls < out.stats["nodes"]                           # batched per frontier wave
    res = pcr.run_pcr(tab, k, [prm, PCRParams(rnd(20), rnd(20), gene_name="absent")], "syn", str(tmp_path) + "/")
    assert res[0]["status"] == "success" and res[0]["product_lengths"][0] == len(want)
    assert res[1] == {"gene_name": "absent", "status": "fail", "n_products": 0, "product_lengths": [],
                      "failure_reason": "forward and reverse primers not found"}
    fa = open(tmp_path / "syn_locus.fasta").read().split("\n")
    assert fa[0].startswith(">syn_locus_0 sample=syn gene=locus product=0 length=")
    assert "".join(fa[1:]) == want and all(len(l) <= 80 for l in fa[1:]) and len(fa[1]) == 80


def test_threshold_sweep_and_multiple_products(oracle):
    """Three alleles at 5 : 5 : 2 copies, tiled exactly (no sampling noise).  Primer sites are shared
    (12 units), so the sweep starts at 6 units: no allele survives; at the next threshold (4 units)
    the two 5-unit alleles do and the search stops there — the 2-unit allele is never reported.
    The two products differ by more than dedup_edit_threshold edits, tie on score, and are ordered
    by sequence."""
    rng = random.Random(7)
    rnd = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    fwd, rev, ins = rnd(20), rnd(20), rnd(400)
    def variant(step, shift):
        v = list(ins)
        for p in range(20 + shift, 380, step):
            v[p] = {"A": "C", "C": "G", "G": "T", "T": "A"}[v[p]]
        return "".join(v)
    alleles = [(ins, 5), (variant(15, 0), 5), (variant(15, 7), 2)]   # SNPs closer than k: one bubble per allele
    flank_a, flank_b = rnd(300), rnd(300)
    L, k = 100, 21
    t = oracle.KmerCounts(k)
    for body, copies in alleles:
        g = flank_a + fwd + body + rc(rev) + flank_b
        for at in range(0, len(g) - L + 1):
            for _ in range(copies):
                t.ingest_seq(g[at:at + L])
    tab = OracleTable(t)
    out = pcr.do_pcr(tab, k, "syn", PCRParams(fwd, rev, gene_name="g", min_count=2, max_length=1500))
    unit = L - k + 1
    assert out.stats["thresholds"] == pcr.compute_coverage_thresholds(12 * unit, 2)
    want = sorted(fwd[-15:] + body + rc(rev)[:15] for body, c in alleles if c == 5)
    assert [r.seq for r in out.records] == want
    assert [r.id for r in out.records] == ["syn_g_0", "syn_g_1"]
    assert all(f"product={i} " in r.desc and f"kmer_count_min={5 * unit} " in r.desc for i, r in enumerate(out.records))
    # one allele alone: a single product
    t1 = oracle.KmerCounts(k)
    g = flank_a + fwd + ins + rc(rev) + flank_b
    for at in range(0, len(g) - L + 1):
        t1.ingest_seq(g[at:at + L])
        t1.ingest_seq(g[at:at + L])
    out = pcr.do_pcr(OracleTable(t1), k, "syn", PCRParams(fwd, rev, gene_name="g", min_count=2, max_length=1500))
    assert [r.seq for r in out.records] == [fwd[-15:] + ins + rc(rev)[:15]]


# ---- the batched / speculative extension against a literal node-at-a-time restatement ------------

def sequential_extend(graph, node_lookup, table, view_min, min_count, params, max_num_nodes, k):
    """Test-only: extend_graph exactly as graph.rs:322-527 reads — one node at a time, four
    get_canonical probes per node, no batching — to pin the product's wave machinery."""
    from collections import deque
    mask = pcr.get_suffix_mask(k)
    shift = 2 * (k - 1)
    found_path = False
    m = pcr.median_f64(graph.edge_counts())
    median = float(min_count) if m is None else m
    last_median_check = 0
    frontier = deque()
    for n in graph.node_indices():
        if graph.nodes[n][1]:
            frontier.append((n, 0))
        if graph.nodes[n][2]:
            frontier.append((n, 1))
    processed = (set(), set())
    added = (set(n for n in graph.node_indices() if graph.nodes[n][1]), set(n for n in graph.node_indices() if graph.nodes[n][2]))
    def get_canonical(kmer):
        c = table.t.get_canonical(kmer)
        return c if c is not None and c >= view_min else None
    while frontier:
        node, d = frontier.popleft()
        if node in processed[d]:
            continue
        processed[d].add(node)
        n_nodes = graph.n_nodes
        if n_nodes > max_num_nodes:
            break
        if n_nodes > last_median_check and n_nodes - last_median_check > 1000:
            m = pcr.median_f64(graph.edge_counts())
            median = float(min_count) if m is None else m
            last_median_check = n_nodes - n_nodes % 1000
        sub = graph.nodes[node][0]
        cands = []
        for base in range(4):
            kmer = (sub << 2) | base if d == 0 else (base << shift) | sub
            c = get_canonical(kmer)
            if c is not None and c >= min_count:
                cands.append(kmer)
        for kmer in cands:
            new_sub = kmer & mask if d == 0 else kmer >> 2
            if new_sub == sub:
                continue
            count = table.t.get_canonical_count(kmer)
            count = count if count >= view_min else 0
            ex = node_lookup.get(new_sub)
            if ex is not None:
                a, b = (node, ex) if d == 0 else (ex, node)
                if graph.find_edge(a, b) is None:
                    graph.add_edge(a, b, count)
                    if ex in added[1 - d]:
                        found_path = True
            else:
                if float(count) > median * params.high_coverage_ratio:
                    continue
                new = graph.add_node(new_sub)
                node_lookup[new_sub] = new
                added[d].add(new)
                if d == 0:
                    graph.add_edge(node, new, count)
                else:
                    graph.add_edge(new, node, count)
                frontier.append((new, d))
    return graph, found_path


@pytest.mark.parametrize("seed,budget", [(1, 500_000), (2, 500_000), (3, 1500), (4, 500_000)])
def test_wave_extension_equals_sequential(oracle, seed, budget):
    """Branchy, repeat-rich input (shared repeats at several copy numbers, sequencing errors, more
    than 1000 nodes so the median refresh and the high-coverage skip fire; one case runs into the
    node budget): the graphs must be identical node for node and edge for edge."""
    from sharkmer_b200.primers import get_primer_kmers
    rng = random.Random(seed)
    rnd = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    k = 21
    fwd, rev = rnd(21), rnd(21)
    repeat = rnd(150)
    body = rnd(500) + repeat + rnd(300) + repeat + rnd(400)
    genome = rnd(600) + fwd + body + rc(rev) + rnd(600)
    extra = "".join(rnd(200) + repeat for _ in range(6))          # the repeat is also elsewhere, deeper
    t = oracle.KmerCounts(k)
    reads = make_reads(rng, genome, 2500, 100, err=0.01) + make_reads(rng, extra, 1500, 100, err=0.01)
    for s in reads:
        t.ingest_seq(s)
    tab = OracleTable(t)
    prm = PCRParams(fwd, rev, gene_name="g", min_count=2, max_length=5000)
    (fk, _), (rk, _) = get_primer_kmers(prm, tab, k)
    assert fk.size and rk.size
    # (the median starts at min_count, so edges above 10 x min_count are "repeats" until the first
    #  refresh at 1001 nodes: thresholds as low as 2 would stop at the seeds, as in the reference)
    for min_count in (8, 20):
        seed_g, lk = pcr.create_seed_graph(fk, rk, k)
        g1, _, f1, calls = pcr.extend_graph(seed_g.copy(), dict(lk), tab, 2, min_count, prm, budget, k)
        g2, f2 = sequential_extend(seed_g.copy(), dict(lk), tab, 2, min_count, prm, budget, k)
        assert g1.nodes == g2.nodes and g1.edges == g2.edges and f1 == f2
        assert calls <= max(2, g1.n_nodes // 4)
        if seed == 1:
            assert g1.n_nodes > 1000    # far enough for the median refresh to fire
    if budget < 10_000 and g1.n_nodes > budget:
        assert g1.n_nodes <= budget + 8  # the check fires when a node is taken from the frontier


# ---- the C++ twin (sharkmer_b200/host/pcr.hpp) through the test-only mock ABI ----------------------

import subprocess


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("pcr") / "host_harness"
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-o", str(out), os.path.join(HERE, "host", "mock_abi.cpp"),
                    "-lz"], check=True)
    return str(out)


def run_cpp_pcr(harness, tmp_path, table, k, specs, sample="syn", extra=()):
    keys, counts = table.export_sorted()
    tf = tmp_path / "table.txt"
    with open(tf, "w") as f:
        for a, b in zip(keys.tolist(), counts.tolist()):
            f.write(f"{a} {b}\n")
    out = tmp_path / "cpp"
    out.mkdir(exist_ok=True)
    for f in out.iterdir():
        f.unlink()
    cmd = [harness, "-k", str(k), "--table", str(tf), "--sample", sample, "--outdir", str(out) + "/", *extra]
    for s in specs:
        cmd += ["--pcr-primers", s]
    return subprocess.run(cmd, capture_output=True, text=True), out


def spec_of(p):
    return (f"forward={p.forward_seq},reverse={p.reverse_seq},name={p.gene_name},max-length={p.max_length},"
            f"min-length={p.min_length},min-count={p.min_count},mismatches={p.mismatches},trim={p.trim}")


def compare_with_python(harness, tmp_path, oracle_table, k, params_list, sample="syn", min_kmer_count=2, max_nodes=None):
    extra = ["--min-kmer-count", str(min_kmer_count)] + (["--max-nodes", str(max_nodes)] if max_nodes else [])
    r, out = run_cpp_pcr(harness, tmp_path, oracle_table, k, [spec_of(p) for p in params_list], sample, extra)
    assert r.returncode == 0, r.stderr
    py = tmp_path / "py"
    py.mkdir(exist_ok=True)
    for f in py.iterdir():
        f.unlink()
    res = pcr.run_pcr(OracleTable(oracle_table), k, params_list, sample, str(py) + "/", min_kmer_count,
                      max_nodes or pcr.DEFAULT_MAX_NUM_NODES)
    lines = [l for l in r.stdout.split("\n") if l]
    assert len(lines) == len(res)
    for line, want in zip(lines, res):
        head, _, reason = line.partition(" | ")
        gene, status, n, *lengths = head.split()
        assert (gene, status, int(n), [int(x) for x in lengths]) == (want["gene_name"], want["status"], want["n_products"],
                                                                    want["product_lengths"])
        assert reason == (want["failure_reason"] or "")
        fa = f"{sample}_{gene}.fasta"
        if status == "success":
            assert open(out / fa).read() == open(py / fa).read()     # byte-identical FASTA
        else:
            assert not os.path.exists(out / fa)
    return res


def test_cpp_18s_and_failures(harness, oracle, tmp_path):
    t = table_18s(oracle)
    absent = PCRParams("ACGTTTGACCATGACCA", "GGGTTTGACCATGACAA", gene_name="absent", min_count=3)
    short = params_18s()
    short.gene_name, short.max_length = "short", 300
    res = compare_with_python(harness, tmp_path, t, 21, [params_18s(), absent, short], sample="smp", min_kmer_count=1)
    assert [r["status"] for r in res] == ["success", "fail", "fail"]
    assert res[1]["failure_reason"] == "forward and reverse primers not found" and res[2]["failure_reason"] == "no path found"
    res = compare_with_python(harness, tmp_path, t, 21, [params_18s()], sample="smp", min_kmer_count=1, max_nodes=50)
    assert res[0]["failure_reason"] == "node budget exceeded"


@pytest.mark.parametrize("seed,k", [(1, 21), (5, 31)])
def test_cpp_matches_python_on_branchy_input(harness, oracle, tmp_path, seed, k):
    """Two loci, sequencing errors, a repeat shared with deeper sequence elsewhere, two alleles at
    one locus: every FASTA byte and every failure reason must agree between the two hosts."""
    rng = random.Random(seed)
    rnd = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    f1, r1, f2, r2 = rnd(22), rnd(22), rnd(20), rnd(24)
    repeat = rnd(120)
    ins = rnd(350)
    ins_b = list(ins)
    for p in range(30, 330, 14):
        ins_b[p] = {"A": "C", "C": "G", "G": "T", "T": "A"}[ins_b[p]]
    ins_b = "".join(ins_b)
    g_a = rnd(500) + f1 + ins + rc(r1) + rnd(400) + f2 + rnd(200) + repeat + rnd(250) + rc(r2) + rnd(500)
    g_b = g_a.replace(ins, ins_b)
    other = "".join(rnd(150) + repeat for _ in range(5))
    t = oracle.KmerCounts(k)
    for s in make_reads(rng, g_a, 2600, 110, err=0.006) + make_reads(rng, g_b, 2200, 110, err=0.006) + make_reads(rng, other, 900, 110, err=0.006):
        t.ingest_seq(s)
    runs = [PCRParams(f1, r1, gene_name="locus1", min_count=2, max_length=1200),
            PCRParams(f2, r2, gene_name="locus2", min_count=2, max_length=1500, min_length=100),
            PCRParams(f1, r2, gene_name="span", min_count=3, max_length=400)]
    res = compare_with_python(harness, tmp_path, t, k, runs)
    assert res[0]["status"] == "success"


def test_cpp_levenshtein_and_primer_spec_errors(harness, tmp_path):
    rng = random.Random(3)
    for _ in range(60):
        a = "".join(rng.choice("ACG") for _ in range(rng.randint(1, 25)))
        b = "".join(rng.choice("ACG") for _ in range(rng.randint(1, 25)))
        for kk in (0, 2, 10):
            got = int(subprocess.run([harness, "--lev", a, b, str(kk)], capture_output=True, text=True).stdout)
            want = pcr.bounded_levenshtein(a, b, kk)
            assert got == (-1 if want is None else want), (a, b, kk)
    tf = tmp_path / "t.txt"
    tf.write_text("5 3\n")
    def spec(s):
        return subprocess.run([harness, "-k", "5", "--table", str(tf), "--pcr-primers", s], capture_output=True, text=True)
    r = spec("forward=ACGT,reverse=TTGA,name=x,forward=AAAA")
    assert r.returncode == 1 and "Duplicate parameter 'forward' in primer specification" in r.stderr
    r = spec("forward=ACGT,reverse=TTGA,bogus=1")
    assert r.returncode == 1 and "Unexpected parameter: bogus" in r.stderr
    r = spec("forward=ACGT,reverse")
    assert r.returncode == 1 and "Invalid parameter (should be key=value): 'reverse'" in r.stderr
    r = spec("forward=ACGT,reverse=TTGA,name=x,trim=abc")
    assert r.returncode == 1 and "Invalid value for trim: abc" in r.stderr
    r = spec("forward=acgt,reverse=ttga,name=x,min-count=1")
    assert r.returncode == 1 and "min-count is 1, must be at least 2" in r.stderr
    r = spec("forward=ACGT,reverse=TTGA")
    assert r.returncode == 1 and "Gene name is empty" in r.stderr


def test_cpp_host_mirror_equals_device_lookups(harness, oracle, tmp_path):
    """KmerCounts::mirror_to_host (the route that leaves src/pcr's one-k-mer-at-a-time calls unchanged:
    src/kmer/counting.rs:328-336 called from src/pcr/graph.rs:419-430): the host open-addressing copy
    answers every lookup mode exactly like the ABI's skm_lookup_batch, and sPCR over the mirror writes the
    same FASTA bytes as sPCR over batched ABI lookups."""
    t = table_18s(oracle)
    keys, counts = t.export_sorted()
    tf = tmp_path / "table.txt"
    with open(tf, "w") as f:
        for a, b in zip(keys.tolist(), counts.tolist()):
            f.write(f"{a} {b}\n")
    r = subprocess.run([harness, "-k", "21", "--table", str(tf), "--mirror-check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith(f"mirror {keys.size} keys") and r.stdout.strip().endswith("0 mismatches")
    spec = [spec_of(params_18s())]
    a, out_a = run_cpp_pcr(harness, tmp_path, t, 21, spec, "smp", ["--min-kmer-count", "1"])
    fasta_a = open(out_a / "smp_18s.fasta").read()
    b, out_b = run_cpp_pcr(harness, tmp_path, t, 21, spec, "smp", ["--min-kmer-count", "1", "--host-mirror"])
    assert a.returncode == 0 and b.returncode == 0 and a.stdout == b.stdout
    assert open(out_b / "smp_18s.fasta").read() == fasta_a


def test_c4_reduced_cpp_equals_python_and_truth(harness, oracle, tmp_path):
    """BASELINE config C4 at 1/27 size on the oracle table (the full-size run on the device table is
    tests/test_gpu_zz_c4.py): the cnidaria panel from the committed fixture, templates at 50x copy number in a
    random genome, 0.5 % read errors, k = 25.  The C++ host and the Python host write byte-identical FASTA for
    every gene and the first product is the planted amplicon."""
    import c4_data
    k, L = 25, 150
    prm = c4_data.load_panel()
    pool, truth = c4_data.build_pool(prm, k, 100_000, copies=50, seed=4)
    n = pool.size * 75 // L
    lines = c4_data.sample_reads(pool, n, L, 0.005, np.random.default_rng(4))
    t = oracle.KmerCounts(k)
    for i in range(n):
        t.ingest_seq(lines[i, :L].tobytes().decode())
    compare_with_python(harness, tmp_path, t, k, prm, sample="c4")
    for p in prm:
        fa = open(tmp_path / "py" / f"c4_{p.gene_name}.fasta").read().split("\n")
        first = []
        for line in fa[1:]:
            if line.startswith(">") or not line:
                break
            first.append(line)
        assert "".join(first) == truth[p.gene_name], p.gene_name
