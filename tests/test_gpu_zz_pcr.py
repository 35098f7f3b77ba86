"""sPCR end to end on the device table: reads -> skm_ingest_batch -> skm_finalize -> primer scans
(skm_scan_oligos) -> graph extension by batched skm_lookup_batch waves -> amplicon records, compared
with the same pipeline over the CPU oracle's table and with the planted truth.  (Named to sort after
the other GPU tests: it widens past the counting path)"""
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


@pytest.fixture(scope="module")
def skm():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from sharkmer_b200 import kmer
    return kmer


def test_18s_amplicon_on_device(skm, oracle):  # src/pcr/mod.rs:1236-1395
    from sharkmer_b200 import pcr
    from sharkmer_b200.primers import PCRParams
    from test_pcr import OracleTable
    read = open(os.path.join(HERE, "golden", "pcr_18s_read.txt")).read().strip()
    e = skm.Engine(21, chunks=0)
    t = oracle.KmerCounts(21)
    for _ in range(10):
        e.ingest_batch(0, (read + "\n").encode())
        t.ingest_seq(read)
    e.finalize()
    prm = PCRParams(forward_seq="AACCTGGTTGATCCTGCCAGT", reverse_seq="TGATCCTTCTGCAGGTTCACCTAC", gene_name="18s",
                    min_count=3, mismatches=2, trim=15, max_length=2500)
    got = pcr.do_pcr(e, 21, "smp", prm, view_min_count=1)
    want = pcr.do_pcr(OracleTable(t), 21, "smp", prm, view_min_count=1)
    assert got.failure_reason is None and len(got.records) == 1
    assert [(r.id, r.desc, r.seq) for r in got.records] == [(r.id, r.desc, r.seq) for r in want.records]
    a = read.index("GTTGATCCTGCCAGT")
    b = read.index(rc("CGCAGGTTCACCTAC")) + 15
    assert got.records[0].seq == read[a:b]
    assert got.stats["nodes"] == want.stats["nodes"] and got.stats["edges"] == want.stats["edges"]
    assert got.stats["lookup_calls"] < got.stats["nodes"] // 4   # batched, speculative waves


def test_planted_amplicon_from_reads_on_device(skm, oracle, tmp_path):
    from sharkmer_b200 import pcr
    from sharkmer_b200.primers import PCRParams
    from test_pcr import OracleTable
    rng = random.Random(11)
    rnd = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    fwd, rev, insert = rnd(22), rnd(22), rnd(450)
    amplicon = fwd + insert + rc(rev)
    genome = rnd(2000) + amplicon + rnd(2000)
    L = 120
    reads = []
    for _ in range(1600):
        at = rng.randint(0, len(genome) - L)
        s = "".join(c if rng.random() > 0.003 else rng.choice("ACGT") for c in genome[at:at + L])
        reads.append(s if rng.random() < 0.5 else rc(s))
    for k in (21, 31):
        e = skm.Engine(k, chunks=0)
        e.ingest_batch(0, ("\n".join(reads) + "\n").encode())
        e.finalize()
        t = oracle.KmerCounts(k)
        for s in reads:
            t.ingest_seq(s)
        prm = PCRParams(fwd, rev, gene_name="locus", min_count=2, max_length=2000)
        got = pcr.do_pcr(e, k, "syn", prm)
        want = pcr.do_pcr(OracleTable(t), k, "syn", prm)
        assert got.failure_reason is None
        assert [(r.id, r.desc, r.seq) for r in got.records] == [(r.id, r.desc, r.seq) for r in want.records]
        trim = min(15, k - 1)
        assert got.records[0].seq == amplicon[len(fwd) - trim:len(amplicon) - (len(rev) - trim)]
    res = pcr.run_pcr(e, 31, [prm], "syn", str(tmp_path) + "/")
    assert res[0]["status"] == "success" and os.path.exists(tmp_path / "syn_locus.fasta")


def test_cli_pcr_primers_fasta(skm, oracle, tmp_path):
    """reads.fastq -> sharkmer_b200_cli --pcr-primers ... -> {sample}_{gene}.fasta and the pcr_results
    block of the stats file; the FASTA must equal what the Python pipeline writes from the oracle's
    table of the same reads."""
    import subprocess
    from sharkmer_b200 import pcr
    from sharkmer_b200.primers import PCRParams
    from test_pcr import OracleTable
    cli = os.path.join(os.path.dirname(HERE), "sharkmer_b200", "host", "sharkmer_b200_cli")
    rng = random.Random(21)
    rnd = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    fwd, rev, insert = rnd(21), rnd(23), rnd(380)
    genome = rnd(1500) + fwd + insert + rc(rev) + rnd(1500)
    L = 100
    reads = []
    for _ in range(1500):
        at = rng.randint(0, len(genome) - L)
        s = "".join(c if rng.random() > 0.004 else rng.choice("ACGT") for c in genome[at:at + L])
        reads.append(s if rng.random() < 0.5 else rc(s))
    fq = tmp_path / "reads.fastq"
    fq.write_text("".join(f"@r{i}\n{s}\n+\n{'I' * len(s)}\n" for i, s in enumerate(reads)))
    out = tmp_path / "out"
    out.mkdir()
    spec = f"forward={fwd.lower()},reverse={rev},name=locus,max-length=1500"
    r = subprocess.run([cli, "-k", "21", "--chunks", "2", "--histo-max", "100", "-s", "smp", "-o", str(out),
                        "--pcr-primers", spec, "--pcr-primers", f"forward={rnd(20)},reverse={rnd(20)},name=absent", str(fq)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    t = oracle.KmerCounts(21)
    for s in reads:
        t.ingest_seq(s)
    py = tmp_path / "py"
    py.mkdir()
    res = pcr.run_pcr(OracleTable(t), 21, [PCRParams(fwd, rev, gene_name="locus", max_length=1500)], "smp", str(py) + "/")
    assert res[0]["status"] == "success"
    assert open(out / "smp_locus.fasta").read() == open(py / "smp_locus.fasta").read()
    assert not os.path.exists(out / "smp_absent.fasta")
    stats = open(out / "smp.stats.yaml").read()
    assert "pcr_results:\n- gene_name: locus\n  status: success\n  n_products: 1\n  product_lengths:\n  - " in stats
    assert "- gene_name: absent\n  status: fail\n  n_products: 0\n  failure_reason: forward and reverse primers not found" in stats
    bad = subprocess.run([cli, "-k", "21", "--pcr-primers", "forward=ACGT,reverse=ACGT,name=x", str(fq)], capture_output=True, text=True)
    assert bad.returncode == 1 and "Forward and reverse primers are identical" in bad.stderr
