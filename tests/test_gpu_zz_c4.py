"""BASELINE config C4 at full size: the cnidaria panel (k = 25) on a device table counted from 10 M reads.

Reads (150 bp, 0.5 % substitutions, both strands) are sampled from a 20 Mbp random genome plus the panel's
amplicon templates at 50x copy number (tests/c4_data.py; SURVEY.md §8's config table).  Two independent
hosts then run the in silico PCR on the device table of those reads:
  * Python (sharkmer_b200.pcr) through skm_scan_oligos + batched skm_lookup_batch waves (route B:
    device lookups), fed by skm_ingest_batch;
  * C++ (sharkmer_b200_cli --pcr-primers ... --host-mirror) from a FASTQ file of the same reads, with
    the table copied once into a host hash table and probed there (route A: src/pcr unchanged).
Required: every gene's FASTA byte-identical between the two, and the first product of every gene equal to
the planted amplicon, base for base.  SKM_C4_READS scales the run down (default 10_000_000)."""
import os
import subprocess
import sys
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
CLI = os.path.join(ROOT, "sharkmer_b200", "host", "sharkmer_b200_cli")


def test_c4_cnidaria_panel_full_size(tmp_path):
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import c4_data
    from sharkmer_b200 import kmer, panels, pcr
    k, L, err = 25, 150, 0.005
    n_reads = int(os.environ.get("SKM_C4_READS", "10000000"))
    genome_len = max(200_000, n_reads * 2)       # 75x genome coverage, 3750x on the templates
    prm = c4_data.load_panel()
    pool, truth = c4_data.build_pool(prm, k, genome_len, copies=50, seed=4)
    rng = np.random.default_rng(4)
    fq = tmp_path / "c4.fastq"
    e = kmer.Engine(k, chunks=0)
    t0 = time.perf_counter()
    with open(fq, "wb") as f:
        done = 0
        while done < n_reads:
            nb = min(500_000, n_reads - done)
            lines = c4_data.sample_reads(pool, nb, L, err, rng)
            e.ingest_batch(0, lines.reshape(-1))
            c4_data.fastq_bytes(lines).tofile(f)
            done += nb
    e.finalize()
    t_count = time.perf_counter() - t0
    tot = e.totals()
    assert int(tot.n_reads) == n_reads and int(tot.n_kmers) == n_reads * (L - k + 1)

    py = tmp_path / "py"
    py.mkdir()
    t0 = time.perf_counter()
    budget = pcr.compute_node_budget(int(tot.n_bases))     # main.rs:146-177, as the C++ host does
    res = pcr.run_pcr(e, k, prm, "c4", str(py) + "/", 2, budget)
    t_py = time.perf_counter() - t0
    assert {r["gene_name"]: r["status"] for r in res} == {p.gene_name: "success" for p in prm}, res

    cc = tmp_path / "cc"
    cc.mkdir()
    args = [CLI, "-k", str(k), "-s", "c4", "-o", str(cc) + "/", "--host-mirror"]
    for p in prm:
        args += ["--pcr-primers", panels.to_pcr_primers_spec(p)]
    t0 = time.perf_counter()
    r = subprocess.run(args + [str(fq)], capture_output=True, text=True)
    t_cli = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-2000:]

    for p in prm:
        name = f"c4_{p.gene_name}.fasta"
        a, b = open(py / name).read(), open(cc / name).read()
        assert a == b, p.gene_name
        first = []
        for line in a.split("\n")[1:]:
            if line.startswith(">") or not line:
                break
            first.append(line)
        assert "".join(first) == truth[p.gene_name], p.gene_name
    print(f"\nC4: {n_reads} reads, {int(tot.n_kmers)} k-mers, {int(tot.n_unique)} distinct | sampling + counting {t_count:.1f} s | "
          f"python sPCR on the device table {t_py:.1f} s | C++ host (FASTQ file -> count -> host mirror -> sPCR) {t_cli:.1f} s")
    e.close()
