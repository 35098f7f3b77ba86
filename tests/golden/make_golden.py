"""Regenerates the committed golden fixtures.  Run in the build container
(needs /root/reference for the 18S string; everything else comes from the
oracle, which is pinned to the reference's KATs by tests/test_oracle_kats.py).

    python tests/golden/make_golden.py
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as o  # noqa: E402

REF_PCR = "/root/reference/src/pcr/mod.rs"


def golden_18s():
    # src/pcr/mod.rs:1236-1247 (build_test_case) and :1336-1342 (test_integration)
    s = re.search(r'let read_string = "([ACGT]+)"', open(REF_PCR).read()).group(1)
    with open(os.path.join(HERE, "pcr_18s_read.txt"), "w") as f:
        f.write(s + "\n")
    kc = o.KmerCounts(21)
    for _ in range(10):
        kc.ingest_seq(s)
    json.dump({"source": "src/pcr/mod.rs:1236-1247,1336-1342; test data string in pcr_18s_read.txt",
               "k": 21, "replicates": 10, "n_bases": len(s), "n_distinct": kc.len(),
               "n_kmers": kc.get_n_kmers(), "digest": kc.digest()},
              open(os.path.join(HERE, "pcr_18s_k21.json"), "w"), indent=1)


SYNTH_CASES = [
    # name, k, chunks, histo_max, seed, genome_len, read_len, sub_rate, n_rate, n_reads
    ("c1_small", 31, 1, 100, 1, 200_000, 150, 0.01, 0.001, 20_000),
    ("c2_small", 21, 10, 100, 2, 100_000, 150, 0.01, 0.001, 25_500),
    ("c4_small", 25, 0, 100, 4, 50_000, 150, 0.005, 0.0, 10_000),
    ("short_reads", 31, 3, 50, 9, 5_000, 40, 0.02, 0.01, 7_777),
]


def golden_synth():
    out = {}
    for name, k, chunks, hmax, seed, G, L, e, n, nreads in SYNTH_CASES:
        reads = o.synth_reads(seed, G, L, e, n, 0, nreads)
        run = o.Run(k, chunks, hmax)
        run.push_lines(reads)
        run.finish()
        t = run.table()
        rec = {"k": k, "chunks": chunks, "histo_max": hmax, "seed": seed, "genome_len": G, "read_len": L,
               "sub_rate": e, "n_rate": n, "n_reads": nreads,
               "n_bases_ingested": run.n_bases_ingested, "n_kmers": run.n_kmers_ingested,
               "n_unique": t.len(), "digest": t.digest(),
               "chunk_totals": [list(run.chunk_totals(c)) for c in range(run.n_chunks)]}
        if chunks:
            rec["histograms"] = [run.histogram(c).tolist() for c in range(chunks)]
        out[name] = rec
    json.dump(out, open(os.path.join(HERE, "synth_cases.json"), "w"))


if __name__ == "__main__":
    golden_18s()
    golden_synth()
    print("golden fixtures written to", HERE)
