"""Full-size goldens for BASELINE.json's CPU-runnable configurations, from the CPU oracle (which is
pinned to the reference's KATs by tests/test_oracle_kats.py).  Takes a few minutes and ~20 GB of RAM:

    python tests/golden/make_golden_full.py            # writes tests/golden/full_cases.json

  C1  sharkmer -k 31 --max-reads 1000000 on 1 M synthetic 150 bp reads, --chunks 0 and --chunks 1
      (10 Mbp genome, 1 % substitutions, 0.1 % N, seed 1)
  C2  k=21, 10 chunks, 10 M reads (50 Mbp genome, 1 % substitutions, 0.1 % N, seed 2) — the bench workload
Stored per case: totals, the table digest (wrapping sum of skm_pair_digest over every (k-mer, count)
pair, include/skm_common.h) and every histogram column (non-zero bins only).
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as o  # noqa: E402

CASES = [
    # name, k, chunks, histo_max, seed, genome_len, read_len, sub_rate, n_rate, n_reads
    ("C1_chunks0", 31, 0, 10000, 1, 10_000_000, 150, 0.01, 0.001, 1_000_000),
    ("C1_chunks1", 31, 1, 10000, 1, 10_000_000, 150, 0.01, 0.001, 1_000_000),
    ("C2", 21, 10, 10000, 2, 50_000_000, 150, 0.01, 0.001, 10_000_000),
]


def sparse(v):
    return {str(i): int(x) for i, x in enumerate(v.tolist()) if x}


def main(only=None):
    path = os.path.join(HERE, "full_cases.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for name, k, chunks, hmax, seed, G, L, e, n, nreads in CASES:
        if only and name not in only:
            continue
        t0 = time.time()
        run = o.Run(k, chunks, hmax)
        step = 500_000
        for first in range(0, nreads, step):          # generate and feed in pieces (1000-read batches inside)
            run.push_lines(o.synth_reads(seed, G, L, e, n, first, min(step, nreads - first)))
        run.finish()
        t = run.table()
        rec = {"k": k, "chunks": chunks, "histo_max": hmax, "seed": seed, "genome_len": G, "read_len": L,
               "sub_rate": e, "n_rate": n, "n_reads": nreads,
               "n_reads_ingested": run.n_reads_ingested, "n_bases_ingested": run.n_bases_ingested,
               "n_kmers": run.n_kmers_ingested, "n_unique": t.len(), "digest": t.digest(),
               "oracle_seconds": round(time.time() - t0, 1)}
        if chunks:
            rec["histograms"] = [sparse(run.histogram(c)) for c in range(chunks)]
        out[name] = rec
        print(name, {kk: vv for kk, vv in rec.items() if kk != "histograms"}, flush=True)
        json.dump(out, open(path, "w"))
        del run, t


if __name__ == "__main__":
    main(set(sys.argv[1:]) or None)
