"""The N>1 path on CPU: world_size-2 `gloo` run of sharkmer_b200.multigpu.ShardedCounter
(the driver code the GPU bench uses) with a TEST-ONLY stand-in for the per-GPU engine that
gets its k-mers from the oracle.  Checks the routing rule (owner = floor(hash * N / 2^64),
buckets = (owner, region)), the split sizes, the per-chunk barrier order and the histogram all-reduce: the merged result
must equal the single-process oracle, independent of N."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K, CHUNKS, HMAX, L, NREADS = 21, 4, 50, 100, 9000


class FakeEngine:
    """Routing interface of kmer.Engine over host memory; counting by a dict."""

    def __init__(self, oracle, common, reads_by_chunk, world, rank):
        self.o, self.common, self.world, self.rank = oracle, common, world, rank
        self.reads = reads_by_chunk
        self.table = {}
        self.cols = {}
        self.routed = {}
        self.keep = []

    def finalize_external(self): pass
    def sync(self): pass

    def _kmers(self, c):
        out = []
        for s in self.reads[c]:
            out.extend(self.o.kmers_from_ascii(s, K))
        return np.array(out, dtype=np.uint64)

    REGIONS = 4

    def route_regions(self):
        return self.REGIONS

    def route_count(self, c, world):
        km = self._kmers(c)
        log2r = self.REGIONS.bit_length() - 1
        buckets = np.array([self.common.route_bucket(int(x), world, log2r) for x in km], dtype=np.int64)
        order = np.argsort(buckets, kind="stable")
        self.routed[c] = km[order]
        return np.bincount(buckets, minlength=world * self.REGIONS).astype(np.uint64).reshape(world, self.REGIONS)

    def route_scatter(self, c, ptr):
        km = self.routed.pop(c)
        dst = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_int64 * max(len(km), 1)).from_address(ptr))
        dst[:len(km)] = km.view(np.int64)

    def insert_runs_device(self, ptr, run_counts):
        n = int(run_counts.sum())
        assert run_counts.shape == (self.world, self.REGIONS)
        if n == 0:
            return
        a = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_int64 * n).from_address(ptr)).view(np.uint64)
        # every run must hold k-mers of exactly that (this rank, region) bucket
        log2r = self.REGIONS.bit_length() - 1
        pos = 0
        for s in range(self.world):
            for r in range(self.REGIONS):
                for x in a[pos:pos + int(run_counts[s, r])].tolist():
                    assert self.common.route_bucket(x, self.world, log2r) == self.rank * self.REGIONS + r
                    self.table[x] = self.table.get(x, 0) + 1
                pos += int(run_counts[s, r])

    def snapshot_histogram(self, c):
        v = np.zeros(HMAX + 2, dtype=np.uint64)
        for n in self.table.values():
            v[n if n <= HMAX else HMAX + 1] += 1
        self.cols[c] = v

    def histogram(self, c):
        return self.cols[c]

    # table services over this rank's shard (what Engine.lookup / Engine.scan_oligos do on the device)
    def lookup(self, kmers, min_count, mode):
        rc = self.common.revcomp_kmer
        counts = np.zeros(len(kmers), dtype=np.uint32)
        for i, x in enumerate(np.asarray(kmers, dtype=np.uint64).tolist()):
            r = rc(x, K)
            c = self.table.get(min(x, r) if mode == 0 else x, 0)
            if mode == 2 and not c:
                c = self.table.get(r, 0)
            counts[i] = c if c >= min_count else 0
        return counts, counts > 0

    def scan_oligos(self, oligos, length, min_count):
        want = set(int(x) for x in oligos)
        out = {}
        for x, c in self.table.items():
            if c < min_count:
                continue
            if (x >> (2 * (K - length))) in want:
                out[x] = c
            else:
                r = self.common.revcomp_kmer(x, K)
                if (r >> (2 * (K - length))) in want:
                    out[r] = c
        keys = np.array(sorted(out), dtype=np.uint64)
        return keys, np.array([out[int(x)] for x in keys], dtype=np.uint32)


def worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as o
    from sharkmer_b200 import common
    from sharkmer_b200.multigpu import ShardedCounter
    reads = o.synth_reads(5, 20_000, L, 0.01, 0.002, 0, NREADS).tobytes().decode().split("\n")[:-1]
    # chunk c = batches b = c (mod CHUNKS); rank r takes the r-th slice of each chunk's batches
    by_chunk = []
    for c in range(CHUNKS):
        batches = list(range(c, (NREADS + 999) // 1000, CHUNKS))
        lo, hi = len(batches) * rank // world, len(batches) * (rank + 1) // world
        mine = []
        for b in batches[lo:hi]:
            mine.extend(reads[b * 1000:(b + 1) * 1000])
        by_chunk.append(mine)
    eng = FakeEngine(o, common, by_chunk, world, rank)
    sc = ShardedCounter(eng, CHUNKS, CHUNKS, HMAX, torch.device("cpu"))
    cols = sc.finalize()
    # every key this rank holds must be one it owns
    assert all(common.owner_rank(common.hash_kmer(x), world) == rank for x in eng.table)
    tot = sc.global_totals({"n_unique": len(eng.table), "n_kmers": sum(eng.table.values())})
    gathered = [None] * world
    dist.all_gather_object(gathered, eng.table)
    if rank == 0:
        merged = {}
        for t in gathered:
            assert not (set(t) & set(merged))  # partitions are disjoint
            merged.update(t)
        q.put((cols, tot, merged))
    dist.destroy_process_group()


def planted_reads():
    import random
    rng = random.Random(17)
    rnd = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    rcs = lambda x: x[::-1].translate(str.maketrans("ACGT", "TGCA"))
    fwd, rev, insert = rnd(22), rnd(22), rnd(300)
    genome = rnd(700) + fwd + insert + rcs(rev) + rnd(700)
    reads = []
    for _ in range(4000):
        at = rng.randint(0, len(genome) - L)
        seq = "".join(c if rng.random() > 0.003 else rng.choice("ACGT") for c in genome[at:at + L])
        reads.append(seq if rng.random() < 0.5 else rcs(seq))
    return fwd, rev, fwd[-15:] + insert + rcs(rev)[:15], reads


def worker_pcr(rank, world, port, q):
    """Count planted-amplicon reads on a sharded (fake) table, then run sPCR on it through
    ShardedCounter.lookup / scan_oligos: every rank must recover the same amplicon."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as o
    from sharkmer_b200 import common, pcr
    from sharkmer_b200.multigpu import ShardedCounter
    from sharkmer_b200.primers import PCRParams
    fwd, rev, want, reads = planted_reads()
    by_chunk = []
    for c in range(CHUNKS):
        batches = list(range(c, (len(reads) + 999) // 1000, CHUNKS))
        lo, hi = len(batches) * rank // world, len(batches) * (rank + 1) // world
        mine = []
        for b in batches[lo:hi]:
            mine.extend(reads[b * 1000:(b + 1) * 1000])
        by_chunk.append(mine)
    eng = FakeEngine(o, common, by_chunk, world, rank)
    sc = ShardedCounter(eng, CHUNKS, CHUNKS, HMAX, torch.device("cpu"))
    sc.finalize()
    out = pcr.do_pcr(sc, K, "smp", PCRParams(fwd, rev, gene_name="locus", max_length=1000))
    probe = np.array(list(eng.table)[:50] + [12345], dtype=np.uint64)   # this rank's keys, asked of everybody
    everybody = [None] * world
    dist.all_gather_object(everybody, probe.tolist())
    probe = np.array(everybody[0], dtype=np.uint64)                       # same query on every rank
    counts, found = sc.lookup(probe, 0, 2)
    if rank == 0:
        q.put(([(r.id, r.seq) for r in out.records], out.failure_reason, want, counts.tolist(), found.tolist(), probe.tolist()))
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_counter_matches_oracle(oracle, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    cols, tot, merged = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    run = oracle.Run(K, CHUNKS, HMAX)
    run.push_lines(oracle.synth_reads(5, 20_000, L, 0.01, 0.002, 0, NREADS))
    run.finish()
    keys, counts = run.table().export_sorted()
    assert merged == dict(zip(keys.tolist(), counts.tolist()))
    assert tot == {"n_unique": keys.size, "n_kmers": int(run.n_kmers_ingested)}
    for c in range(CHUNKS):
        assert (cols[c] == run.histogram(c)).all(), c


@pytest.mark.parametrize("world", [2, 3])
def test_spcr_on_sharded_table(oracle, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker_pcr, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    records, failure, want, counts, found, probe = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert failure is None and records == [("smp_locus_0", want)]
    # the collective lookup answers like one table holding everything
    t = oracle.KmerCounts(K)
    for s in planted_reads()[3]:
        t.ingest_seq(s)
    for x, c, f in zip(probe, counts, found):
        ref = t.get_canonical(x)
        assert (c, f) == ((ref, True) if ref is not None else (0, False))
