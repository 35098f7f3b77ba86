"""The N>1 path on CPU: world_size-2 and -3 `gloo` runs of sharkmer_b200.multigpu.ShardedCounter
(the driver the GPU bench uses) with a TEST-ONLY stand-in for the per-GPU engine.

On a GPU, skm_mg_finalize (C, CUDA) does the work and borrows one primitive from the host: an
all-gather of host bytes (skm_comm).  The stand-in below follows the same protocol in Python and
reaches the other ranks ONLY through that callback, called through its C function pointer exactly
as the library calls it: header all-gather, directory all-gather, count in chunk order, totals
all-gather, histogram-column all-gather + sum.  (Its k-mers travel inside the directory message;
on GPUs they travel by peer copies.)  Checked: the ownership rule (owner = floor(hash * N / 2^64)),
the callback marshalling over gloo, the chunk order, the column sum — the merged result must equal
the single-process oracle, independent of N."""
import ctypes as C
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
    """The multi-GPU interface of kmer.Engine over host memory; counting by a dict."""

    def __init__(self, oracle, common, reads_by_chunk, world, rank):
        self.o, self.common, self.world, self.rank = oracle, common, world, rank
        self.reads = reads_by_chunk
        self.table = {}
        self.cols = {}
        self.chunks, self.n_chunks = CHUNKS, CHUNKS
        self.sent = 0

    def _kmers(self, c):
        out = []
        for s in self.reads[c]:
            out.extend(self.o.kmers_from_ascii(s, K))
        return np.array(out, dtype=np.uint64)

    def _allgather(self, comm, arr: np.ndarray) -> np.ndarray:
        """One call of the borrowed primitive, through the C function pointer."""
        send = np.ascontiguousarray(arr)
        recv = np.empty(self.world * send.size, dtype=send.dtype)
        rc = comm.allgather(comm.user, send.ctypes.data, recv.ctypes.data, send.nbytes)
        assert rc == 0
        return recv.reshape(self.world, -1)

    def mg_finalize(self, comm):
        own = self.common.owner_rank
        # (1) header: how many k-mers this rank sends to each owner, per chunk
        per_chunk = [self._kmers(c) for c in range(CHUNKS)]
        owners = [np.array([own(self.common.hash_kmer(int(x)), self.world) for x in km], dtype=np.int64) for km in per_chunk]
        counts = np.array([[int((o == d).sum()) for d in range(self.world)] for o in owners], dtype=np.uint64)  # [chunk, dst]
        allc = self._allgather(comm, counts.reshape(-1)).reshape(self.world, CHUNKS, self.world)              # [src, chunk, dst]
        # (2) directory + payload, padded to the largest message
        width = int(allc.sum(axis=(1, 2)).max())
        msg = np.zeros(max(width, 1), dtype=np.uint64)
        at = 0
        for c in range(CHUNKS):
            for d in range(self.world):
                sel = per_chunk[c][owners[c] == d]
                msg[at:at + sel.size] = sel
                at += sel.size
        self.sent = 8 * int(counts.sum() - counts[:, self.rank].sum())
        allm = self._allgather(comm, msg)
        # (3) count what this rank owns, chunk by chunk, a histogram column after each
        for c in range(CHUNKS):
            for s in range(self.world):
                start = int(allc[s, :c, :].sum() + allc[s, c, :self.rank].sum())
                for x in allm[s, start:start + int(allc[s, c, self.rank])].tolist():
                    assert own(self.common.hash_kmer(x), self.world) == self.rank
                    self.table[x] = self.table.get(x, 0) + 1
            v = np.zeros(HMAX + 2, dtype=np.uint64)
            for n in self.table.values():
                v[n if n <= HMAX else HMAX + 1] += 1
            self.cols[c] = v
        # (4) totals: the conservation identity on the global sums (src/io.rs:1042-1047)
        tot = self._allgather(comm, np.array([sum(self.table.values()), int(counts.sum())], dtype=np.uint64))
        assert int(tot[:, 0].sum()) == int(tot[:, 1].sum())
        # (5) columns summed over the partitions
        cols = self._allgather(comm, np.stack([self.cols[c] for c in range(CHUNKS)]).reshape(-1))
        total = cols.reshape(self.world, CHUNKS, HMAX + 2).sum(axis=0)
        for c in range(CHUNKS):
            self.cols[c] = total[c]

    def mg_bytes_sent(self):
        return self.sent

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
    sc = ShardedCounter(eng, torch.device("cpu"))
    cols = sc.finalize()
    assert sc.comm.calls == 4 and sc.bytes_sent == eng.sent
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
    sc = ShardedCounter(eng, torch.device("cpu"))
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
