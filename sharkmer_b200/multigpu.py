"""Hash-sharded k-mer counting over N GPUs: one process per GPU, torch.distributed
for the plumbing (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests).

The path shards by k-mer hash: every rank packs and extracts the
k-mers of ITS slice of the reads, buckets them by owner rank
(skm_owner_rank(hash, n_ranks), include/skm_common.h), and an all-to-all delivers
each bucket to the rank that owns that slice of the table; the owner inserts them
with no cross-GPU atomics.  With --chunks n > 0 every chunk boundary is a global
barrier: all ranks finish exchanging and inserting chunk i before histogram column i
is taken (src/io.rs:1016-1028 merges chunks in index order).  The result is
independent of N.

The exchange of chunk c+1 (route on the compute stream, all-to-all on the
collective's stream) overlaps the insert of chunk c: send/receive buffers are
double-buffered torch tensors.

`engine` is anything with the routing interface of sharkmer_b200.kmer.Engine
(route_count / route_scatter / insert_kmers_device / snapshot_histogram / histogram /
finalize_external); the CPU tests drive this same code with a recording stand-in.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


class ShardedCounter:
    def __init__(self, engine, n_chunks: int, chunks_arg: int, histo_max: int, device: torch.device,
                 group=None, stream=None):
        self.e = engine
        self.n_chunks = n_chunks
        self.chunks_arg = chunks_arg
        self.histo_max = histo_max
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.stream = stream  # torch.cuda.Stream the engine launches on (None on CPU)
        self.bytes_sent = 0
        self.kmers_received = 0

    # -- small helpers ---------------------------------------------------------
    def _exchange_counts(self, counts: np.ndarray) -> np.ndarray:
        send = torch.as_tensor(counts.astype(np.int64), device=self.device)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        return recv.cpu().numpy()

    def _alloc(self, n: int) -> torch.Tensor:
        return torch.empty(max(int(n), 1), dtype=torch.int64, device=self.device)

    def _ptr(self, t: torch.Tensor) -> int:
        return t.data_ptr()

    def _on_stream(self):
        if self.stream is not None:
            return torch.cuda.stream(self.stream)
        import contextlib
        return contextlib.nullcontext()

    # -- the chunk loop ----------------------------------------------------------
    def finalize(self) -> np.ndarray | None:
        """Counts every chunk in order.  Returns the (n_chunks, histo_max+2) cumulative
        histogram columns summed over ranks (None when chunks == 0)."""
        # Everything below is issued with the engine's stream current: tensors are allocated on
        # it, collectives are ordered after the kernels already queued on it, and work.wait()
        # makes it (not the default stream) wait for the collective.
        with self._on_stream():
            return self._finalize()

    def _finalize(self):
        e = self.e
        e.finalize_external()  # ingest is complete on every rank; the chunk loop is ours

        def launch_exchange(c):
            counts = e.route_count(c, self.world)            # per-owner counts of this rank's k-mers
            rcounts = self._exchange_counts(counts)          # how many each rank sends us
            n_send, n_recv = int(counts.sum()), int(rcounts.sum())
            send, recv = self._alloc(n_send), self._alloc(n_recv)
            e.route_scatter(c, self._ptr(send))              # k-mers grouped by destination (engine stream)
            work = dist.all_to_all_single(recv[:n_recv], send[:n_send],
                                          output_split_sizes=[int(x) for x in rcounts],
                                          input_split_sizes=[int(x) for x in counts],
                                          group=self.group, async_op=True)
            self.bytes_sent += 8 * (n_send - int(counts[self.rank]))
            self.kmers_received += n_recv
            return work, send, recv, n_recv, c

        pending = launch_exchange(0)
        for c in range(self.n_chunks):
            work, send, recv, n_recv, cc = pending
            # start routing the next chunk while this chunk's k-mers are in flight
            nxt = launch_exchange(c + 1) if c + 1 < self.n_chunks else None
            work.wait()                                       # engine stream waits for the collective
            e.insert_kmers_device(self._ptr(recv), n_recv)
            if self.chunks_arg > 0:
                e.snapshot_histogram(cc)                      # this rank's partial column (syncs the stream)
            else:
                e.sync()
            del send, recv
            pending = nxt
        if self.chunks_arg == 0:
            return None
        cols = np.stack([e.histogram(c) for c in range(self.n_chunks)]).astype(np.int64)
        t = torch.as_tensor(cols, device=self.device)
        dist.all_reduce(t, group=self.group)                  # histogram = sum of the partitions' histograms
        return t.cpu().numpy().astype(np.uint64)

    def global_totals(self, local: dict) -> dict:
        keys = sorted(local)
        t = torch.as_tensor([int(local[k]) for k in keys], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, group=self.group)
        return dict(zip(keys, t.cpu().tolist()))
