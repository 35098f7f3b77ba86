"""Hash-sharded k-mer counting over N GPUs: one process per GPU, torch.distributed
for the plumbing (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests).

The path shards by k-mer hash: every rank packs and extracts the
k-mers of ITS slice of the reads and buckets them by (owner rank, table region of that
owner) — owner = floor(hash * N / 2^64), include/skm_common.h — so one pass both groups
the k-mers by destination and leaves every destination's run sorted by the receiver's
table regions.  An all-to-all delivers each rank's range; the owner inserts the runs
region by region (L2-resident) with no cross-GPU atomics.  With --chunks n > 0 every chunk boundary is a global
barrier: all ranks finish exchanging and inserting chunk i before histogram column i
is taken (src/io.rs:1016-1028 merges chunks in index order).  The result is
independent of N.

Two exchange paths:
  * "p2p" (default on GPUs, N <= 16): fused route + exchange.  The scatter kernel stores every
    destination's runs straight into that rank's receive arena through a CUDA-IPC mapping (peer
    stores over NVLink/NVSwitch) — no local list, no collective copy; the transfer overlaps the
    extraction tile by tile inside one kernel.  Steps are ordered by a 1-element all-reduce used
    as a stream-ordered barrier; two arenas per rank give double buffering.
  * "nccl": route to a local list, then all_to_all_single; the exchange of chunk c+1 overlaps
    the insert of chunk c (double-buffered torch tensors).  Also the path of the CPU tests (gloo).

`engine` is anything with the routing interface of sharkmer_b200.kmer.Engine
(route_count / route_scatter / insert_runs_device / snapshot_histogram / histogram /
finalize_external); the CPU tests drive this same code with a recording stand-in.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


class ShardedCounter:
    def __init__(self, engine, n_chunks: int, chunks_arg: int, histo_max: int, device: torch.device,
                 group=None, stream=None, exchange: str = "nccl", arena_entries: int = 0):
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
        self.exchange = exchange
        if exchange == "p2p":
            self._setup_p2p(arena_entries)

    def _setup_p2p(self, arena_entries: int):
        """Allocate this rank's receive arenas and map every peer's through CUDA IPC."""
        e = self.e
        e.p2p_arena_create(int(arena_entries))
        mine = [e.p2p_arena_handle(slot) for slot in range(2)]
        allh = [None] * self.world
        dist.all_gather_object(allh, mine, group=self.group)
        for r in range(self.world):
            if r != self.rank:
                for slot in range(2):
                    e.p2p_open_peer(r, slot, allh[r][slot])
        self._tick = torch.zeros(1, dtype=torch.int32, device=self.device)

    # -- small helpers ---------------------------------------------------------
    def _exchange_counts(self, counts: np.ndarray) -> np.ndarray:
        """counts[d, r] = k-mers this rank sends to rank d for d's table region r.
        Returns recv[s, r] = k-mers rank s sends to this rank for our region r."""
        send = torch.as_tensor(np.ascontiguousarray(counts).astype(np.int64), device=self.device)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)  # row d goes to rank d
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
        if self.exchange == "p2p":
            return self._finalize_p2p()
        return self._finalize_nccl()

    def _finalize_p2p(self):
        e = self.e
        e.finalize_external()
        for c in range(self.n_chunks):
            slot = c & 1
            counts = e.route_count(c, self.world)            # (world, regions)
            rcounts = self._exchange_counts(counts)          # (world, regions): per source and region
            per_dst = counts.sum(axis=1).astype(np.int64)
            # where this rank's block starts in every destination's arena: after the blocks of lower ranks
            allt = torch.empty(self.world * self.world, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(allt, torch.as_tensor(per_dst, device=self.device), group=self.group)
            m = allt.cpu().numpy().reshape(self.world, self.world)   # m[s, d] = k-mers s sends to d
            dst_offsets = m[:self.rank, :].sum(axis=0)
            need = int(m[:, self.rank].sum())
            e.route_scatter_p2p(c, slot, dst_offsets)        # fused: extract + bucket + peer stores
            dist.all_reduce(self._tick, group=self.group)    # stream-ordered barrier: every rank's stores landed
            e.insert_runs_device(e.p2p_arena_ptr(slot), rcounts)
            self.bytes_sent += 8 * (int(per_dst.sum()) - int(per_dst[self.rank]))
            self.kmers_received += need
            if self.chunks_arg > 0:
                e.snapshot_histogram(c)
            else:
                e.sync()
        if self.chunks_arg == 0:
            return None
        cols = np.stack([e.histogram(c) for c in range(self.n_chunks)]).astype(np.int64)
        t = torch.as_tensor(cols, device=self.device)
        dist.all_reduce(t, group=self.group)
        return t.cpu().numpy().astype(np.uint64)

    def _finalize_nccl(self):
        e = self.e
        e.finalize_external()  # ingest is complete on every rank; the chunk loop is ours

        def launch_exchange(c):
            counts = e.route_count(c, self.world)            # (world, regions): per destination and region
            rcounts = self._exchange_counts(counts)          # (world, regions): per source and region
            per_dst, per_src = counts.sum(axis=1), rcounts.sum(axis=1)
            n_send, n_recv = int(per_dst.sum()), int(per_src.sum())
            send, recv = self._alloc(n_send), self._alloc(n_recv)
            e.route_scatter(c, self._ptr(send))              # bucket order = destination-major (engine stream)
            work = dist.all_to_all_single(recv[:n_recv], send[:n_send],
                                          output_split_sizes=[int(x) for x in per_src],
                                          input_split_sizes=[int(x) for x in per_dst],
                                          group=self.group, async_op=True)
            self.bytes_sent += 8 * (n_send - int(per_dst[self.rank]))
            self.kmers_received += n_recv
            return work, send, recv, rcounts, c

        pending = launch_exchange(0)
        for c in range(self.n_chunks):
            work, send, recv, rcounts, cc = pending
            # start routing the next chunk while this chunk's k-mers are in flight
            nxt = launch_exchange(c + 1) if c + 1 < self.n_chunks else None
            work.wait()                                       # engine stream waits for the collective
            e.insert_runs_device(self._ptr(recv), rcounts)   # region-major over all sources' runs
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
