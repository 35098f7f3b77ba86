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
  * "dma": batches are bucketed at ingest time; per chunk the copy engines push each
    destination's block into its arena over NVLink while a CPU (gloo) group carries the counts
    and the two barriers: the exchange needs no SM at all and hides behind the inserts.
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


def bind_to_gpu_numa(device_index: int) -> list[int] | None:
    """Pin the calling thread (and the threads it spawns later) to the CPUs NVML reports as local to
    GPU `device_index`, so that pinned staging buffers allocated afterwards land on that GPU's NUMA
    node. With one process per GPU and no binding, half the ranks of an 8-GPU box read their FASTQ
    bytes across the socket interconnect and host->device bandwidth drops by 2x (profiles/
    experiments_r01.md #24). Call before the first pinned allocation; returns the CPU list, or None
    when NVML is unavailable (binding is an optimisation, never a requirement)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        if not uuid.startswith("GPU-"):
            uuid = "GPU-" + uuid
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except TypeError:
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


class _CudaArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}


def _device_i64(ptr: int, n: int, device) -> torch.Tensor:
    """Zero-copy int64 view of n u64 values at a raw device pointer."""
    return torch.as_tensor(_CudaArray(ptr, n), device=device)


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
        if exchange in ("p2p", "dma"):
            self._setup_p2p(arena_entries)

    def _setup_p2p(self, arena_entries: int):
        """Allocate this rank's receive arenas and map every peer's through CUDA IPC."""
        e = self.e
        # "dma": one arena per chunk (up to 16), so that every chunk can be exchanged ahead of its
        # insert while the chip is busy; "p2p": two (double buffering)
        self._n_slots = min(16, self.n_chunks) if self.exchange == "dma" else 2
        e.p2p_arena_create(int(arena_entries), self._n_slots)
        mine = [e.p2p_arena_handle(slot) for slot in range(self._n_slots)]
        allh = [None] * self.world
        dist.all_gather_object(allh, mine, group=self.group)
        for r in range(self.world):
            if r != self.rank:
                for slot in range(self._n_slots):
                    e.p2p_open_peer(r, slot, allh[r][slot])
        self._tick = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._regions = e.route_regions()
        # control plane of the "dma" exchange: a CPU group, so that metadata and barriers never
        # need an SM (created collectively by every rank)
        self._cpu_group = dist.new_group(backend="gloo") if self.exchange == "dma" else None

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
        # Issued with the engine's main stream current: tensors are allocated on it and the final
        # all-reduce is ordered after the inserts.
        with self._on_stream():
            out = self._finalize()
        self.e.sync()  # waits for every stream; raises if a pack kernel met an invalid base
        return out

    def _finalize(self):
        if self.exchange == "dma":
            return self._finalize_dma()
        if self.exchange == "p2p":
            return self._finalize_p2p()
        return self._finalize_nccl()

    def _streams(self):
        """(main, routing) torch streams wrapping the engine's CUDA streams; None on CPU."""
        if self.stream is None:
            return None, None
        main = torch.cuda.ExternalStream(self.e.stream_handle(0), device=self.device)
        part = torch.cuda.ExternalStream(self.e.stream_handle(1), device=self.device)
        return main, part

    def _finalize_p2p(self):
        """Software pipeline: chunk c+1 is routed (extract + bucket + peer stores + barrier, routing
        stream) while chunk c is inserted (main stream)."""
        e = self.e
        e.finalize_external()
        main, part = self._streams()

        def route(c, prev_insert_done):
            slot = c & 1
            with torch.cuda.stream(part):
                # per-(destination, region) counts of this rank's k-mers, left on the device and
                # all-gathered from there: one collective and one host sync per chunk
                d_counts = e.route_count_device(c)
                mine = _device_i64(d_counts, self.world * self._regions, self.device)
                allt = torch.empty(self.world * self.world * self._regions, dtype=torch.int64, device=self.device)
                dist.all_gather_into_tensor(allt, mine, group=self.group)
                allc = allt.cpu().numpy().reshape(self.world, self.world, self._regions)  # [src, dst, region]
                counts = allc[self.rank].astype(np.uint64)
                rcounts = np.ascontiguousarray(allc[:, self.rank, :]).astype(np.uint64)
                m = allc.sum(axis=2)                              # m[s, d] = k-mers s sends to d
                per_dst = m[self.rank]
                e.route_set_counts(c, counts)
                # this rank's block in every destination's arena starts after the lower ranks' blocks
                if self.exchange == "dma":   # bucket locally, copy engines push the runs to the peers
                    e.route_scatter_dma(c, slot, m[:self.rank, :].sum(axis=0))
                else:                          # fused: extract + bucket + peer stores in one kernel
                    e.route_scatter_p2p(c, slot, m[:self.rank, :].sum(axis=0))
                # Barrier c tells the peers two things: (1) my stores of chunk c have landed, and
                # (2) my insert of chunk c-1 is finished, so after this barrier they may overwrite
                # the other arena slot (their scatter of chunk c+1).
                if prev_insert_done is not None:
                    part.wait_event(prev_insert_done)
                dist.all_reduce(self._tick, group=self.group)    # stream-ordered barrier
                done = torch.cuda.Event()
                done.record(part)
            self.bytes_sent += 8 * (int(per_dst.sum()) - int(per_dst[self.rank]))
            self.kmers_received += int(m[:, self.rank].sum())
            return slot, rcounts, done

        nxt = route(0, None)
        for c in range(self.n_chunks):
            slot, rcounts, done = nxt
            main.wait_event(done)
            e.insert_runs_device(e.p2p_arena_ptr(slot), rcounts)     # asynchronous, main stream
            inserted = torch.cuda.Event()
            inserted.record(main)
            if c + 1 < self.n_chunks:
                nxt = route(c + 1, inserted)                      # overlaps the insert just queued
            if self.chunks_arg > 0:
                e.snapshot_histogram(c)                          # syncs the main stream
            else:
                e.sync()
        if self.chunks_arg == 0:
            return None
        cols = np.stack([e.histogram(c) for c in range(self.n_chunks)]).astype(np.int64)
        t = torch.as_tensor(cols, device=self.device)
        dist.all_reduce(t, group=self.group)
        return t.cpu().numpy().astype(np.uint64)

    def _finalize_dma(self):
        """Exchange with zero SM time: batches were bucketed by (owner, region) at ingest time; per
        chunk the HOST gathers the counts and runs the two barriers over a CPU (gloo) group, and
        the copy engines push every destination's block into its arena over NVLink.  Nothing here
        needs a free SM, so it all proceeds while the persistent insert kernel of the previous
        chunk owns the chip (a small NCCL kernel would have to wait for it to finish).

            host, chunk c:  wait own insert(c-2) -> all-gather counts (= slot free everywhere)
                            -> peer copies(c) -> wait for them -> BARRIER (all blocks landed)
                            -> queue insert(c), queue histogram column c
        """
        e = self.e
        e.finalize_external()
        main, _ = self._streams()
        cpu = self._cpu_group
        regions = self._regions
        n_slots = self._n_slots
        inserted, pending = {}, {}

        def exchange(c):
            """Queue the peer copies of chunk c into arena slot c % n_slots (asynchronous)."""
            slot = c % n_slots
            counts = e.route_count(c, self.world)                 # host: waits for chunk c's bucketing only
            if c >= n_slots:
                inserted[c - n_slots].synchronize()               # my reads of this arena slot are over ...
            mine = torch.from_numpy(counts.astype(np.int64).reshape(-1))
            allt = torch.empty(self.world * mine.numel(), dtype=torch.int64)
            dist.all_gather_into_tensor(allt, mine, group=cpu)    # ... and, once this returns, everybody's
            allc = allt.numpy().reshape(self.world, self.world, regions)   # [src, dst, region]
            m = allc.sum(axis=2)                                  # m[s, d] = k-mers s sends to d
            e.route_scatter_dma(c, slot, m[:self.rank, :].sum(axis=0))
            pending[c] = np.ascontiguousarray(allc[:, self.rank, :]).astype(np.uint64)
            self.bytes_sent += 8 * (int(m[self.rank].sum()) - int(m[self.rank, self.rank]))
            self.kmers_received += int(m[:, self.rank].sum())

        nxt = 0
        for c in range(self.n_chunks):
            # keep the exchange as far ahead of the inserts as there are arena slots: the copies then
            # run while the chip is bucketing (SM-bound) instead of inserting (memory-bound), which
            # slows the copy engines down 3-4x
            # ... but only as far as every rank has its batches on the device already: a chunk that
            # is still arriving from the host must not hold up the inserts of the earlier ones
            mine = torch.tensor([e.chunks_ready()], dtype=torch.int64)
            allr = torch.empty(self.world, dtype=torch.int64)
            dist.all_gather_into_tensor(allr, mine, group=cpu)
            ahead = max(int(allr.min()), c + 1)
            while nxt < self.n_chunks and nxt < min(ahead, c + n_slots):
                exchange(nxt)
                nxt += 1
            e.dma_wait(c % n_slots)                               # my blocks of chunk c have landed at the peers
            dist.barrier(group=cpu)                               # ... and everybody's at mine
            e.insert_runs_device(e.p2p_arena_ptr(c % n_slots), pending.pop(c))  # asynchronous, main stream
            ev = torch.cuda.Event()
            ev.record(main)
            inserted[c] = ev
            if self.chunks_arg > 0:
                e.snapshot_histogram_async(c)
        if self.chunks_arg == 0:
            e.sync()
            return None
        cols = np.stack([e.histogram(c) for c in range(self.n_chunks)]).astype(np.int64)
        t = torch.as_tensor(cols, device=self.device)
        dist.all_reduce(t, group=self.group)
        return t.cpu().numpy().astype(np.uint64)

    def _finalize_nccl(self):
        """Route to a local list, all_to_all_single, insert.  On GPUs the routing + collective of
        chunk c+1 run on the routing stream while chunk c is inserted on the main stream."""
        import contextlib
        e = self.e
        e.finalize_external()  # ingest is complete on every rank; the chunk loop is ours
        main, part = self._streams()
        on_part = (lambda: torch.cuda.stream(part)) if part is not None else contextlib.nullcontext

        def route(c):
            with on_part():
                counts = e.route_count(c, self.world)            # (world, regions): per destination and region
                rcounts = self._exchange_counts(counts)          # (world, regions): per source and region
                per_dst, per_src = counts.sum(axis=1), rcounts.sum(axis=1)
                n_send, n_recv = int(per_dst.sum()), int(per_src.sum())
                send, recv = self._alloc(n_send), self._alloc(n_recv)
                e.route_scatter(c, self._ptr(send))              # bucket order = destination-major
                dist.all_to_all_single(recv[:n_recv], send[:n_send],
                                       output_split_sizes=[int(x) for x in per_src],
                                       input_split_sizes=[int(x) for x in per_dst], group=self.group)
                done = None
                if part is not None:
                    done = torch.cuda.Event()
                    done.record(part)
                    recv.record_stream(main)                     # consumed by the insert on the main stream
            self.bytes_sent += 8 * (n_send - int(per_dst[self.rank]))
            self.kmers_received += n_recv
            return recv, rcounts, done

        nxt = route(0)
        for c in range(self.n_chunks):
            recv, rcounts, done = nxt
            if done is not None:
                main.wait_event(done)
            e.insert_runs_device(self._ptr(recv), rcounts)       # region-major over all sources' runs
            if c + 1 < self.n_chunks:
                nxt = route(c + 1)                                # overlaps the insert just queued
            if self.chunks_arg > 0:
                e.snapshot_histogram(c)                          # this rank's partial column (syncs main)
            else:
                e.sync()
            del recv
        if self.chunks_arg == 0:
            return None
        cols = np.stack([e.histogram(c) for c in range(self.n_chunks)]).astype(np.int64)
        t = torch.as_tensor(cols, device=self.device)
        dist.all_reduce(t, group=self.group)                  # histogram = sum of the partitions' histograms
        return t.cpu().numpy().astype(np.uint64)

    # -- table services over the sharded table (what sPCR needs; collective: every rank calls them
    #    with the same arguments and gets the same answer) ----------------------------------------
    def lookup(self, kmers, min_count: int = 0, mode: int = 0):
        """skm_lookup_batch over all shards: a k-mer lives on exactly one rank (ownership is by the
        hash of the canonical k-mer), so the answer is the maximum over the ranks' local answers."""
        counts, _ = self.e.lookup(kmers, min_count, mode)
        t = torch.as_tensor(np.ascontiguousarray(counts).astype(np.int64), device=self.device)
        if t.numel():
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        c = t.cpu().numpy().astype(np.uint32)
        return c, c > 0

    def scan_oligos(self, oligos, oligo_length: int, min_count: int):
        """skm_scan_oligos over all shards: every rank scans its own part of the table, the matches
        are gathered and merged in ascending k-mer order."""
        keys, counts = self.e.scan_oligos(oligos, oligo_length, min_count)
        n = torch.tensor([keys.size], dtype=torch.int64, device=self.device)
        sizes = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(self.world)]
        dist.all_gather(sizes, n, group=self.group)
        sizes = [int(x.item()) for x in sizes]
        width = max(max(sizes), 1)
        mine = torch.zeros(2 * width, dtype=torch.int64, device=self.device)
        mine[:keys.size] = torch.as_tensor(keys.astype(np.int64), device=self.device)
        mine[width:width + keys.size] = torch.as_tensor(counts.astype(np.int64), device=self.device)
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        all_k = np.concatenate([p[:sizes[r]].cpu().numpy().astype(np.uint64) for r, p in enumerate(parts)])
        all_c = np.concatenate([p[width:width + sizes[r]].cpu().numpy().astype(np.uint32) for r, p in enumerate(parts)])
        order = np.argsort(all_k, kind="stable")
        return all_k[order], all_c[order]

    def global_totals(self, local: dict) -> dict:
        keys = sorted(local)
        t = torch.as_tensor([int(local[k]) for k in keys], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, group=self.group)
        return dict(zip(keys, t.cpu().tolist()))
