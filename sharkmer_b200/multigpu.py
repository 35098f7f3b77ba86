"""Hash-sharded k-mer counting over N GPUs.

The multi-GPU path lives behind the C ABI (include/sharkmer_b200.h, skm_mg_* / skm_group_*):
every rank buckets and tile-sorts its batches as on one GPU, the copy engines push the other
owners' slices into their receive arenas over NVLink at ingest time, and the collective
skm_mg_finalize counts everything in chunk order and sums the histogram columns.  The library
only borrows ONE primitive from its host: an all-gather of host bytes among the ranks.

This module supplies that primitive for the two ways of running N ranks:
  * ShardedCounter — one process per GPU (torchrun): the all-gather is torch.distributed on a
    CPU (gloo) group; arenas are wired through CUDA IPC handles.
  * Group          — all ranks in one process (skm_group_*): threads + an in-process all-gather
    inside the library; arenas are wired by device pointers.  Several ranks may share a GPU,
    which is how the multi-rank path is tested on a one-GPU box.
Both also offer the table services sPCR needs over the sharded table (lookup, scan_oligos).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def bind_to_gpu_numa(device_index: int) -> list[int] | None:
    """Pin the calling thread (and the threads it spawns later) to the CPUs NVML reports as local to
    GPU `device_index`, so that pinned staging buffers allocated afterwards land on that GPU's NUMA
    node.  Call before the first pinned allocation; returns the CPU list, or None when NVML is
    unavailable (binding is an optimisation, never a requirement)."""
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


# ---- the all-gather the library borrows (skm_comm) ------------------------------------------------

_ALLGATHER = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)


class SkmComm(C.Structure):
    _fields_ = [("user", C.c_void_p), ("allgather", _ALLGATHER)]


class TorchComm:
    """skm_comm over a torch.distributed group of CPU tensors (gloo)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.calls = 0
        self.bytes = 0
        self.error = None

        def _allgather(_user, send, recv, n_bytes):
            try:
                n = int(n_bytes)
                src = np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_uint8)), shape=(n,))
                dst = np.ctypeslib.as_array(C.cast(recv, C.POINTER(C.c_uint8)), shape=(n * self.world,))
                dist.all_gather_into_tensor(torch.from_numpy(dst), torch.from_numpy(src.copy()), group=self.group)
                self.calls += 1
                self.bytes += n
                return 0
            except Exception as e:  # never let an exception cross the C boundary
                self.error = e
                return 1

        self._cb = _ALLGATHER(_allgather)       # keep the callback object alive
        self.struct = SkmComm(None, self._cb)


class ShardedCounter:
    """One rank of a hash-sharded count (one process per GPU).  `engine` is this rank's
    sharkmer_b200.kmer.Engine, created with n_ranks / rank; ingest into it as usual, then call
    finalize() on every rank."""

    def __init__(self, engine, device: torch.device | None = None, group=None, cpu_group=None, arena_bytes: int = 0):
        self.e = engine
        self.device = device
        self.group = group                      # for the table services (device tensors: NCCL on GPUs)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        # the library's all-gather runs on a CPU group: metadata never needs an SM
        if cpu_group is None:
            cpu_group = group if dist.get_backend(group) == "gloo" else dist.new_group(backend="gloo")
        self.comm = TorchComm(cpu_group)
        if arena_bytes and self.world > 1:
            self._wire(arena_bytes, cpu_group)

    def _wire(self, arena_bytes: int, cpu_group):
        """Allocate this rank's receive arena and map every peer's through CUDA IPC."""
        e = self.e
        e.mg_arena_create(int(arena_bytes))
        handles = [None] * self.world
        dist.all_gather_object(handles, e.mg_arena_handle(), group=cpu_group)
        for r in range(self.world):
            if r != self.rank:
                e.mg_open_peer(r, handles[r])

    def finalize(self) -> np.ndarray | None:
        """Collective: counts every chunk in order.  Returns the (n_chunks, histo_max+2) cumulative
        histogram columns summed over ranks (None when chunks == 0)."""
        e = self.e
        try:
            e.mg_finalize(self.comm.struct)
        except Exception:
            if self.comm.error is not None:
                raise self.comm.error
            raise
        if e.chunks == 0:
            return None
        return np.stack([e.histogram(c) for c in range(e.n_chunks)])

    def flush(self):
        """Collective (chunks == 0): count what has been ingested so far and free the lists and arenas —
        for inputs that do not fit the GPUs' memory in one piece."""
        self.e.mg_flush(self.comm.struct)

    @property
    def bytes_sent(self) -> int:
        return self.e.mg_bytes_sent()

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


class Group:
    """All ranks in one process (skm_group_*): n ctx's, possibly several per GPU."""

    def __init__(self, k: int, chunks: int, histo_max: int, devices, arena_bytes_per_rank: int,
                 capacity_hint: int = 0, insert_mode: int = _lib.INSERT_AUTO):
        from .kmer import Engine, SkmError
        self.L = _lib.load()
        self.n = len(devices)
        p = _lib.SkmParams(struct_size=C.sizeof(_lib.SkmParams), k=k, chunks=chunks, insert_mode=insert_mode,
                           histo_max=histo_max, capacity_hint=capacity_hint, device=-1, n_ranks=self.n, rank=0,
                           reserved=0, stream=0)
        devs = (C.c_int32 * self.n)(*devices)
        h = C.c_void_p()
        rc = self.L.skm_group_create(C.byref(p), self.n, devs, int(arena_bytes_per_rank), C.byref(h))
        self._h = h
        if rc:
            msg = self.L.skm_group_last_error(h).decode() if h else "skm_group_create failed"
            if h:
                self.L.skm_group_destroy(h)
            self._h = None
            raise SkmError(rc, msg)
        # borrowed engines over the members' ctx's (the group owns them)
        self.engines = []
        for r in range(self.n):
            e = Engine.__new__(Engine)
            e.L, e.k, e.chunks, e.histo_max, e.n_chunks = self.L, k, chunks, histo_max, max(1, chunks)
            e._h = C.c_void_p(self.L.skm_group_ctx(h, r))
            e.close = lambda: None  # owned by the group
            self.engines.append(e)

    def _ck(self, rc):
        from .kmer import SkmError
        if rc:
            raise SkmError(rc, self.L.skm_group_last_error(self._h).decode(errors="replace"))

    def finalize(self):
        self._ck(self.L.skm_group_finalize(self._h))

    def flush(self):
        self._ck(self.L.skm_group_flush(self._h))

    def reset(self):
        self._ck(self.L.skm_group_reset(self._h))

    def histogram(self, chunk_i: int) -> np.ndarray:
        return self.engines[0].histogram(chunk_i)     # every member holds the global columns

    def export_sorted(self):
        parts = [e.export(sorted=False) for e in self.engines]
        keys = np.concatenate([p[0] for p in parts])
        counts = np.concatenate([p[1] for p in parts])
        order = np.argsort(keys, kind="stable")
        return keys[order], counts[order]

    def lookup(self, kmers, min_count: int = 0, mode: int = 0):
        """A k-mer lives in exactly one partition: the answer is the maximum over the members."""
        counts = None
        for e in self.engines:
            c, _ = e.lookup(kmers, min_count, mode)
            counts = c if counts is None else np.maximum(counts, c)
        return counts, counts > 0

    def scan_oligos(self, oligos, oligo_length: int, min_count: int):
        parts = [e.scan_oligos(oligos, oligo_length, min_count) for e in self.engines]
        keys = np.concatenate([p[0] for p in parts])
        counts = np.concatenate([p[1] for p in parts])
        order = np.argsort(keys, kind="stable")
        return keys[order], counts[order]

    def close(self):
        if getattr(self, "_h", None):
            for e in self.engines:
                e._h = None
            self.L.skm_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
