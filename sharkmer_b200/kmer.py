"""Host-side mirror of the reference's k-mer API over the C ABI.

Names, argument meaning and error behaviour follow caseywdunn/sharkmer v3.1.0
`src/kmer/mod.rs:10-17` (Chunk, KmerCounts, FilteredKmerCounts, Histogram,
kmers_from_ascii, revcomp_kmer, ...) so that parity tests read like the
reference's own tests.  All counting runs in the CUDA library; this module only
marshals buffers.  (The Rust shim a sharkmer maintainer would add is in
INTEGRATION.md; Rust is not available in this build environment.)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import common
from ._lib import SkmParams, SkmStageMs, SkmTotals

EMPTY = common.EMPTY_KEY


class SkmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


def _u8(buf) -> np.ndarray:
    if isinstance(buf, str):
        buf = buf.encode()
    if isinstance(buf, (bytes, bytearray, memoryview)):
        return np.frombuffer(buf, dtype=np.uint8)
    a = np.asarray(buf)
    if a.dtype != np.uint8:
        raise TypeError("expected bytes or uint8 array")
    return np.ascontiguousarray(a)


class Engine:
    """One skm_ctx: one GPU, one table partition."""

    def __init__(self, k: int, chunks: int = 0, histo_max: int = 10000, capacity_hint: int = 0,
                 device: int = -1, insert_mode: int = _lib.INSERT_AUTO, n_ranks: int = 1, rank: int = 0,
                 stream: int = 0):
        self.L = _lib.load()
        self.k, self.chunks, self.histo_max = k, chunks, histo_max
        self.n_chunks = max(1, chunks)
        p = SkmParams(struct_size=C.sizeof(SkmParams), k=k, chunks=chunks, insert_mode=insert_mode,
                      histo_max=histo_max, capacity_hint=capacity_hint, device=device,
                      n_ranks=n_ranks, rank=rank, reserved=0, stream=stream)
        h = C.c_void_p()
        rc = self.L.skm_create(C.byref(p), C.byref(h))
        self._h = h
        if rc:
            msg = self.L.skm_last_error(h).decode() if h else "skm_create failed"
            if h:
                self.L.skm_destroy(h)
            self._h = None
            raise SkmError(rc, msg)

    def close(self):
        if getattr(self, "_h", None):
            self.L.skm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise SkmError(rc, self.L.skm_last_error(self._h).decode(errors="replace"))

    # ---- ingest ---------------------------------------------------------
    def ingest_batch(self, chunk_index: int, seqs, flags: int = 0):
        """drain_batch (src/io.rs:355-361): newline-terminated sequence lines."""
        a = _u8(seqs)
        self._ck(self.L.skm_ingest_batch(self._h, chunk_index, a.ctypes.data, a.size, flags))

    def ingest_ptr(self, chunk_index: int, ptr: int, n_bytes: int, flags: int = 0):
        self._ck(self.L.skm_ingest_batch(self._h, chunk_index, ptr, n_bytes, flags))

    def ingest_reads(self, chunk_index: int, bases, offsets):
        a = _u8(bases)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._ck(self.L.skm_ingest_reads(self._h, chunk_index, a.ctypes.data, off.ctypes.data, off.size - 1))

    def ingest_device(self, chunk_index: int, d_ptr: int, n_bytes: int):
        self._ck(self.L.skm_ingest_device(self._h, chunk_index, d_ptr, n_bytes))

    def sync(self): self._ck(self.L.skm_sync(self._h))
    def finalize(self): self._ck(self.L.skm_finalize(self._h))
    def reset(self): self._ck(self.L.skm_reset(self._h))

    # ---- results --------------------------------------------------------
    def histogram(self, chunk_i: int) -> np.ndarray:
        out = np.zeros(self.histo_max + 2, dtype=np.uint64)
        self._ck(self.L.skm_histogram(self._h, chunk_i, out.ctypes.data, out.size))
        return out

    def totals(self) -> SkmTotals:
        t = SkmTotals()
        self._ck(self.L.skm_totals_get(self._h, C.byref(t)))
        return t

    def chunk_totals(self, chunk: int) -> SkmTotals:
        t = SkmTotals()
        self._ck(self.L.skm_chunk_totals(self._h, chunk, C.byref(t)))
        return t

    def stage_times(self) -> SkmStageMs:
        t = SkmStageMs()
        self._ck(self.L.skm_stage_times(self._h, C.byref(t)))
        return t

    def table_len(self) -> int:
        n = C.c_uint64()
        self._ck(self.L.skm_table_len(self._h, C.byref(n)))
        return n.value

    def export(self, sorted: bool = True):
        n = self.table_len()
        keys = np.empty(n, dtype=np.uint64)
        counts = np.empty(n, dtype=np.uint32)
        got = C.c_uint64()
        self._ck(self.L.skm_export(self._h, keys.ctypes.data, counts.ctypes.data, n, int(sorted), C.byref(got)))
        return keys[:got.value], counts[:got.value]

    def digest(self) -> int:
        d = C.c_uint64()
        self._ck(self.L.skm_table_digest(self._h, C.byref(d)))
        return d.value

    def lookup(self, kmers, min_count: int = 0, mode: int = _lib.LOOKUP_CANONICAL):
        q = np.ascontiguousarray(kmers, dtype=np.uint64)
        counts = np.zeros(q.size, dtype=np.uint32)
        found = np.zeros(q.size, dtype=np.uint8)
        self._ck(self.L.skm_lookup_batch(self._h, q.ctypes.data, q.size, min_count, mode,
                                         counts.ctypes.data, found.ctypes.data))
        return counts, found.astype(bool)

    def scan_oligos(self, oligos, oligo_length: int, min_count: int, cap: int = 1 << 16):
        """find_oligos_in_kmers (src/pcr/primers.rs:163-226): sorted (kmers, counts).  One table pass
        (the matches of a primer are a handful of k-mers); a second one only if `cap` was too small."""
        o = np.ascontiguousarray(oligos, dtype=np.uint64)
        n = C.c_uint64()
        while True:
            keys = np.empty(cap, dtype=np.uint64)
            counts = np.empty(cap, dtype=np.uint32)
            rc = self.L.skm_scan_oligos(self._h, o.ctypes.data, o.size, oligo_length, min_count,
                                        keys.ctypes.data, counts.ctypes.data, cap, C.byref(n))
            if rc == _lib.ERR_INVALID_ARG and n.value > cap:
                cap = int(n.value)
                continue
            self._ck(rc)
            return keys[:n.value].copy(), counts[:n.value].copy()

    def insert_counts(self, keys, counts):
        k = np.ascontiguousarray(keys, dtype=np.uint64)
        c = np.ascontiguousarray(counts, dtype=np.uint32)
        assert k.size == c.size
        self._ck(self.L.skm_insert_counts(self._h, k.ctypes.data, c.ctypes.data, k.size))

    # ---- multi-GPU (include/sharkmer_b200.h: skm_mg_*) ---------------------
    def mg_arena_create(self, n_bytes: int):
        self._ck(self.L.skm_mg_arena_create(self._h, int(n_bytes)))

    def mg_arena_handle(self) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._ck(self.L.skm_mg_arena_handle(self._h, buf))
        return bytes(buf)

    def mg_arena_ptr(self) -> int:
        p = C.c_void_p()
        self._ck(self.L.skm_mg_arena_ptr(self._h, C.byref(p)))
        return p.value or 0

    def mg_open_peer(self, peer_rank: int, handle: bytes):
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        self._ck(self.L.skm_mg_open_peer(self._h, peer_rank, buf))

    def mg_set_peer(self, peer_rank: int, d_ptr: int, peer_device: int = -1):
        self._ck(self.L.skm_mg_set_peer(self._h, peer_rank, d_ptr, peer_device))

    def mg_finalize(self, comm_struct):
        """Collective.  comm_struct: a ctypes skm_comm (sharkmer_b200.multigpu builds one over torch.distributed)."""
        self._ck(self.L.skm_mg_finalize(self._h, C.byref(comm_struct)))

    def mg_flush(self, comm_struct):
        """Collective, chunks == 0: count what was ingested so far, free lists and arenas."""
        self._ck(self.L.skm_mg_flush(self._h, C.byref(comm_struct)))

    def mg_bytes_sent(self) -> int:
        n = C.c_uint64()
        self._ck(self.L.skm_mg_bytes_sent(self._h, C.byref(n)))
        return n.value

    def insert_kmers_device(self, d_ptr: int, n: int):
        self._ck(self.L.skm_insert_kmers_device(self._h, d_ptr, n))

    def snapshot_histogram(self, chunk_i: int):
        self._ck(self.L.skm_snapshot_histogram(self._h, chunk_i))

    # ---- diagnostics -----------------------------------------------------
    def extract_kmers(self, seqs) -> np.ndarray:
        a = _u8(seqs)
        out = np.empty(a.size, dtype=np.uint64)
        self._ck(self.L.skm_extract_kmers(self._h, a.ctypes.data, a.size, out.ctypes.data))
        return out

    def pack(self, seqs):
        a = _u8(seqs)
        n_units = (a.size + 31) // 32
        codes = np.zeros(n_units, dtype=np.uint64)
        breaks = np.zeros(n_units, dtype=np.uint32)
        self._ck(self.L.skm_pack(self._h, a.ctypes.data, a.size, codes.ctypes.data, breaks.ctypes.data))
        return codes, breaks

    def device_alloc(self, n_bytes: int) -> int:
        p = C.c_void_p()
        self._ck(self.L.skm_device_alloc(self._h, n_bytes, C.byref(p)))
        return p.value

    def device_free(self, ptr: int): self._ck(self.L.skm_device_free(self._h, ptr))

    def pinned_alloc(self, n_bytes: int) -> int:
        p = C.c_void_p()
        self._ck(self.L.skm_pinned_alloc(self._h, n_bytes, C.byref(p)))
        return p.value

    def pinned_free(self, ptr: int): self._ck(self.L.skm_pinned_free(self._h, ptr))

    def memcpy_d2h(self, dst: np.ndarray, d_src: int, n_bytes: int):
        self._ck(self.L.skm_memcpy_d2h(self._h, dst.ctypes.data, d_src, n_bytes))

    def memcpy_d2h_ptr(self, dst_ptr: int, d_src: int, n_bytes: int):
        self._ck(self.L.skm_memcpy_d2h(self._h, dst_ptr, d_src, n_bytes))

    def synth_device(self, seed, genome_len, read_len, sub_thresh, n_thresh, chunk_index, n_chunks, first, n, d_out):
        self._ck(self.L.skm_synth_device(self._h, seed, genome_len, read_len, sub_thresh, n_thresh,
                                         chunk_index, n_chunks, first, n, d_out))

    def bench_gups(self, log2_slots: int, n_updates: int, iters: int = 3, variant: int = 0) -> float:
        ms = C.c_float()
        self._ck(self.L.skm_bench_gups(self._h, log2_slots, n_updates, iters, variant, C.byref(ms)))
        return ms.value


# ---------------------------------------------------------------------------
# Reference-shaped API (src/kmer/mod.rs:10-17)
# ---------------------------------------------------------------------------

def kmers_from_ascii(seq, k: int, engine: Engine | None = None) -> list[int]:
    """src/kmer/encoding.rs:332-371, computed by the device extract kernel."""
    if not (0 < k < 32):
        raise SkmError(_lib.ERR_INVALID_ARG, f"k must be between 1 and 31, got {k}")
    e = engine or Engine(k if k % 2 else k)  # the ctx enforces odd k (src/cli.rs:665)
    s = seq.encode() if isinstance(seq, str) else bytes(seq)
    out = e.extract_kmers(s + b"\n")
    return [int(v) for v in out if int(v) != EMPTY]


def revcomp_kmer(kmer: int, k: int) -> int:
    """src/kmer/encoding.rs:235-262 (host helper; the device uses the same skm_common.h function)."""
    return common.revcomp_kmer(kmer, k)


def kmer_to_seq(kmer: int, k: int) -> str:
    """src/kmer/encoding.rs:311-325."""
    return "".join("ACGT"[(kmer >> (2 * (k - i - 1))) & 3] for i in range(k))


def kmer_last_base(kmer: int) -> str:
    """src/kmer/encoding.rs:301-309."""
    return "ACGT"[kmer & 3]


def seq_to_kmer(seq: str) -> int:
    """src/kmer/encoding.rs:379-392."""
    v = 0
    for ch in seq:
        if ch not in "ACGT":
            raise SkmError(_lib.ERR_INVALID_BASE, f"Invalid base '{ch}' in sequence '{seq}'")
        v = (v << 2) | "ACGT".index(ch)
    return v


def count_valid_bases(seq: str) -> int:
    """src/kmer/encoding.rs:374-376."""
    return sum(1 for ch in seq if ch != "N")


class KmerCounts:
    """src/kmer/counting.rs:113-312 over a device table."""

    def __init__(self, k: int, capacity: int = 0, _engine: Engine | None = None):
        self._e = _engine or Engine(k, chunks=0, capacity_hint=capacity)
        self.k = k
        self._dirty = False  # staged reads not yet counted

    @classmethod
    def new_with_capacity(cls, k: int, capacity: int):
        return cls(k, capacity)

    def get_k(self): return self.k

    def ingest_seq(self, seq):
        """counting.rs:144-149.  An invalid base fails the call (encoding.rs:353-356)."""
        s = seq.encode() if isinstance(seq, str) else bytes(seq)
        for ch in s:
            if ch not in b"ACGTN":
                raise SkmError(_lib.ERR_INVALID_BASE,
                               f"Invalid character '{chr(ch)}' in sequence. Only ACGTN allowed.")
        kmers = self._e.extract_kmers(s + b"\n")
        kmers = kmers[kmers != np.uint64(EMPTY)]
        if kmers.size:
            u, c = np.unique(kmers, return_counts=True)
            self._e.insert_counts(u, c.astype(np.uint32))

    def insert(self, kmer: int, count: int):
        """counting.rs:152-154 (saturating)."""
        self._e.insert_counts([kmer], [count])

    def extend(self, other: "KmerCounts"):
        """counting.rs:157-166."""
        if self.k != other.k:
            raise SkmError(_lib.ERR_K_MISMATCH, "Cannot extend KmerCounts with different k")
        keys, counts = other._e.export(sorted=False)
        self._e.insert_counts(keys, counts)

    def get(self, kmer: int):
        c, f = self._e.lookup([kmer], 0, _lib.LOOKUP_EXACT)
        return int(c[0]) if f[0] else None

    def get_count(self, kmer: int) -> int: return self.get(kmer) or 0
    def contains(self, kmer: int) -> bool: return self.get(kmer) is not None

    def get_canonical_count(self, kmer: int) -> int:
        """counting.rs:205-209."""
        c, _ = self._e.lookup([kmer], 0, _lib.LOOKUP_CANONICAL)
        return int(c[0])

    def get_canonical(self, kmer: int):
        """counting.rs:218-222: probe the k-mer, else its reverse complement."""
        c, f = self._e.lookup([kmer], 0, _lib.LOOKUP_EITHER)
        return int(c[0]) if f[0] else None

    def len(self) -> int: return self._e.table_len()
    __len__ = len
    def is_empty(self) -> bool: return self.len() == 0
    def get_n_kmers(self) -> int: return int(self._e.totals().n_kmers)
    def get_n_unique_kmers(self) -> int: return self.len()
    def iter(self):
        keys, counts = self._e.export(sorted=False)
        return zip(keys.tolist(), counts.tolist())
    def kmers(self): return self._e.export(sorted=False)[0].tolist()
    def counts(self): return self._e.export(sorted=False)[1].tolist()
    def export_sorted(self): return self._e.export(sorted=True)
    def digest(self) -> int: return self._e.digest()

    def get_max_count(self) -> int:
        c = self._e.export(sorted=False)[1]
        return int(c.max()) if c.size else 0

    def get_median_count(self) -> int:
        """counting.rs:279-300: even length -> lower/2 + upper/2."""
        c = np.sort(self._e.export(sorted=False)[1])
        n = c.size
        if n == 0:
            return 0
        return int(c[n // 2]) if n % 2 else int(c[n // 2 - 1]) // 2 + int(c[n // 2]) // 2

    def filtered_view(self, min_count: int) -> "FilteredKmerCounts":
        return FilteredKmerCounts(self, min_count)


class FilteredKmerCounts:
    """src/kmer/counting.rs:316-350: lazy `count >= min_count` view."""

    def __init__(self, inner: KmerCounts, min_count: int):
        self.inner, self.min_count = inner, min_count

    def get_k(self): return self.inner.k

    def get_canonical(self, kmer: int):
        c, f = self.inner._e.lookup([kmer], self.min_count, _lib.LOOKUP_EITHER)
        return int(c[0]) if f[0] else None

    def get_canonical_count(self, kmer: int) -> int:
        c, _ = self.inner._e.lookup([kmer], self.min_count, _lib.LOOKUP_CANONICAL)
        return int(c[0])

    def get_canonical_counts(self, kmers) -> np.ndarray:
        """Batched form (what src/pcr/graph.rs:419-430 would call once per frontier)."""
        return self.inner._e.lookup(kmers, self.min_count, _lib.LOOKUP_CANONICAL)[0]

    def iter(self): return self.inner.iter()


class Histogram:
    """src/kmer/histogram.rs: the vector form (get_vector, :125-134)."""

    def __init__(self, vec: np.ndarray, histo_max: int):
        self.vec, self.histo_max = vec, histo_max

    @classmethod
    def from_kmer_counts(cls, kc: KmerCounts, histo_max: int):
        e = Engine(kc.k, chunks=1, histo_max=histo_max)
        keys, counts = kc._e.export(sorted=False)
        e.insert_counts(keys, counts)
        e.snapshot_histogram(0)
        return cls(e.histogram(0), histo_max)

    def get_vector(self): return self.vec
    def get_n_unique_kmers(self): return int(self.vec[1:].sum())


class Chunk:
    """src/kmer/chunk.rs: a stripe of reads routed to one chunk index of an Engine."""

    def __init__(self, engine: Engine, index: int):
        self._e, self.index = engine, index

    def ingest_seq(self, seq):
        s = seq.encode() if isinstance(seq, str) else bytes(seq)
        self._e.ingest_batch(self.index, s + b"\n")

    def get_n_reads(self): return int(self._e.chunk_totals(self.index).n_reads)
    def get_n_bases(self): return int(self._e.chunk_totals(self.index).n_bases)
    def get_n_kmers(self): return int(self._e.chunk_totals(self.index).n_kmers)
