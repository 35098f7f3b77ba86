"""ctypes binding of the C oracle + a brute-force pure-Python second opinion.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by sharkmer_b200/.

The C library restates caseywdunn/sharkmer v3.1.0 src/kmer/* and the
ingest/consolidate halves of src/io.rs (citations in oracle/skm_oracle.c).
The pure-Python functions at the bottom are an independent brute-force
statement (sort | uniq -c) used to cross-check the C oracle on small inputs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from collections import Counter

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so (gcc; seconds)."""
    src = [os.path.join(_HERE, f) for f in ("skm_oracle.c", "skm_oracle.h", "oracle_cli.c", "Makefile")]
    src.append(os.path.join(_HERE, "..", "include", "skm_common.h"))
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src if os.path.exists(s))
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    u64, u32, i64, vp, cp = C.c_uint64, C.c_uint32, C.c_int64, C.c_void_p, C.c_char_p
    sz = C.c_size_t
    sig = {
        "orc_kmers_from_ascii": (i64, [cp, sz, u32, vp]),
        "orc_count_valid_bases": (u64, [cp, sz]),
        "orc_revcomp_kmer": (u64, [u64, u32]),
        "orc_seq_to_kmer": (u64, [cp, sz, C.POINTER(C.c_int)]),
        "orc_kmer_to_seq": (None, [u64, u32, cp]),
        "orc_read_pack": (i64, [cp, sz, vp]),
        "orc_read_get_kmers": (i64, [vp, sz, sz, u32, vp]),
        "orc_kmers_via_reads": (i64, [cp, sz, u32, vp]),
        "orc_counts_new": (vp, [u32]),
        "orc_counts_with_capacity": (vp, [u32, u64]),
        "orc_counts_free": (None, [vp]),
        "orc_counts_k": (u32, [vp]),
        "orc_counts_ingest_seq": (C.c_int, [vp, cp, sz]),
        "orc_counts_insert": (None, [vp, u64, u32]),
        "orc_counts_insert_get": (None, [vp, u64, u32, C.POINTER(u32), C.POINTER(u32)]),
        "orc_counts_extend": (C.c_int, [vp, vp]),
        "orc_counts_get": (C.c_int, [vp, u64, C.POINTER(u32)]),
        "orc_counts_get_canonical_count": (u32, [vp, u64]),
        "orc_counts_get_canonical": (C.c_int, [vp, u64, C.POINTER(u32)]),
        "orc_filtered_get_canonical": (C.c_int, [vp, u32, u64, C.POINTER(u32)]),
        "orc_filtered_get_canonical_count": (u32, [vp, u32, u64]),
        "orc_counts_len": (u64, [vp]),
        "orc_counts_n_kmers": (u64, [vp]),
        "orc_counts_max_count": (u32, [vp]),
        "orc_counts_median_count": (u32, [vp]),
        "orc_counts_remove_low": (None, [vp, u32]),
        "orc_counts_export_sorted": (u64, [vp, vp, vp, u64]),
        "orc_counts_digest": (u64, [vp]),
        "orc_find_oligos": (u64, [vp, vp, u64, u32, u32, vp, vp, u64]),
        "orc_histo_new": (vp, [u64]),
        "orc_histo_free": (None, [vp]),
        "orc_histo_move_count": (None, [vp, u64, u64]),
        "orc_histo_ingest": (None, [vp, vp]),
        "orc_histo_get_vector": (None, [vp, vp]),
        "orc_histo_n_kmers": (u64, [vp]),
        "orc_histo_n_unique": (u64, [vp]),
        "orc_counts_extend_with_histogram": (C.c_int, [vp, vp, vp, C.POINTER(C.c_int)]),
        "orc_run_new": (vp, [u32, u32, u64]),
        "orc_run_free": (None, [vp]),
        "orc_run_push_seq": (C.c_int, [vp, cp, sz]),
        "orc_run_push_lines": (C.c_int, [vp, vp, sz]),
        "orc_run_read_fastq": (C.c_int, [vp, cp, u64, u64]),
        "orc_run_read_fastq_paired": (C.c_int, [vp, cp, cp, u64, u64]),
        "orc_run_finish_ingest": (C.c_int, [vp]),
        "orc_run_consolidate": (C.c_int, [vp]),
        "orc_run_write_histo": (C.c_int, [vp, cp, cp]),
        "orc_run_write_stats": (C.c_int, [vp, cp, cp, cp]),
        "orc_run_error": (cp, [vp]),
        "orc_run_n_chunks": (u32, [vp]),
        "orc_run_n_reads_read": (u64, [vp]),
        "orc_run_n_bases_read": (u64, [vp]),
        "orc_run_n_reads_ingested": (u64, [vp]),
        "orc_run_n_bases_ingested": (u64, [vp]),
        "orc_run_n_kmers_ingested": (u64, [vp]),
        "orc_run_chunk_n_reads": (u64, [vp, u32]),
        "orc_run_chunk_n_bases": (u64, [vp, u32]),
        "orc_run_chunk_n_kmers": (u64, [vp, u32]),
        "orc_run_table": (vp, [vp]),
        "orc_run_histogram": (C.c_int, [vp, u32, vp]),
        "orc_run_n_singletons": (C.c_int, [vp, C.POINTER(u64)]),
        "orc_synth_reads": (None, [u64, u64, u32, u32, u32, u64, u64, vp]),
        "orc_synth_fastq": (C.c_int, [u64, u64, u32, u32, u32, u64, u64, cp, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class OracleError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


def _b(seq) -> bytes:
    return seq.encode() if isinstance(seq, str) else bytes(seq)


# ---- encoding.rs ---------------------------------------------------------

def kmers_from_ascii(seq, k: int) -> list[int]:
    s = _b(seq)
    out = np.empty(max(len(s), 1), dtype=np.uint64)
    n = lib().orc_kmers_from_ascii(s, len(s), k, out.ctypes.data)
    if n < 0:
        raise OracleError(n)
    return out[:n].tolist()


def kmers_via_reads(seq, k: int) -> list[int]:
    s = _b(seq)
    out = np.empty(len(s) + 8, dtype=np.uint64)  # get_kmers emits <=3 padding k-mers before truncating
    n = lib().orc_kmers_via_reads(s, len(s), k, out.ctypes.data)
    if n < 0:
        raise OracleError(n)
    return out[:n].tolist()


def count_valid_bases(seq) -> int:
    s = _b(seq)
    return lib().orc_count_valid_bases(s, len(s))


def revcomp_kmer(kmer: int, k: int) -> int:
    return lib().orc_revcomp_kmer(kmer, k)


def seq_to_kmer(seq) -> int:
    s = _b(seq)
    err = C.c_int(0)
    v = lib().orc_seq_to_kmer(s, len(s), C.byref(err))
    if err.value:
        raise OracleError(err.value)
    return v


def kmer_to_seq(kmer: int, k: int) -> str:
    buf = C.create_string_buffer(k + 1)
    lib().orc_kmer_to_seq(kmer, k, buf)
    return buf.value.decode()


def read_pack(seq) -> tuple[list[int], int]:
    """Read::from_str -> (bytes, length)."""
    s = _b(seq)
    out = np.zeros(len(s) // 4 + 2, dtype=np.uint8)
    n = lib().orc_read_pack(s, len(s), out.ctypes.data)
    if n < 0:
        raise OracleError(n)
    return out[:n].tolist(), len(s)


def read_get_kmers(packed: list[int], length: int, k: int) -> list[int]:
    p = np.asarray(packed, dtype=np.uint8)
    out = np.empty(len(p) * 4 + 8, dtype=np.uint64)
    n = lib().orc_read_get_kmers(p.ctypes.data, len(p), length, k, out.ctypes.data)
    if n < 0:
        raise OracleError(n)
    return out[:n].tolist()


# ---- counting.rs / histogram.rs ------------------------------------------

class KmerCounts:
    def __init__(self, k: int, capacity: int = 0, _borrowed=None):
        self._own = _borrowed is None
        self._h = lib().orc_counts_with_capacity(k, capacity) if self._own else _borrowed

    def __del__(self):
        if getattr(self, "_own", False) and self._h:
            lib().orc_counts_free(self._h)
            self._h = None

    def get_k(self): return lib().orc_counts_k(self._h)
    def ingest_seq(self, seq):
        s = _b(seq)
        rc = lib().orc_counts_ingest_seq(self._h, s, len(s))
        if rc:
            raise OracleError(rc)
    def insert(self, kmer, count): lib().orc_counts_insert(self._h, kmer, count)
    def insert_get(self, kmer, count):
        o, n = C.c_uint32(), C.c_uint32()
        lib().orc_counts_insert_get(self._h, kmer, count, C.byref(o), C.byref(n))
        return o.value, n.value
    def extend(self, other):
        rc = lib().orc_counts_extend(self._h, other._h)
        if rc:
            raise OracleError(rc, "Cannot extend KmerCounts with different k")
    def extend_with_histogram(self, other, histo):
        sat = C.c_int(0)
        rc = lib().orc_counts_extend_with_histogram(self._h, other._h, histo._h, C.byref(sat))
        if rc:
            raise OracleError(rc, "Cannot extend KmerCounts with different k")
        return bool(sat.value)
    def get(self, kmer):
        c = C.c_uint32()
        return c.value if lib().orc_counts_get(self._h, kmer, C.byref(c)) else None
    def get_count(self, kmer): return self.get(kmer) or 0
    def contains(self, kmer): return self.get(kmer) is not None
    def get_canonical_count(self, kmer): return lib().orc_counts_get_canonical_count(self._h, kmer)
    def get_canonical(self, kmer):
        c = C.c_uint32()
        return c.value if lib().orc_counts_get_canonical(self._h, kmer, C.byref(c)) else None
    def len(self): return lib().orc_counts_len(self._h)
    __len__ = len
    def is_empty(self): return self.len() == 0
    def get_n_kmers(self): return lib().orc_counts_n_kmers(self._h)
    def get_n_unique_kmers(self): return lib().orc_counts_len(self._h)
    def get_max_count(self): return lib().orc_counts_max_count(self._h)
    def get_median_count(self): return lib().orc_counts_median_count(self._h)
    def remove_low_count_kmers(self, m): lib().orc_counts_remove_low(self._h, m)
    def digest(self): return lib().orc_counts_digest(self._h)
    def export_sorted(self):
        n = self.len()
        keys = np.empty(n, dtype=np.uint64)
        counts = np.empty(n, dtype=np.uint32)
        lib().orc_counts_export_sorted(self._h, keys.ctypes.data, counts.ctypes.data, n)
        return keys, counts
    def find_oligos(self, oligos, oligo_length, min_count):
        """find_oligos_in_kmers (src/pcr/primers.rs:163-226); sorted (kmers, counts)."""
        o = np.ascontiguousarray(oligos, dtype=np.uint64)
        n = lib().orc_find_oligos(self._h, o.ctypes.data, o.size, oligo_length, min_count, None, None, 0)
        keys = np.empty(n, dtype=np.uint64)
        counts = np.empty(n, dtype=np.uint32)
        lib().orc_find_oligos(self._h, o.ctypes.data, o.size, oligo_length, min_count,
                              keys.ctypes.data, counts.ctypes.data, n)
        return keys, counts
    def filtered_view(self, min_count): return FilteredKmerCounts(self, min_count)


class FilteredKmerCounts:
    def __init__(self, inner: KmerCounts, min_count: int):
        self.inner, self.min_count = inner, min_count
    def get_k(self): return self.inner.get_k()
    def get_canonical(self, kmer):
        c = C.c_uint32()
        ok = lib().orc_filtered_get_canonical(self.inner._h, self.min_count, kmer, C.byref(c))
        return c.value if ok else None
    def get_canonical_count(self, kmer):
        return lib().orc_filtered_get_canonical_count(self.inner._h, self.min_count, kmer)


class Histogram:
    def __init__(self, histo_max: int):
        self.histo_max = histo_max
        self._h = lib().orc_histo_new(histo_max)
    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_histo_free(self._h)
            self._h = None
    @classmethod
    def from_kmer_counts(cls, kc: KmerCounts, histo_max: int):
        h = cls(histo_max)
        lib().orc_histo_ingest(h._h, kc._h)
        return h
    def move_count(self, old, new): lib().orc_histo_move_count(self._h, old, new)
    def get_vector(self):
        out = np.zeros(self.histo_max + 2, dtype=np.uint64)
        lib().orc_histo_get_vector(self._h, out.ctypes.data)
        return out
    def get_n_kmers(self): return lib().orc_histo_n_kmers(self._h)
    def get_n_unique_kmers(self): return lib().orc_histo_n_unique(self._h)


# ---- chunk.rs + io.rs ------------------------------------------------------

class Run:
    """ingest_reads + consolidate_and_histogram (src/io.rs:366-595, 977-1161)."""

    def __init__(self, k: int, chunks: int = 0, histo_max: int = 10000):
        self.k, self.chunks, self.histo_max = k, chunks, histo_max
        self._h = lib().orc_run_new(k, chunks, histo_max)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_run_free(self._h)
            self._h = None

    def _check(self, rc):
        if rc < 0:
            raise OracleError(rc, lib().orc_run_error(self._h).decode(errors="replace"))
        return rc

    def push_seq(self, seq):
        s = _b(seq)
        self._check(lib().orc_run_push_seq(self._h, s, len(s)))

    def push_lines(self, buf):
        """buf: bytes / numpy uint8 of newline-terminated sequences."""
        a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
        self._check(lib().orc_run_push_lines(self._h, a.ctypes.data, a.size))

    def read_fastq(self, path, max_reads=0, validate_every=0):
        return self._check(lib().orc_run_read_fastq(self._h, os.fsencode(path), max_reads, validate_every))

    def read_fastq_paired(self, p1, p2, max_reads=0, validate_every=0):
        if max_reads > 0 and max_reads % 2:
            max_reads += 1
        return self._check(lib().orc_run_read_fastq_paired(
            self._h, os.fsencode(p1), os.fsencode(p2), max_reads, validate_every))

    def finish_ingest(self): self._check(lib().orc_run_finish_ingest(self._h))
    def consolidate(self): self._check(lib().orc_run_consolidate(self._h))
    def finish(self):
        self.finish_ingest()
        self.consolidate()
        return self

    def write_histo(self, directory, sample):
        self._check(lib().orc_run_write_histo(self._h, os.fsencode(directory), sample.encode()))

    def write_stats(self, directory, sample, command="sharkmer"):
        self._check(lib().orc_run_write_stats(self._h, os.fsencode(directory), sample.encode(), command.encode()))

    @property
    def n_chunks(self): return lib().orc_run_n_chunks(self._h)
    @property
    def n_reads_read(self): return lib().orc_run_n_reads_read(self._h)
    @property
    def n_bases_read(self): return lib().orc_run_n_bases_read(self._h)
    @property
    def n_reads_ingested(self): return lib().orc_run_n_reads_ingested(self._h)
    @property
    def n_bases_ingested(self): return lib().orc_run_n_bases_ingested(self._h)
    @property
    def n_kmers_ingested(self): return lib().orc_run_n_kmers_ingested(self._h)
    def chunk_totals(self, c):
        L = lib()
        return (L.orc_run_chunk_n_reads(self._h, c), L.orc_run_chunk_n_bases(self._h, c),
                L.orc_run_chunk_n_kmers(self._h, c))
    def table(self) -> KmerCounts:
        t = KmerCounts(self.k, _borrowed=lib().orc_run_table(self._h))
        t._keepalive = self
        return t
    def histogram(self, chunk_i):
        out = np.zeros(self.histo_max + 2, dtype=np.uint64)
        self._check(lib().orc_run_histogram(self._h, chunk_i, out.ctypes.data))
        return out
    def n_singletons(self):
        v = C.c_uint64()
        self._check(lib().orc_run_n_singletons(self._h, C.byref(v)))
        return v.value


# ---- synthetic reads -------------------------------------------------------

def rate_to_thresh(rate: float) -> int:
    return min(0xFFFFFFFF, int(round(rate * 4294967296.0)))


def synth_reads(seed, genome_len, read_len, sub_rate, n_rate, first, n) -> np.ndarray:
    """Newline-terminated sequence lines of reads [first, first+n) as uint8."""
    out = np.empty(n * (read_len + 1), dtype=np.uint8)
    lib().orc_synth_reads(seed, genome_len, read_len, rate_to_thresh(sub_rate), rate_to_thresh(n_rate),
                          first, n, out.ctypes.data)
    return out


def synth_fastq(path, seed, genome_len, read_len, sub_rate, n_rate, first, n, gzip=False):
    rc = lib().orc_synth_fastq(seed, genome_len, read_len, rate_to_thresh(sub_rate), rate_to_thresh(n_rate),
                               first, n, os.fsencode(path), int(gzip))
    if rc:
        raise OracleError(rc, "synth_fastq")


# ---- brute force, pure Python (independent of the C code above) -------------

_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def py_revcomp(kmer: int, k: int) -> int:
    rc = 0
    for _ in range(k):
        rc = (rc << 2) | (3 - (kmer & 3))
        kmer >>= 2
    return rc


def py_kmers(seq: str, k: int) -> list[int]:
    """Every window of k bases with no N, in order; canonical = min(fwd, revcomp)."""
    out = []
    for i in range(len(seq) - k + 1):
        w = seq[i:i + k]
        if "N" in w:
            continue
        f = 0
        for ch in w:
            f = (f << 2) | _CODE[ch]
        out.append(min(f, py_revcomp(f, k)))
    return out


def py_count(seqs, k: int) -> Counter:
    c = Counter()
    for s in seqs:
        c.update(py_kmers(s, k))
    return c


def py_histogram(counts: Counter, histo_max: int) -> list[int]:
    v = [0] * (histo_max + 2)
    for n in counts.values():
        n = min(n, 0xFFFFFFFF)
        v[n if n <= histo_max else histo_max + 1] += 1
    return v
