"""BASELINE config C4 as data: a random genome plus the amplicon templates of a
primer panel at high copy number, and error-bearing reads sampled from both strands.  Shared by the CPU
check on the oracle table (tests/test_panels.py) and the full-size run on the device table
(tests/test_gpu_zz_c4.py).  numpy only."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_IUPAC = {"A": "A", "C": "C", "G": "G", "T": "T", "R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT", "M": "AC",
          "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT"}
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def rc(s: str) -> str:
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def load_panel():
    """The cnidaria panel as PCRParams, from the committed fixture (tests/golden/make_panel_fixture.py)."""
    from sharkmer_b200.primers import PCRParams
    d = json.load(open(os.path.join(HERE, "golden", "cnidaria_panel.json")))
    return [PCRParams(**p) for p in d["primers"]]


def build_pool(primers, k: int, genome_len: int, copies: int, seed: int):
    """-> (pool: uint8 ASCII array, truth: {gene_name: expected product}).  Every pair gets one concrete reading
    of its degenerate primers, a random insert that puts the amplicon in the middle of [min_length, max_length],
    300 bp flanks; the unit is repeated `copies` times in tandem behind the genome."""
    rng = np.random.default_rng(seed)
    rnd = lambda n: "".join("ACGT"[i] for i in rng.integers(0, 4, n))
    concrete = lambda s: "".join(_IUPAC[c][rng.integers(0, len(_IUPAC[c]))] for c in s.upper())
    parts = [rng.integers(0, 4, genome_len, dtype=np.uint8)]
    truth = {}
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    for p in primers:
        f, r = concrete(p.forward_seq), concrete(p.reverse_seq)
        total = (p.min_length + p.max_length) // 2
        amplicon = f + rnd(total - len(f) - len(r)) + rc(r)
        trim_f, trim_r = min(p.trim, k - 1, len(f)), min(p.trim, k - 1, len(r))
        truth[p.gene_name] = amplicon[len(f) - trim_f:len(amplicon) - (len(r) - trim_r)]
        unit = rnd(300) + amplicon + rnd(300)
        parts.append(np.tile(np.frombuffer(unit.encode(), dtype=np.uint8), copies))
    pool = np.concatenate([lut[parts[0]]] + parts[1:])
    return pool, truth


def sample_reads(pool: np.ndarray, n: int, L: int, err: float, rng) -> np.ndarray:
    """n reads of L bases, newline-terminated, as one (n, L+1) uint8 array: uniform start, each base replaced
    by a random other base with probability err, half of the reads reverse-complemented."""
    pos = rng.integers(0, pool.size - L, n)
    reads = pool[pos[:, None] + np.arange(L, dtype=np.int64)[None, :]]
    hit = rng.random((n, L)) < err
    if hit.any():
        lut = np.frombuffer(b"ACGT", dtype=np.uint8)
        code = np.zeros(256, dtype=np.uint8)
        code[lut] = np.arange(4, dtype=np.uint8)
        old = code[reads[hit]]
        reads[hit] = lut[(old + rng.integers(1, 4, old.size, dtype=np.uint8)) & 3]
    flip = rng.random(n) < 0.5
    reads[flip] = _COMP[reads[flip][:, ::-1]]
    out = np.empty((n, L + 1), dtype=np.uint8)
    out[:, :L] = reads
    out[:, L] = 10
    return out


def fastq_bytes(lines: np.ndarray) -> np.ndarray:
    """(n, L+1) newline-terminated reads -> the bytes of a FASTQ file ('@r' header, constant quality)."""
    n, L1 = lines.shape
    rec = np.empty((n, 3 + L1 + 2 + L1), dtype=np.uint8)
    rec[:, 0:3] = np.frombuffer(b"@r\n", dtype=np.uint8)
    rec[:, 3:3 + L1] = lines
    rec[:, 3 + L1:5 + L1] = np.frombuffer(b"+\n", dtype=np.uint8)
    rec[:, 5 + L1:5 + 2 * L1 - 1] = ord("I")
    rec[:, -1] = 10
    return rec.reshape(-1)
