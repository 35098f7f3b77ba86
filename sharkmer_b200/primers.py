"""Primer preprocessing and primer k-mer discovery for sPCR over the device table — the caller on
the far side of the counting path: what `get_primer_kmers`
(caseywdunn/sharkmer v3.1.0, src/pcr/primers.rs:448-478) computes, with the reference's one
expensive step — `find_oligos_in_kmers`, a scan of the WHOLE count table per primer direction and
mismatch level (primers.rs:163-226) — done by `skm_scan_oligos` as one streaming pass over HBM.

Not a translation of the reference's string sets: primer variants are 2-bit packed integers from the
start (the form the device wants), IUPAC expansion is a product over per-position base lists, and a
mismatch level is the Hamming-1 neighbourhood of everything seen so far, built with vectorised
bit-field substitutions and sorted-array set differences.

Same arguments, defaults, limits and error texts as the reference:
  PCRParams fields forward_seq / reverse_seq / trim (15) / mismatches (2) / min_count (2) /
  max_primer_kmers (40)                                   pcr/mod.rs:281, cli.rs:22-24
  trim >= k is clamped to k-1; a primer longer than trim keeps its 3' end   primers.rs:236-263
  more than 10 000 ambiguity-resolved variants is an error                  primers.rs:268-277
  levels are disjoint; level m holds the variants first reached at m mismatches  :279-299
  per level: matches not seen at a lower level, by count descending then k-mer ascending, fill
  what is left of the cap                                                    :376-446
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_BASE = {"A": 0, "C": 1, "G": 2, "T": 3}
_IUPAC = {
    "A": (0,), "C": (1,), "G": (2,), "T": (3,),
    "R": (0, 2), "Y": (1, 3), "S": (2, 1), "W": (0, 3), "K": (2, 3), "M": (0, 1),
    "B": (1, 2, 3), "D": (0, 2, 3), "H": (0, 1, 3), "V": (0, 1, 2), "N": (0, 1, 2, 3),
}
MAX_RESOLVED_VARIANTS = 10_000
DEFAULT_MAX_NUM_PRIMER_KMERS = 40


class PrimerError(ValueError):
    pass


@dataclass
class PCRParams:
    """The fields of PCRParams (src/pcr/mod.rs:148-247) the pipeline reads, with the defaults of
    cli.rs:20-27 and pcr/mod.rs:54,278-283."""
    forward_seq: str
    reverse_seq: str
    gene_name: str = "gene"
    min_count: int = 2
    mismatches: int = 2
    trim: int = 15
    max_primer_kmers: int = DEFAULT_MAX_NUM_PRIMER_KMERS
    min_length: int = 0
    max_length: int = 10000
    dedup_edit_threshold: int = 10
    max_dfs_states: int = 100_000
    max_paths_per_pair: int = 20
    max_node_visits: int = 2
    high_coverage_ratio: float = 10.0
    tip_coverage_fraction: float = 0.1


def string_to_oligo(seq: str):
    """(length, 2-bit packed value); only A/C/G/T (primers.rs:33-54)."""
    if len(seq) > 32:
        raise PrimerError(f"Oligo sequence length {len(seq)} exceeds maximum of 32 bases")
    v = 0
    for c in seq:
        if c not in _BASE:
            raise PrimerError(f"Invalid nucleotide {c} in {seq}")
        v = (v << 2) | _BASE[c]
    return len(seq), v


def oligo_to_string(value: int, length: int) -> str:
    return "".join("ACGT"[(int(value) >> (2 * (length - 1 - i))) & 3] for i in range(length))


def trimmed_primer(params: PCRParams, reverse: bool, k: int) -> str:
    primer = params.reverse_seq if reverse else params.forward_seq
    trim = min(params.trim, k - 1)
    return primer[len(primer) - trim:] if len(primer) > trim else primer


def resolve_primer(primer: str) -> np.ndarray:
    """All ambiguity-free readings of `primer`, packed, ascending.  A character that is neither a
    base nor an IUPAC code is an error (the reference fails on it one step later, in string_to_oligo,
    with the same message)."""
    if len(primer) > 32:
        raise PrimerError(f"Oligo sequence length {len(primer)} exceeds maximum of 32 bases")
    out = np.zeros(1, dtype=np.uint64) if primer else np.zeros(0, dtype=np.uint64)
    n = 1
    for c in primer:
        if c not in _IUPAC:
            raise PrimerError(f"Invalid nucleotide {c} in {primer}")
        n *= len(_IUPAC[c])
    if n > MAX_RESOLVED_VARIANTS and primer:
        raise PrimerError(
            f"Primer {primer} has too many ambiguous bases: {n} resolved variants exceeds limit of "
            f"{MAX_RESOLVED_VARIANTS}. Reduce ambiguity or use a more specific primer.")
    for c in primer:
        choices = np.array(_IUPAC[c], dtype=np.uint64)
        out = ((out[:, None] << np.uint64(2)) | choices[None, :]).reshape(-1)
    return np.unique(out)


def hamming1(variants: np.ndarray, length: int) -> np.ndarray:
    """Everything within one substitution of any element of `variants` (itself included)."""
    if variants.size == 0 or length == 0:
        return variants
    shifts = (np.arange(length, dtype=np.uint64) * np.uint64(2))[None, :, None]
    cleared = variants[:, None, None] & ~(np.uint64(3) << shifts)
    subs = cleared | (np.arange(4, dtype=np.uint64)[None, None, :] << shifts)
    return np.unique(subs.reshape(-1))


def preprocess_primer_by_mismatch(params: PCRParams, reverse: bool, k: int):
    """-> (oligo_length, [level_0, level_1, ...]) with sorted packed oligos per level."""
    primer = trimmed_primer(params, reverse, k)
    base = resolve_primer(primer)
    levels = [base]
    seen = base
    for _ in range(min(params.mismatches, len(primer))):
        ball = hamming1(seen, len(primer))
        levels.append(np.setdiff1d(ball, seen, assume_unique=True))
        seen = ball
    return len(primer), levels


def discover_primer_kmers(engine, length: int, levels, min_count: int, cap: int):
    """-> (kmers, counts), ascending by k-mer: at most `cap` primer k-mers, lower mismatch levels
    first.  `engine.scan_oligos(oligos, length, min_count)` is the device scan."""
    keys = np.zeros(0, dtype=np.uint64)
    counts = np.zeros(0, dtype=np.uint32)
    for oligos in levels:
        if keys.size >= cap:
            break
        if oligos.size == 0:
            continue
        k_new, c_new = engine.scan_oligos(oligos, length, min_count)
        fresh = ~np.isin(k_new, keys)
        k_new, c_new = k_new[fresh], c_new[fresh]
        order = np.lexsort((k_new, -c_new.astype(np.int64)))[:cap - keys.size]
        keys = np.concatenate([keys, k_new[order]])
        counts = np.concatenate([counts, c_new[order]])
    order = np.argsort(keys, kind="stable")
    return keys[order], counts[order]


def get_primer_kmers(params: PCRParams, engine, k: int):
    """((fwd_kmers, fwd_counts), (rev_kmers, rev_counts)) — get_primer_kmers, primers.rs:448-478.
    The scan threshold is params.min_count alone: the reference's FilteredKmerCounts::iter() yields
    every entry whatever the view's threshold (counting.rs:343-349)."""
    out = []
    for reverse in (False, True):
        length, levels = preprocess_primer_by_mismatch(params, reverse, k)  # length <= k-1 by the trim clamp
        out.append(discover_primer_kmers(engine, length, levels, params.min_count, params.max_primer_kmers))
    return out[0], out[1]
