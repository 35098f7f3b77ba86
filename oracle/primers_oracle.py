"""TEST INFRASTRUCTURE ONLY — CPU restatement of sharkmer's primer preprocessing and primer k-mer
discovery (caseywdunn/sharkmer v3.1.0, src/pcr/primers.rs), the stage of sPCR that scans the whole
count table.  Only tests/ may import this; the product path is sharkmer_b200/primers.py (Python) and
sharkmer_b200/host/primers.hpp (C++) over skm_scan_oligos.

String based, like the reference, so that every step can be checked against the reference's own
unit tests (primers.rs:484-833, pcr/mod.rs:1236-1311): 991 variants of the 18S reverse primer at two
mismatches, one primer k-mer per direction on the 18S table.
"""
from __future__ import annotations

from dataclasses import dataclass
from itertools import combinations as _combinations

import numpy as np

IUPAC = {  # primers.rs:63-76
    "R": "AG", "Y": "CT", "S": "GC", "W": "AT", "K": "GT", "M": "AC",
    "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT",
}
MAX_RESOLVED_VARIANTS = 10_000  # primers.rs:268


class PrimerError(Exception):
    pass


@dataclass
class PCRParams:  # the fields of PCRParams (pcr/mod.rs) this stage reads; defaults of cli.rs:22-24, mod.rs:281
    forward_seq: str
    reverse_seq: str
    gene_name: str = "gene"
    min_count: int = 2
    mismatches: int = 2
    trim: int = 15
    max_primer_kmers: int = 40


def is_valid_nucleotide(c: str) -> bool:  # primers.rs:11-30
    return c in "ACGTRYSWKMBDHVN"


def string_to_oligo(seq: str):  # primers.rs:33-54 -> (length, kmer)
    if len(seq) > 32:
        raise PrimerError(f"Oligo sequence length {len(seq)} exceeds maximum of 32 bases")
    kmer = 0
    for c in seq:
        if c not in "ACGT":
            raise PrimerError(f"Invalid nucleotide {c} in {seq}")
        kmer = (kmer << 2) | "ACGT".index(c)
    return len(seq), kmer


def resolve_primer(primer: str) -> set:  # primers.rs:60-98
    seqs: set = set()
    for nuc in primer:
        poss = IUPAC.get(nuc, nuc)
        seqs = {p for p in poss} if not seqs else {s + p for s in seqs for p in poss}
    return seqs


def combinations(n: int, r: int):  # primers.rs:117-137 (as a list of position lists)
    if r > n:
        return []
    return [list(c) for c in _combinations(range(n), r)]


def permute_sequences(sequences: set, r: int) -> set:  # primers.rs:103-160
    out = set()
    for seq in sequences:
        for positions in combinations(len(seq), r):
            stack = [(seq, 0)]
            while stack:
                s, cur = stack.pop()
                if cur == len(positions):
                    out.add(s)
                    continue
                p = positions[cur]
                for nuc in "ATCG":
                    stack.append((s[:p] + nuc + s[p + 1:], cur + 1))
    return out


def trimmed_primer(params: PCRParams, reverse: bool, k: int) -> str:  # primers.rs:236-263
    primer = params.reverse_seq if reverse else params.forward_seq
    trim = params.trim
    if trim >= k:
        trim = k - 1
    if len(primer) > trim:
        primer = primer[len(primer) - trim:]
    return primer


def preprocess_primer_by_mismatch(params: PCRParams, reverse: bool, k: int):  # primers.rs:231-307
    primer = trimmed_primer(params, reverse, k)
    base = resolve_primer(primer)
    if len(base) > MAX_RESOLVED_VARIANTS:
        raise PrimerError(
            f"Primer {primer} has too many ambiguous bases: {len(base)} resolved variants exceeds limit of "
            f"{MAX_RESOLVED_VARIANTS}. Reduce ambiguity or use a more specific primer.")
    mismatches = min(params.mismatches, len(primer))
    levels = [set(base)]
    seen = set(base)
    for _ in range(mismatches):
        new = permute_sequences(set(seen), 1) - seen
        seen |= new
        levels.append(new)
    return levels


def preprocess_primer(params: PCRParams, reverse: bool, k: int) -> set:  # primers.rs:311-319
    return set().union(*preprocess_primer_by_mismatch(params, reverse, k))


def get_kmers_from_primers(variants, table, min_count: int):  # primers.rs:322-335 -> {kmer: count}
    oligos = [string_to_oligo(v) for v in variants]
    assert oligos, "find_oligos_in_kmers called with no oligos"  # primers.rs:168-171
    length = oligos[0][0]
    keys, counts = table.find_oligos(np.array([o[1] for o in oligos], dtype=np.uint64), length, min_count)
    return {int(a): int(b) for a, b in zip(keys, counts)}


def filter_primer_kmers(matches: dict, cap: int) -> dict:  # primers.rs:347-370
    if len(matches) <= cap:
        return dict(matches)
    entries = sorted(matches.items(), key=lambda kv: (-kv[1], kv[0]))
    return dict(entries[:cap])


def discover_primer_kmers_by_round(levels, table, min_count: int, cap: int) -> dict:  # primers.rs:376-446
    result: dict = {}
    for variants in levels:
        if len(result) >= cap:
            break
        if not variants:
            continue
        rnd = get_kmers_from_primers(variants, table, min_count)
        new = sorted(((k, c) for k, c in rnd.items() if k not in result), key=lambda kv: (-kv[1], kv[0]))
        for k, c in new[:cap - len(result)]:
            result[k] = c
    return result


def get_primer_kmers(params: PCRParams, table):  # primers.rs:448-478
    """`table`: oracle KmerCounts.  The FilteredKmerCounts view the reference passes in does not
    matter here: its iter() yields every entry (counting.rs:343-349) and find_oligos_in_kmers
    filters by params.min_count alone."""
    k = table.get_k()
    fwd = discover_primer_kmers_by_round(preprocess_primer_by_mismatch(params, False, k), table, params.min_count,
                                         params.max_primer_kmers)
    rev = discover_primer_kmers_by_round(preprocess_primer_by_mismatch(params, True, k), table, params.min_count,
                                         params.max_primer_kmers)
    return fwd, rev
