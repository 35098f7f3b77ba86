"""In silico PCR (sPCR) over the device count table, without read threading: primers -> seed graph ->
bidirectional de Bruijn extension under a coverage-threshold sweep -> tip / reachability pruning ->
coverage-ordered path search -> scored, de-duplicated amplicon records.  This is the consumer on the
far side of the counting path and follows, stage by stage and tie-break by
tie-break, caseywdunn/sharkmer v3.1.0:

    get_primer_kmers            src/pcr/primers.rs:448-478     (sharkmer_b200/primers.py)
    create_seed_graph           src/pcr/graph.rs:192-278
    compute_coverage_thresholds src/pcr/mod.rs:405-432
    compute_node_budget         src/pcr/graph.rs:40-52
    extend_graph                src/pcr/graph.rs:322-527
    remove_low_coverage_tips / reachability_pruning   src/pcr/pruning.rs:19-216
    annotate_coverage_ratios    src/pcr/graph.rs:531-544
    get_assembly_paths / generate_sequences_from_paths / sort_and_deduplicate   src/pcr/paths.rs:79-428
    do_pcr / run_pcr            src/pcr/mod.rs:434-795, src/stats.rs:49-155

What is different from the reference is where the table lives.  The reference probes a host hash
map four times per graph node, one node at a time; here the table is in HBM behind
`skm_lookup_batch`, so the extension asks for the four candidate k-mers of EVERY node currently
in the frontier in one device call (one call per frontier wave) and then replays the reference's
node-by-node logic on the host from the prefetched answers.  The FIFO order, the periodically
refreshed median, the node budget and therefore the resulting graph are identical; the number of
device round trips drops from 4 x nodes to the depth of the graph.

`table` everywhere below is anything with `lookup(kmers, min_count, mode) -> (counts, found)` and
`scan_oligos(oligos, length, min_count) -> (kmers, counts)`: `sharkmer_b200.kmer.Engine`.
Read threading (src/pcr/threading.rs, bubble.rs; opt-in `--read-threading` in the reference) is not
built: scores use the reference's "no threading data" branch.
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .common import revcomp_kmer
from .primers import PCRParams, get_primer_kmers

COVERAGE_MULTIPLIER = 2            # pcr/mod.rs:45
COVERAGE_STEPS = 4                 # pcr/mod.rs:48
EXTENSION_EVALUATION_FREQUENCY = 1000   # graph.rs:16
DEFAULT_MAX_NUM_NODES = 500_000    # graph.rs:22
MIN_NODE_BUDGET = 100_000          # graph.rs:25
BUDGET_LERP_LOW_BP = 150_000_000   # graph.rs:28
BUDGET_LERP_HIGH_BP = 750_000_000  # graph.rs:31
MAX_NUM_AMPLICONS = 20             # paths.rs:20
FASTA_LINE_WIDTH = 80              # io.rs:14


# ---- a digraph with petgraph's StableDiGraph iteration orders ------------------------------------

class DiGraph:
    """Nodes and edges keep their indices when others are removed; a node's edge lists are walked
    newest edge first (petgraph pushes new edges at the head of its adjacency lists), which is what
    decides ties in the path search."""

    def __init__(self):
        self.nodes = []      # [sub_kmer, is_start, is_end] or None
        self.edges = []      # [src, dst, count, coverage_ratio] or None
        self._out = []       # edge ids per node, oldest first
        self._in = []
        self.n_nodes = 0
        self.n_edges = 0

    def copy(self):
        g = DiGraph()
        g.nodes = [None if n is None else list(n) for n in self.nodes]
        g.edges = [None if e is None else list(e) for e in self.edges]
        g._out = [list(x) for x in self._out]
        g._in = [list(x) for x in self._in]
        g.n_nodes, g.n_edges = self.n_nodes, self.n_edges
        return g

    def add_node(self, sub_kmer, is_start=False, is_end=False):
        self.nodes.append([sub_kmer, is_start, is_end])
        self._out.append([])
        self._in.append([])
        self.n_nodes += 1
        return len(self.nodes) - 1

    def add_edge(self, a, b, count, coverage_ratio=0.0):
        self.edges.append([a, b, count, coverage_ratio])
        e = len(self.edges) - 1
        self._out[a].append(e)
        self._in[b].append(e)
        self.n_edges += 1
        return e

    def remove_node(self, n):
        for e in list(self._out[n]) + list(self._in[n]):
            if self.edges[e] is None:
                continue
            a, b = self.edges[e][0], self.edges[e][1]
            self._out[a].remove(e)
            self._in[b].remove(e)
            self.edges[e] = None
            self.n_edges -= 1
        self.nodes[n] = None
        self.n_nodes -= 1

    def node_indices(self):
        return [i for i, n in enumerate(self.nodes) if n is not None]

    def edge_indices(self):
        return [i for i, e in enumerate(self.edges) if e is not None]

    def out_edges(self, n):   # newest first
        return self._out[n][::-1]

    def in_edges(self, n):
        return self._in[n][::-1]

    def out_degree(self, n):
        return len(self._out[n])

    def in_degree(self, n):
        return len(self._in[n])

    def find_edge(self, a, b):
        for e in self.out_edges(a):
            if self.edges[e][1] == b:
                return e
        return None

    def edge_counts(self):
        return [e[2] for e in self.edges if e is not None]


def median_f64(values):
    """graph.rs:81-113: None when empty; the mean of the two middle values for an even length."""
    if not values:
        return None
    v = sorted(values)
    mid = len(v) // 2
    return float(v[mid]) if len(v) % 2 else (float(v[mid - 1]) + float(v[mid])) / 2.0


def compute_mean(values):       # graph.rs:176-182
    return sum(values) / len(values) if values else 0.0


def compute_median(values):     # graph.rs:184-189
    m = median_f64(list(values))
    return 0.0 if m is None else m


def compute_node_budget(n_bases_ingested: int) -> int:   # graph.rs:40-52
    if n_bases_ingested <= BUDGET_LERP_LOW_BP:
        return MIN_NODE_BUDGET
    if n_bases_ingested >= BUDGET_LERP_HIGH_BP:
        return DEFAULT_MAX_NUM_NODES
    fraction = (n_bases_ingested - BUDGET_LERP_LOW_BP) / float(BUDGET_LERP_HIGH_BP - BUDGET_LERP_LOW_BP)
    return int(MIN_NODE_BUDGET + fraction * (DEFAULT_MAX_NUM_NODES - MIN_NODE_BUDGET))


def get_suffix_mask(k: int) -> int:   # graph.rs:57-60
    return (1 << (2 * (k - 1))) - 1


def compute_coverage_thresholds(primer_count: int, min_count: int):   # pcr/mod.rs:405-432
    high = primer_count // COVERAGE_MULTIPLIER
    if high <= min_count:
        t = [min_count]
    else:
        step = (high - min_count) // (COVERAGE_STEPS - 1)
        t = [max(0, high - i * step) for i in range(COVERAGE_STEPS)]
        t[-1] = min_count
    out = []
    for x in t:   # Vec::dedup: consecutive repeats only
        if not out or out[-1] != x:
            out.append(x)
    return out


def kmer_to_seq(kmer: int, k: int) -> str:
    return "".join("ACGT"[(int(kmer) >> (2 * (k - 1 - i))) & 3] for i in range(k))


def primer_counts_max_median(counts):
    """KmerCounts::get_max_count / get_median_count (counting.rs:271-298): integer median, the even
    case as half-sums."""
    c = sorted(int(x) for x in counts)
    if not c:
        return 0, 0
    mid = len(c) // 2
    med = c[mid] if len(c) % 2 else c[mid - 1] // 2 + c[mid] // 2
    return c[-1], med


# ---- seed graph + extension ---------------------------------------------------------------------

def create_seed_graph(forward_kmers, reverse_kmers, k: int):   # graph.rs:192-278
    g = DiGraph()
    lookup = {}
    mask = get_suffix_mask(k)
    for kmer in sorted(int(x) for x in forward_kmers):
        sub = kmer >> 2
        if sub in lookup:
            g.nodes[lookup[sub]][1] = True
        else:
            lookup[sub] = g.add_node(sub, True, False)
    for kmer in sorted(int(x) for x in reverse_kmers):
        sub = revcomp_kmer(kmer, k) & mask
        if sub in lookup:
            g.nodes[lookup[sub]][2] = True
        else:
            lookup[sub] = g.add_node(sub, False, True)
    return g, lookup


FORWARD, REVERSE = 0, 1


class _WaveLookups:
    """Candidate k-mer lookups for the extension, one device call per frontier wave.  When the wave
    is narrow (a linear stretch has two entries, one per direction) the call also asks, blindly, for
    the candidates of the candidates, several levels deep, up to `budget` k-mers: what a node needs
    depends only on its sub-k-mer and direction, so answers are cached under that key and a level
    of the graph that was guessed right costs no round trip at all.  The guesses are enumerated
    with array arithmetic, not per k-mer."""

    def __init__(self, table, view_min_count: int, k: int, budget: int = 4096):
        self.table, self.view_min, self.k, self.budget = table, view_min_count, k, budget
        self.shift = np.uint64(2 * (k - 1))
        self.mask = np.uint64(get_suffix_mask(k))
        self.row = {}         # (sub_kmer << 1 | dir) -> (wave, row) of the cached answer
        self.waves = {}       # wave id -> (kmers[n,4], counts[n,4], found[n,4], rows still cached)
        self.next_wave = 0
        self.calls = 0
        self.kmers_asked = 0

    def _candidates(self, keys):
        """keys: packed (sub_kmer << 1 | dir) -> candidate k-mers [n, 4] and the children's keys [n, 4]."""
        sub, d = keys >> np.uint64(1), keys & np.uint64(1)
        base = np.arange(4, dtype=np.uint64)[None, :]
        fwd = (sub[:, None] << np.uint64(2)) | base
        rev = (base << self.shift) | sub[:, None]
        is_fwd = (d == 0)[:, None]
        cands = np.where(is_fwd, fwd, rev)
        child = np.where(is_fwd, cands & self.mask, cands >> np.uint64(2))
        return cands, (child << np.uint64(1)) | d[:, None]

    def get(self, graph: DiGraph, frontier, node: int, direction: int):
        key = (graph.nodes[node][0] << 1) | direction
        if key not in self.row:
            first = [key] + [(graph.nodes[n][0] << 1) | d for (n, d) in frontier]
            first = [q for q in dict.fromkeys(first) if q not in self.row]
            level = np.array(first, dtype=np.uint64)
            todo, cands_all = [], []
            n_todo = 0
            while level.size and (not todo or 4 * (n_todo + level.size) <= self.budget):
                cands, child = self._candidates(level)
                todo.append(level)
                cands_all.append(cands)
                n_todo += level.size
                nxt = np.unique(child.reshape(-1))
                known = np.concatenate(todo)
                nxt = nxt[~np.isin(nxt, known)]
                if self.row and nxt.size:
                    nxt = np.array([q for q in nxt.tolist() if q not in self.row], dtype=np.uint64)
                level = nxt
            keys = np.concatenate(todo)
            kmers = np.concatenate(cands_all)
            counts, found = self.table.lookup(kmers.reshape(-1), self.view_min, _lib.LOOKUP_EITHER)
            self.calls += 1
            self.kmers_asked += int(kmers.size)
            w = self.next_wave
            self.next_wave += 1
            self.waves[w] = [kmers, np.asarray(counts).reshape(-1, 4), np.asarray(found).reshape(-1, 4), keys.size]
            self.row.update(zip(keys.tolist(), ((w, i) for i in range(keys.size))))
            if len(self.row) > 1_000_000:   # guesses that were never needed
                keep = {(graph.nodes[n][0] << 1) | d for (n, d) in frontier}
                keep.add(key)
                self.row = {q: v for q, v in self.row.items() if q in keep}
                alive = {v[0] for v in self.row.values()}
                self.waves = {i: x for i, x in self.waves.items() if i in alive}
        w, i = self.row.pop(key)
        wave = self.waves[w]
        out = (wave[0][i].tolist(), wave[1][i].tolist(), wave[2][i].tolist())
        wave[3] -= 1
        if wave[3] == 0:
            del self.waves[w]
        return out


def extend_graph(graph: DiGraph, node_lookup: dict, table, view_min_count: int, min_count: int,
                 params: PCRParams, max_num_nodes: int, k: int, log=None):
    """graph.rs:322-527.  Returns (graph, node_lookup, found_path, n_device_calls)."""
    mask = get_suffix_mask(k)
    found_path = False
    counts_now = graph.edge_counts()
    median_edge_count = median_f64(counts_now) if counts_now else None
    if median_edge_count is None:
        median_edge_count = float(min_count)
    last_median_check = 0
    frontier = deque()
    for n in graph.node_indices():
        if graph.nodes[n][1]:
            frontier.append((n, FORWARD))
        if graph.nodes[n][2]:
            frontier.append((n, REVERSE))
    processed = (set(), set())
    added_by = (set(n for n in graph.node_indices() if graph.nodes[n][1]),
                set(n for n in graph.node_indices() if graph.nodes[n][2]))
    waves = _WaveLookups(table, view_min_count, k)
    while frontier:
        node, d = frontier.popleft()
        if node in processed[d]:
            continue
        processed[d].add(node)
        n_nodes = graph.n_nodes
        if n_nodes > max_num_nodes:
            if log:
                log(f"There are {n_nodes} nodes in the graph. This exceeds the maximum of {max_num_nodes}, abandoning search.")
            break
        if n_nodes > last_median_check and (n_nodes - last_median_check) > EXTENSION_EVALUATION_FREQUENCY:
            m = median_f64(graph.edge_counts())
            median_edge_count = float(min_count) if m is None else m
            last_median_check = n_nodes - (n_nodes % EXTENSION_EVALUATION_FREQUENCY)
        sub = graph.nodes[node][0]
        kmers, counts, found = waves.get(graph, frontier, node, d)
        cands = [(kmers[b], counts[b]) for b in range(4) if found[b] and counts[b] >= min_count]
        for kmer, count in cands:
            new_sub = (kmer & mask) if d == FORWARD else (kmer >> 2)
            if new_sub == sub:      # self loop
                continue
            existing = node_lookup.get(new_sub)
            if existing is not None:
                if d == FORWARD:
                    if graph.find_edge(node, existing) is None:
                        graph.add_edge(node, existing, count)
                        if existing in added_by[REVERSE]:
                            found_path = True
                else:
                    if graph.find_edge(existing, node) is None:
                        graph.add_edge(existing, node, count)
                        if existing in added_by[FORWARD]:
                            found_path = True
            else:
                if float(count) > median_edge_count * params.high_coverage_ratio:
                    continue        # likely repetitive
                new = graph.add_node(new_sub)
                node_lookup[new_sub] = new
                added_by[d].add(new)
                if d == FORWARD:
                    graph.add_edge(node, new, count)
                else:
                    graph.add_edge(new, node, count)
                frontier.append((new, d))
    return graph, node_lookup, found_path, waves.calls


def annotate_coverage_ratios(graph: DiGraph):   # graph.rs:531-544
    median = median_f64(graph.edge_counts())
    if median is None or median <= 0.0:
        return
    for e in graph.edge_indices():
        graph.edges[e][3] = graph.edges[e][2] / median


# ---- pruning ------------------------------------------------------------------------------------

def _tip_length_backward(g: DiGraph, node: int) -> int:   # pruning.rs:99-124
    length, cur = 0, node
    while True:
        length += 1
        inc = g.in_edges(cur)
        if len(inc) != 1:
            break
        parent = g.edges[inc[0]][0]
        if g.out_degree(parent) > 1 or g.nodes[parent][1]:
            break
        cur = parent
    return length


def _tip_length_forward(g: DiGraph, node: int) -> int:    # pruning.rs:128-149
    length, cur = 0, node
    while True:
        length += 1
        out = g.out_edges(cur)
        if len(out) != 1:
            break
        child = g.edges[out[0]][1]
        if g.in_degree(child) > 1 or g.nodes[child][2]:
            break
        cur = child
    return length


def remove_low_coverage_tips(g: DiGraph, k: int, tip_coverage_fraction: float):   # pruning.rs:19-95
    median = median_f64(g.edge_counts())
    min_tip_count = max((1.0 if median is None else median) * tip_coverage_fraction, 1.0)
    removed = 1
    while removed > 0:
        removed = 0
        doomed = []
        for n in g.node_indices():
            if g.nodes[n][1] or g.nodes[n][2]:
                continue
            no_out, no_in = g.out_degree(n) == 0, g.in_degree(n) == 0
            if not no_out and not no_in:
                continue
            if no_out:
                if _tip_length_backward(g, n) >= k:
                    continue
                if float(max((g.edges[e][2] for e in g.in_edges(n)), default=0)) >= min_tip_count:
                    continue
            if no_in:
                if _tip_length_forward(g, n) >= k:
                    continue
                if float(max((g.edges[e][2] for e in g.out_edges(n)), default=0)) >= min_tip_count:
                    continue
            doomed.append(n)
        for n in doomed:
            g.remove_node(n)
            removed += 1


def reachability_pruning(g: DiGraph):   # pruning.rs:166-216
    fwd, stack = set(), [n for n in g.node_indices() if g.nodes[n][1]]
    while stack:
        n = stack.pop()
        if n not in fwd:
            fwd.add(n)
            stack.extend(g.edges[e][1] for e in g.out_edges(n))
    bwd, stack = set(), [n for n in g.node_indices() if g.nodes[n][2]]
    while stack:
        n = stack.pop()
        if n not in bwd:
            bwd.add(n)
            stack.extend(g.edges[e][0] for e in g.in_edges(n))
    for n in [n for n in g.node_indices() if n not in fwd or n not in bwd]:
        g.remove_node(n)


# ---- paths --------------------------------------------------------------------------------------

def sorted_children(g: DiGraph, node: int):   # paths.rs:42-65 (no edge preferences without threading)
    out = [(g.edges[e][1], e, float(g.edges[e][2])) for e in g.out_edges(node)]
    out.sort(key=lambda t: t[2])   # stable, ascending: pop() takes the highest score
    return out


def get_assembly_paths(g: DiGraph, k: int, params: PCRParams):   # paths.rs:79-196
    """-> list of paths, each a list of (node, edge or None)."""
    min_path_nodes = 1 if params.min_length <= k else params.min_length - k + 2
    max_path_nodes = 1 if params.max_length <= k else params.max_length - k + 2
    end_nodes = set(n for n in g.node_indices() if g.nodes[n][2])
    all_paths = []
    for start in [n for n in g.node_indices() if g.nodes[n][1]]:
        paths_from_start = states = 0
        path = [(start, None)]
        visits = {start: 1}
        child_stack = [sorted_children(g, start)]
        while True:
            if paths_from_start >= params.max_paths_per_pair or states >= params.max_dfs_states:
                break
            frame = child_stack[-1]
            if frame:
                neighbor, edge, _ = frame.pop()
                states += 1
                if visits.get(neighbor, 0) >= params.max_node_visits:
                    continue
                path.append((neighbor, edge))
                visits[neighbor] = visits.get(neighbor, 0) + 1
                if neighbor in end_nodes and len(path) >= min_path_nodes:
                    all_paths.append(list(path))
                    paths_from_start += 1
                    visits[neighbor] -= 1
                    path.pop()
                    continue
                if len(path) >= max_path_nodes:
                    visits[neighbor] -= 1
                    path.pop()
                    continue
                child_stack.append(sorted_children(g, neighbor))
            else:
                child_stack.pop()
                if not child_stack:
                    break
                back, _ = path.pop()
                visits[back] -= 1
    return all_paths


@dataclass
class PathScore:   # pcr/mod.rs:58-112, the branch without threading data
    kmer_min_count: int
    kmer_median_count: float
    coverage_cv: float
    max_coverage_ratio: float

    def composite(self) -> float:
        cv_penalty = 1.0 / self.coverage_cv if self.coverage_cv > 1.0 else 1.0
        repeat_penalty = 5.0 / self.max_coverage_ratio if self.max_coverage_ratio > 5.0 else 1.0
        return self.kmer_median_count * cv_penalty * repeat_penalty * 1.0


@dataclass
class Record:
    id: str
    desc: str
    seq: str
    score: PathScore = field(repr=False, default=None)


def _f64_display(x: float) -> str:
    """Rust's `{}` for an f64: shortest round-trip digits, no exponent, no trailing '.0'."""
    if x == int(x) and abs(x) < 1e16:
        return str(int(x))
    return repr(float(x))


def generate_sequences_from_paths(g: DiGraph, paths, k: int, sample_name: str, params: PCRParams,
                                  amplicon_index: int = 0):   # paths.rs:200-377
    records = []
    for path in paths:
        seq = ""
        edge_counts, path_edges = [], []
        for node, edge in path:
            sub = g.nodes[node][0]
            if not seq:
                seq = kmer_to_seq(sub, k - 1)
            else:
                seq += "ACGT"[sub & 3]
                edge_counts.append(g.edges[edge][2])
                path_edges.append(edge)
        if len(seq) < params.min_length or not edge_counts:
            continue
        mean = compute_mean(edge_counts)
        median = compute_median(edge_counts)
        cmin, cmax = min(edge_counts), max(edge_counts)
        if mean > 0.0:
            var = sum((c - mean) * (c - mean) for c in edge_counts) / len(edge_counts)
            cv = var ** 0.5 / mean
        else:
            cv = 0.0
        max_ratio = 0.0
        for e in path_edges:
            max_ratio = max(max_ratio, g.edges[e][3])
        score = PathScore(cmin, median, cv, max_ratio)
        rid = f"{sample_name}_{params.gene_name}_{amplicon_index}"
        desc = (f"sample={sample_name} gene={params.gene_name} product={amplicon_index} length={len(seq)} "
                f"kmer_count_mean={mean:.2f} kmer_count_median={_f64_display(median)} kmer_count_min={cmin} "
                f"kmer_count_max={cmax} score={score.composite():.2f}")
        amplicon_index += 1
        records.append(Record(rid, desc, seq, score))
    return records, amplicon_index


def bounded_levenshtein(a: str, b: str, k: int):
    """Edit distance if it is <= k, else None (bio::alignment::distance::simd::bounded_levenshtein)."""
    if abs(len(a) - len(b)) > k:
        return None
    if len(a) > len(b):
        a, b = b, a
    big = k + 1
    prev = {j: j for j in range(0, min(len(b), k) + 1)}
    for i in range(1, len(a) + 1):
        lo, hi = max(0, i - k), min(len(b), i + k)
        cur = {}
        for j in range(lo, hi + 1):
            if j == 0:
                cur[j] = i
                continue
            best = prev.get(j - 1, big) + (a[i - 1] != b[j - 1])
            best = min(best, prev.get(j, big) + 1, cur.get(j - 1, big) + 1)
            cur[j] = min(best, big)
        if min(cur.values()) > k:
            return None
        prev = cur
    d = prev.get(len(b), big)
    return d if d <= k else None


def sort_and_deduplicate(records, params: PCRParams):   # paths.rs:381-428
    ordered = sorted(records, key=lambda r: (-r.score.composite(), r.seq.encode()))
    kept = []
    for r in ordered:
        if not any(bounded_levenshtein(r.seq, q.seq, params.dedup_edit_threshold) is not None for q in kept):
            kept.append(r)
    return kept[:MAX_NUM_AMPLICONS]


# ---- the pipeline -------------------------------------------------------------------------------

def validate_pcr_params(p: PCRParams):   # pcr/mod.rs:296-401 -> [(error, suggestion)]
    from .primers import _IUPAC
    errors = []
    if len(p.forward_seq) < 2:
        errors.append((f"Forward primer sequence is too short: '{p.forward_seq}'", "Primer sequences must be at least 2 bases"))
    if len(p.reverse_seq) < 2:
        errors.append((f"Reverse primer sequence is too short: '{p.reverse_seq}'", "Primer sequences must be at least 2 bases"))
    for name, seq in (("forward", p.forward_seq), ("reverse", p.reverse_seq)):
        bad = [c for c in seq if c not in _IUPAC] if len(seq) >= 2 else []
        if bad:
            errors.append((f"Invalid nucleotide(s) {', '.join(bad)} in {name} primer {seq}",
                           "Valid characters: A C G T R Y W S M K B D H V N"))
    if p.min_length > p.max_length:
        errors.append((f"min-length ({p.min_length}) is greater than max-length ({p.max_length})", "Swap the values or adjust the range"))
    if p.min_count < 2:
        errors.append((f"min-count is {p.min_count}, must be at least 2", "Set min-count to at least 2"))
    if p.max_length == 0:
        errors.append(("max-length is 0", "Set max-length to a positive value"))
    if not p.gene_name:
        errors.append(("Gene name is empty", "Provide a unique name for the primer pair via the 'name' field"))
    if p.forward_seq == p.reverse_seq and len(p.forward_seq) >= 2:
        errors.append((f"Forward and reverse primers are identical: {p.forward_seq}",
                       "Check that forward and reverse sequences are not swapped"))
    return errors


@dataclass
class PcrOutcome:   # pcr/mod.rs:286-291
    records: list
    failure_reason: str | None
    stats: dict = field(default_factory=dict)


def do_pcr(table, k: int, sample_name: str, params: PCRParams, max_num_nodes: int = DEFAULT_MAX_NUM_NODES,
           view_min_count: int = 2, log=None) -> PcrOutcome:
    """pcr/mod.rs:434-795 without read threading.  `view_min_count`: the `--min-kmer-count` filter
    of the FilteredKmerCounts view (stats.rs:81), default 2."""
    (fk, fc), (rk, rc) = get_primer_kmers(params, table, k)
    if fk.size == 0 or rk.size == 0:
        which = ("forward and reverse primers" if fk.size == 0 and rk.size == 0 else
                 "forward primer" if fk.size == 0 else "reverse primer")
        return PcrOutcome([], f"{which} not found")
    seed, _ = create_seed_graph(fk, rk, k)
    max_f, _ = primer_counts_max_median(fc)
    max_r, _ = primer_counts_max_median(rc)
    thresholds = compute_coverage_thresholds(min(max_f, max_r), params.min_count)
    failure = "no path found"
    found_signal = False
    current = seed.copy()
    calls = 0
    for min_count in thresholds:
        fresh = seed.copy()
        lookup = {fresh.nodes[n][0]: n for n in fresh.node_indices()}
        current, _, found, c = extend_graph(fresh, lookup, table, view_min_count, min_count, params, max_num_nodes, k, log)
        calls += c
        if found:
            found_signal = True
            break
    if current.n_nodes >= max_num_nodes:
        failure = "node budget exceeded"
    stats = {"thresholds": thresholds, "nodes": current.n_nodes, "edges": current.n_edges, "lookup_calls": calls,
             "forward_primer_kmers": int(fk.size), "reverse_primer_kmers": int(rk.size)}
    records = []
    if found_signal:
        pruned = current.copy()
        remove_low_coverage_tips(pruned, k, params.tip_coverage_fraction)
        reachability_pruning(pruned)
        annotate_coverage_ratios(pruned)
        stats.update(pruned_nodes=pruned.n_nodes, pruned_edges=pruned.n_edges)
        paths = get_assembly_paths(pruned, k, params)
        stats["paths"] = len(paths)
        if paths:
            records, _ = generate_sequences_from_paths(pruned, paths, k, sample_name, params, 0)
            if records:
                failure = None
    if not records:
        return PcrOutcome([], failure, stats)
    out = []
    for i, r in enumerate(sort_and_deduplicate(records, params)):   # renumber after dedup (mod.rs:759-790)
        desc = " ".join(f"product={i}" if f.startswith("product=") else f for f in r.desc.split())
        out.append(Record(f"{sample_name}_{params.gene_name}_{i}", desc, r.seq, r.score))
    return PcrOutcome(out, None, stats)


def write_fasta(path: str, records):   # io.rs:144-158
    with open(path, "w") as f:
        for r in records:
            f.write(f">{r.id} {r.desc}\n")
            for i in range(0, len(r.seq), FASTA_LINE_WIDTH):
                f.write(r.seq[i:i + FASTA_LINE_WIDTH] + "\n")


def run_pcr(table, k: int, pcr_runs, sample: str, directory: str, min_kmer_count: int = 2,
            max_nodes: int = DEFAULT_MAX_NUM_NODES):
    """stats.rs:49-155: one {directory}{sample}_{gene}.fasta per gene with products; returns the
    per-gene result dicts of the stats file."""
    results = []
    for p in pcr_runs:
        if p.min_count < min_kmer_count:      # cli.rs:556-570
            p = PCRParams(**{**p.__dict__, "min_count": min_kmer_count})
        out = do_pcr(table, k, sample, p, max_nodes, min_kmer_count)
        if out.records:
            prefix = directory if (not directory or directory.endswith("/")) else directory + "/"
            write_fasta(f"{prefix}{sample}_{p.gene_name}.fasta", out.records)
            results.append({"gene_name": p.gene_name, "status": "success", "n_products": len(out.records),
                            "product_lengths": [len(r.seq) for r in out.records], "failure_reason": None})
        else:
            results.append({"gene_name": p.gene_name, "status": "fail", "n_products": 0, "product_lengths": [],
                            "failure_reason": out.failure_reason or "unknown (no reason reported by PCR pipeline)"})
    return results
