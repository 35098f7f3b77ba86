// pcr.hpp — in silico PCR (sPCR) over the device count table, without read threading; the C++ twin
// of sharkmer_b200/pcr.py (same stages, same tie-breaks, byte-identical FASTA records).  Follows
// caseywdunn/sharkmer v3.1.0 stage by stage:
//   parse_pcr_primers_string     src/cli.rs:12-140
//   validate_pcr_params          src/pcr/mod.rs:296-401
//   get_primer_kmers             src/pcr/primers.rs:448-478         (primers.hpp)
//   create_seed_graph            src/pcr/graph.rs:192-278
//   compute_coverage_thresholds  src/pcr/mod.rs:405-432
//   compute_node_budget          src/pcr/graph.rs:40-52
//   extend_graph                 src/pcr/graph.rs:322-527
//   remove_low_coverage_tips / reachability_pruning     src/pcr/pruning.rs:19-216
//   annotate_coverage_ratios     src/pcr/graph.rs:531-544
//   get_assembly_paths / generate_sequences_from_paths / sort_and_deduplicate   src/pcr/paths.rs:42-428
//   do_pcr / run_pcr / FASTA     src/pcr/mod.rs:434-795, src/stats.rs:49-155, src/io.rs:144-158
// The table stays in HBM: the extension asks skm_lookup_batch for the candidate k-mers of the whole
// frontier (plus, when the frontier is narrow, blind guesses several levels ahead) in one call and
// replays the reference's node-at-a-time logic from the cached answers — identical graph, one device
// round trip per wave instead of four probes per node.
#pragma once

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <functional>
#include <set>
#include <unordered_map>
#include <unordered_set>

#include "primers.hpp"

namespace skm {
namespace pcr {

constexpr uint32_t COVERAGE_MULTIPLIER = 2, COVERAGE_STEPS = 4;   // pcr/mod.rs:45,48
constexpr size_t EXTENSION_EVALUATION_FREQUENCY = 1000;           // graph.rs:16
constexpr size_t DEFAULT_MAX_NUM_NODES = 500000, MIN_NODE_BUDGET = 100000;   // graph.rs:22,25
constexpr uint64_t BUDGET_LERP_LOW_BP = 150000000ull, BUDGET_LERP_HIGH_BP = 750000000ull;   // graph.rs:28,31
constexpr size_t MAX_NUM_AMPLICONS = 20;                          // paths.rs:20
constexpr size_t FASTA_LINE_WIDTH = 80;                           // io.rs:14

// The fields of PCRParams (pcr/mod.rs:148-247) the pipeline reads beyond the primer stage.
struct Params : PCRParams {
    size_t min_length = 0, max_length = 10000;
    uint32_t dedup_edit_threshold = 10;
    size_t max_dfs_states = 100000, max_paths_per_pair = 20, max_node_visits = 2;
    double high_coverage_ratio = 10.0, tip_coverage_fraction = 0.1;
};

// ---- a digraph with petgraph's StableDiGraph iteration orders (newest edge first) ---------------
struct DiGraph {
    struct Node { uint64_t sub_kmer; bool is_start, is_end, alive; };
    struct Edge { uint32_t src, dst, count; double coverage_ratio; bool alive; };
    std::vector<Node> nodes;
    std::vector<Edge> edges;
    std::vector<std::vector<uint32_t>> out, in;   // edge ids per node, oldest first
    size_t n_nodes = 0, n_edges = 0;

    uint32_t add_node(uint64_t sub, bool s = false, bool e = false) {
        nodes.push_back(Node{sub, s, e, true});
        out.emplace_back();
        in.emplace_back();
        n_nodes++;
        return (uint32_t)nodes.size() - 1;
    }
    uint32_t add_edge(uint32_t a, uint32_t b, uint32_t count, double ratio = 0.0) {
        edges.push_back(Edge{a, b, count, ratio, true});
        const uint32_t e = (uint32_t)edges.size() - 1;
        out[a].push_back(e);
        in[b].push_back(e);
        n_edges++;
        return e;
    }
    void remove_node(uint32_t n) {
        std::vector<uint32_t> inc(out[n]);
        inc.insert(inc.end(), in[n].begin(), in[n].end());
        for (uint32_t e : inc) {
            if (!edges[e].alive) continue;
            auto &o = out[edges[e].src];
            o.erase(std::find(o.begin(), o.end(), e));
            auto &i = in[edges[e].dst];
            i.erase(std::find(i.begin(), i.end(), e));
            edges[e].alive = false;
            n_edges--;
        }
        nodes[n].alive = false;
        n_nodes--;
    }
    std::vector<uint32_t> node_indices() const {
        std::vector<uint32_t> v;
        for (uint32_t i = 0; i < nodes.size(); i++)
            if (nodes[i].alive) v.push_back(i);
        return v;
    }
    // a node's edges, newest first
    std::vector<uint32_t> out_edges(uint32_t n) const { return std::vector<uint32_t>(out[n].rbegin(), out[n].rend()); }
    std::vector<uint32_t> in_edges(uint32_t n) const { return std::vector<uint32_t>(in[n].rbegin(), in[n].rend()); }
    bool has_edge(uint32_t a, uint32_t b) const {
        for (uint32_t e : out[a])
            if (edges[e].dst == b) return true;
        return false;
    }
    std::vector<uint64_t> edge_counts() const {
        std::vector<uint64_t> v;
        for (auto &e : edges)
            if (e.alive) v.push_back(e.count);
        return v;
    }
};

// graph.rs:81-113: false when empty; the mean of the two middle values for an even length
inline bool median_f64(std::vector<uint64_t> v, double &out) {
    if (v.empty()) return false;
    std::sort(v.begin(), v.end());
    const size_t mid = v.size() / 2;
    out = v.size() % 2 ? (double)v[mid] : ((double)v[mid - 1] + (double)v[mid]) / 2.0;
    return true;
}
inline double compute_mean(const std::vector<uint64_t> &v) {   // graph.rs:176-182
    if (v.empty()) return 0.0;
    uint64_t s = 0;
    for (uint64_t x : v) s += x;
    return (double)s / (double)v.size();
}
inline double compute_median(const std::vector<uint64_t> &v) {   // graph.rs:184-189
    double m = 0.0;
    return median_f64(v, m) ? m : 0.0;
}
inline size_t compute_node_budget(uint64_t n_bases_ingested) {   // graph.rs:40-52
    if (n_bases_ingested <= BUDGET_LERP_LOW_BP) return MIN_NODE_BUDGET;
    if (n_bases_ingested >= BUDGET_LERP_HIGH_BP) return DEFAULT_MAX_NUM_NODES;
    const double fraction = (double)(n_bases_ingested - BUDGET_LERP_LOW_BP) / (double)(BUDGET_LERP_HIGH_BP - BUDGET_LERP_LOW_BP);
    return (size_t)((double)MIN_NODE_BUDGET + fraction * (double)(DEFAULT_MAX_NUM_NODES - MIN_NODE_BUDGET));
}
inline uint64_t get_suffix_mask(size_t k) { return (uint64_t(1) << (2 * (k - 1))) - 1; }   // graph.rs:57-60

inline std::vector<uint32_t> compute_coverage_thresholds(uint32_t primer_count, uint32_t min_count) {   // mod.rs:405-432
    const uint32_t high = primer_count / COVERAGE_MULTIPLIER;
    std::vector<uint32_t> t;
    if (high <= min_count) {
        t.push_back(min_count);
    } else {
        const uint32_t step = (high - min_count) / (COVERAGE_STEPS - 1);
        for (uint32_t i = 0; i < COVERAGE_STEPS; i++) t.push_back(high >= i * step ? high - i * step : 0);
        t.back() = min_count;
    }
    t.erase(std::unique(t.begin(), t.end()), t.end());
    return t;
}

// KmerCounts::get_max_count / get_median_count (counting.rs:271-298)
inline std::pair<uint32_t, uint32_t> primer_counts_max_median(const PrimerKmers &pk) {
    std::vector<uint32_t> c;
    for (auto &e : pk) c.push_back(e.second);
    if (c.empty()) return {0, 0};
    std::sort(c.begin(), c.end());
    const size_t mid = c.size() / 2;
    return {c.back(), c.size() % 2 ? c[mid] : c[mid - 1] / 2 + c[mid] / 2};
}

inline std::pair<DiGraph, std::unordered_map<uint64_t, uint32_t>> create_seed_graph(const PrimerKmers &fwd, const PrimerKmers &rev,
                                                                                  size_t k) {   // graph.rs:192-278
    DiGraph g;
    std::unordered_map<uint64_t, uint32_t> lookup;
    const uint64_t mask = get_suffix_mask(k);
    for (auto &e : fwd) {   // PrimerKmers is ascending by k-mer already
        const uint64_t sub = e.first >> 2;
        auto it = lookup.find(sub);
        if (it != lookup.end()) g.nodes[it->second].is_start = true;
        else lookup[sub] = g.add_node(sub, true, false);
    }
    for (auto &e : rev) {
        const uint64_t sub = revcomp_kmer(e.first, (uint32_t)k) & mask;
        auto it = lookup.find(sub);
        if (it != lookup.end()) g.nodes[it->second].is_end = true;
        else lookup[sub] = g.add_node(sub, false, true);
    }
    return {std::move(g), std::move(lookup)};
}

enum { FORWARD = 0, REVERSE = 1 };

// Candidate k-mer lookups of the extension: one skm_lookup_batch per frontier wave (see pcr.py).
class WaveLookups {
  public:
    struct Answer { uint64_t kmer[4]; uint32_t count[4]; bool found[4]; };
    WaveLookups(KmerCounts &table, uint32_t view_min, size_t k, size_t budget = 4096)
        : table_(table), view_min_(view_min), shift_(2 * (k - 1)), mask_(get_suffix_mask(k)), budget_(budget) {}
    size_t calls = 0, kmers_asked = 0;

    Answer get(const DiGraph &g, const std::deque<std::pair<uint32_t, int>> &frontier, uint32_t node, int dir) {
        const uint64_t key = pack(g.nodes[node].sub_kmer, dir);
        auto it = ready_.find(key);
        if (it == ready_.end()) {
            std::vector<uint64_t> level{key};
            std::unordered_set<uint64_t> seen{key};
            for (auto &e : frontier) {
                const uint64_t q = pack(g.nodes[e.first].sub_kmer, e.second);
                if (!ready_.count(q) && seen.insert(q).second) level.push_back(q);
            }
            std::vector<uint64_t> todo;
            bool first = true;
            while (!level.empty() && (first || 4 * (todo.size() + level.size()) <= budget_)) {
                first = false;
                todo.insert(todo.end(), level.begin(), level.end());
                std::vector<uint64_t> next;
                for (uint64_t q : level) {
                    const int d = (int)(q & 1);
                    for (uint64_t b = 0; b < 4; b++) {
                        const uint64_t c = candidate(q >> 1, d, b);
                        const uint64_t child = pack(d == FORWARD ? (c & mask_) : (c >> 2), d);
                        if (!ready_.count(child) && seen.insert(child).second) next.push_back(child);
                    }
                }
                level.swap(next);
            }
            std::vector<uint64_t> kmers;
            kmers.reserve(todo.size() * 4);
            for (uint64_t q : todo)
                for (uint64_t b = 0; b < 4; b++) kmers.push_back(candidate(q >> 1, (int)(q & 1), b));
            std::vector<uint32_t> counts;
            std::vector<uint8_t> found;
            table_.lookup_found(kmers, view_min_, SKM_LOOKUP_EITHER, counts, found);
            calls++;
            kmers_asked += kmers.size();
            for (size_t i = 0; i < todo.size(); i++) {
                Answer a;
                for (int b = 0; b < 4; b++) {
                    a.kmer[b] = kmers[4 * i + b];
                    a.count[b] = counts[4 * i + b];
                    a.found[b] = found[4 * i + b] != 0;
                }
                ready_[todo[i]] = a;
            }
            if (ready_.size() > 1000000) {   // guesses that were never needed
                std::unordered_map<uint64_t, Answer> keep;
                keep[key] = ready_[key];
                for (auto &e : frontier) {
                    const uint64_t q = pack(g.nodes[e.first].sub_kmer, e.second);
                    auto f = ready_.find(q);
                    if (f != ready_.end()) keep[q] = f->second;
                }
                ready_.swap(keep);
            }
            it = ready_.find(key);
        }
        Answer a = it->second;
        ready_.erase(it);
        return a;
    }

  private:
    static uint64_t pack(uint64_t sub, int dir) { return (sub << 1) | (uint64_t)dir; }
    uint64_t candidate(uint64_t sub, int dir, uint64_t base) const {
        return dir == FORWARD ? ((sub << 2) | base) : ((base << shift_) | sub);
    }
    KmerCounts &table_;
    uint32_t view_min_;
    size_t shift_;
    uint64_t mask_;
    size_t budget_;
    std::unordered_map<uint64_t, Answer> ready_;
};

// graph.rs:322-527.  Returns found_path; `graph` and `lookup` are extended in place.
inline bool extend_graph(DiGraph &g, std::unordered_map<uint64_t, uint32_t> &lookup, KmerCounts &table, uint32_t view_min,
                         uint32_t min_count, const Params &p, size_t max_num_nodes, size_t k, size_t *n_calls = nullptr) {
    const uint64_t mask = get_suffix_mask(k);
    bool found_path = false;
    double median = (double)min_count;
    if (!median_f64(g.edge_counts(), median)) median = (double)min_count;
    size_t last_median_check = 0;
    std::deque<std::pair<uint32_t, int>> frontier;
    std::unordered_set<uint32_t> processed[2], added_by[2];
    for (uint32_t n : g.node_indices()) {
        if (g.nodes[n].is_start) {
            frontier.emplace_back(n, FORWARD);
            added_by[FORWARD].insert(n);
        }
        if (g.nodes[n].is_end) {
            frontier.emplace_back(n, REVERSE);
            added_by[REVERSE].insert(n);
        }
    }
    WaveLookups waves(table, view_min, k);
    while (!frontier.empty()) {
        const auto [node, d] = frontier.front();
        frontier.pop_front();
        if (!processed[d].insert(node).second) continue;
        const size_t n_nodes = g.n_nodes;
        if (n_nodes > max_num_nodes) break;
        if (n_nodes > last_median_check && n_nodes - last_median_check > EXTENSION_EVALUATION_FREQUENCY) {
            if (!median_f64(g.edge_counts(), median)) median = (double)min_count;
            last_median_check = n_nodes - n_nodes % EXTENSION_EVALUATION_FREQUENCY;
        }
        const uint64_t sub = g.nodes[node].sub_kmer;
        const WaveLookups::Answer a = waves.get(g, frontier, node, d);
        for (int b = 0; b < 4; b++) {
            if (!a.found[b] || a.count[b] < min_count) continue;
            const uint64_t kmer = a.kmer[b];
            const uint32_t count = a.count[b];
            const uint64_t new_sub = d == FORWARD ? (kmer & mask) : (kmer >> 2);
            if (new_sub == sub) continue;   // self loop
            auto ex = lookup.find(new_sub);
            if (ex != lookup.end()) {
                const uint32_t other = ex->second;
                const uint32_t from = d == FORWARD ? node : other, to = d == FORWARD ? other : node;
                if (!g.has_edge(from, to)) {
                    g.add_edge(from, to, count);
                    if (added_by[1 - d].count(other)) found_path = true;
                }
            } else {
                if ((double)count > median * p.high_coverage_ratio) continue;   // likely repetitive
                const uint32_t nn = g.add_node(new_sub);
                lookup[new_sub] = nn;
                added_by[d].insert(nn);
                if (d == FORWARD) g.add_edge(node, nn, count);
                else g.add_edge(nn, node, count);
                frontier.emplace_back(nn, d);
            }
        }
    }
    if (n_calls) *n_calls += waves.calls;
    return found_path;
}

inline void annotate_coverage_ratios(DiGraph &g) {   // graph.rs:531-544
    double median = 0.0;
    if (!median_f64(g.edge_counts(), median) || median <= 0.0) return;
    for (auto &e : g.edges)
        if (e.alive) e.coverage_ratio = (double)e.count / median;
}

// ---- pruning (pruning.rs) ----------------------------------------------------------------------
inline size_t tip_length_backward(const DiGraph &g, uint32_t node) {   // :99-124
    size_t length = 0;
    uint32_t cur = node;
    for (;;) {
        length++;
        if (g.in[cur].size() != 1) break;
        const uint32_t parent = g.edges[g.in[cur][0]].src;
        if (g.out[parent].size() > 1 || g.nodes[parent].is_start) break;
        cur = parent;
    }
    return length;
}
inline size_t tip_length_forward(const DiGraph &g, uint32_t node) {   // :128-149
    size_t length = 0;
    uint32_t cur = node;
    for (;;) {
        length++;
        if (g.out[cur].size() != 1) break;
        const uint32_t child = g.edges[g.out[cur][0]].dst;
        if (g.in[child].size() > 1 || g.nodes[child].is_end) break;
        cur = child;
    }
    return length;
}
inline void remove_low_coverage_tips(DiGraph &g, size_t k, double tip_coverage_fraction) {   // :19-95
    double median = 1.0;
    if (!median_f64(g.edge_counts(), median)) median = 1.0;
    const double min_tip_count = std::max(median * tip_coverage_fraction, 1.0);
    size_t removed = 1;
    while (removed > 0) {
        removed = 0;
        std::vector<uint32_t> doomed;
        for (uint32_t n : g.node_indices()) {
            if (g.nodes[n].is_start || g.nodes[n].is_end) continue;
            const bool no_out = g.out[n].empty(), no_in = g.in[n].empty();
            if (!no_out && !no_in) continue;
            bool keep = false;
            if (no_out) {
                uint32_t mx = 0;
                for (uint32_t e : g.in[n]) mx = std::max(mx, g.edges[e].count);
                if (tip_length_backward(g, n) >= k || (double)mx >= min_tip_count) keep = true;
            }
            if (!keep && no_in) {
                uint32_t mx = 0;
                for (uint32_t e : g.out[n]) mx = std::max(mx, g.edges[e].count);
                if (tip_length_forward(g, n) >= k || (double)mx >= min_tip_count) keep = true;
            }
            if (!keep) doomed.push_back(n);
        }
        for (uint32_t n : doomed) {
            g.remove_node(n);
            removed++;
        }
    }
}
inline void reachability_pruning(DiGraph &g) {   // :166-216
    std::unordered_set<uint32_t> fwd, bwd;
    std::vector<uint32_t> stack;
    for (uint32_t n : g.node_indices())
        if (g.nodes[n].is_start) stack.push_back(n);
    while (!stack.empty()) {
        const uint32_t n = stack.back();
        stack.pop_back();
        if (fwd.insert(n).second)
            for (uint32_t e : g.out[n]) stack.push_back(g.edges[e].dst);
    }
    for (uint32_t n : g.node_indices())
        if (g.nodes[n].is_end) stack.push_back(n);
    while (!stack.empty()) {
        const uint32_t n = stack.back();
        stack.pop_back();
        if (bwd.insert(n).second)
            for (uint32_t e : g.in[n]) stack.push_back(g.edges[e].src);
    }
    for (uint32_t n : g.node_indices())
        if (!fwd.count(n) || !bwd.count(n)) g.remove_node(n);
}

// ---- paths (paths.rs) ----------------------------------------------------------------------------
struct Child { uint32_t node, edge; double score; };
inline std::vector<Child> sorted_children(const DiGraph &g, uint32_t node) {   // :42-65
    std::vector<Child> c;
    for (uint32_t e : g.out_edges(node)) c.push_back(Child{g.edges[e].dst, e, (double)g.edges[e].count});
    std::stable_sort(c.begin(), c.end(), [](const Child &a, const Child &b) { return a.score < b.score; });
    return c;   // ascending: back() is the highest score
}
constexpr uint32_t NO_EDGE = 0xffffffffu;
using Path = std::vector<std::pair<uint32_t, uint32_t>>;   // (node, edge into it or NO_EDGE)

inline std::vector<Path> get_assembly_paths(const DiGraph &g, size_t k, const Params &p) {   // :79-196
    const size_t min_path_nodes = p.min_length <= k ? 1 : p.min_length - k + 2;
    const size_t max_path_nodes = p.max_length <= k ? 1 : p.max_length - k + 2;
    std::vector<Path> all;
    for (uint32_t start : g.node_indices()) {
        if (!g.nodes[start].is_start) continue;
        size_t paths_from_start = 0, states = 0;
        Path path{{start, NO_EDGE}};
        std::unordered_map<uint32_t, size_t> visits{{start, 1}};
        std::vector<std::vector<Child>> stack{sorted_children(g, start)};
        for (;;) {
            if (paths_from_start >= p.max_paths_per_pair || states >= p.max_dfs_states) break;
            auto &frame = stack.back();
            if (!frame.empty()) {
                const Child c = frame.back();
                frame.pop_back();
                states++;
                if (visits[c.node] >= p.max_node_visits) continue;
                path.emplace_back(c.node, c.edge);
                visits[c.node]++;
                if (g.nodes[c.node].is_end && path.size() >= min_path_nodes) {
                    all.push_back(path);
                    paths_from_start++;
                    visits[c.node]--;
                    path.pop_back();
                    continue;
                }
                if (path.size() >= max_path_nodes) {
                    visits[c.node]--;
                    path.pop_back();
                    continue;
                }
                stack.push_back(sorted_children(g, c.node));
            } else {
                stack.pop_back();
                if (stack.empty()) break;
                visits[path.back().first]--;
                path.pop_back();
            }
        }
    }
    return all;
}

struct PathScore {   // pcr/mod.rs:58-112, the branch without threading data
    uint32_t kmer_min_count;
    double kmer_median_count, coverage_cv, max_coverage_ratio;
    double composite() const {
        const double cv_penalty = coverage_cv > 1.0 ? 1.0 / coverage_cv : 1.0;
        const double repeat_penalty = max_coverage_ratio > 5.0 ? 5.0 / max_coverage_ratio : 1.0;
        return kmer_median_count * cv_penalty * repeat_penalty * 1.0;
    }
};
struct Record {
    std::string id, desc, seq;
    PathScore score;
};

inline std::string fixed2(double x) {   // Rust's {:.2}
    char b[64];
    std::snprintf(b, sizeof b, "%.2f", x);
    return b;
}
inline std::string f64_display(double x) {   // Rust's {} for an f64: shortest round trip, no exponent, no ".0"
    if (x == std::floor(x) && std::fabs(x) < 1e16) {
        char b[32];
        std::snprintf(b, sizeof b, "%lld", (long long)x);
        return b;
    }
    for (int prec = 1; prec <= 17; prec++) {
        char b[64];
        std::snprintf(b, sizeof b, "%.*g", prec, x);
        if (std::strtod(b, nullptr) == x) return b;
    }
    return std::to_string(x);
}

inline std::vector<Record> generate_sequences_from_paths(const DiGraph &g, const std::vector<Path> &paths, size_t k,
                                                         const std::string &sample, const Params &p,
                                                         size_t amplicon_index = 0) {   // paths.rs:200-377
    std::vector<Record> out;
    for (const Path &path : paths) {
        std::string seq;
        std::vector<uint64_t> counts;
        std::vector<uint32_t> path_edges;
        for (auto &[node, edge] : path) {
            const uint64_t sub = g.nodes[node].sub_kmer;
            if (seq.empty()) {
                seq = kmer_to_seq(sub, (uint32_t)(k - 1));
            } else {
                seq.push_back(kmer_last_base(sub));
                counts.push_back(g.edges[edge].count);
                path_edges.push_back(edge);
            }
        }
        if (seq.size() < p.min_length || counts.empty()) continue;
        const double mean = compute_mean(counts), median = compute_median(counts);
        const uint64_t cmin = *std::min_element(counts.begin(), counts.end());
        const uint64_t cmax = *std::max_element(counts.begin(), counts.end());
        double cv = 0.0;
        if (mean > 0.0) {
            double var = 0.0;
            for (uint64_t c : counts) var += ((double)c - mean) * ((double)c - mean);
            cv = std::sqrt(var / (double)counts.size()) / mean;
        }
        double max_ratio = 0.0;
        for (uint32_t e : path_edges) max_ratio = std::max(max_ratio, g.edges[e].coverage_ratio);
        Record r;
        r.score = PathScore{(uint32_t)cmin, median, cv, max_ratio};
        r.id = sample + "_" + p.gene_name + "_" + std::to_string(amplicon_index);
        r.desc = "sample=" + sample + " gene=" + p.gene_name + " product=" + std::to_string(amplicon_index) +
                 " length=" + std::to_string(seq.size()) + " kmer_count_mean=" + fixed2(mean) +
                 " kmer_count_median=" + f64_display(median) + " kmer_count_min=" + std::to_string(cmin) +
                 " kmer_count_max=" + std::to_string(cmax) + " score=" + fixed2(r.score.composite());
        r.seq = std::move(seq);
        amplicon_index++;
        out.push_back(std::move(r));
    }
    return out;
}

// edit distance if <= k, else -1 (bio::alignment::distance::simd::bounded_levenshtein)
inline long bounded_levenshtein(const std::string &x, const std::string &y, size_t k) {
    const std::string &a = x.size() <= y.size() ? x : y, &b = x.size() <= y.size() ? y : x;
    if (b.size() - a.size() > k) return -1;
    const size_t big = k + 1, m = b.size();
    std::vector<size_t> prev(m + 1, big), cur(m + 1, big);
    for (size_t j = 0; j <= std::min(m, k); j++) prev[j] = j;
    for (size_t i = 1; i <= a.size(); i++) {
        const size_t lo = i > k ? i - k : 0, hi = std::min(m, i + k);
        size_t best_row = big;
        for (size_t j = lo; j <= hi; j++) {
            size_t v;
            if (j == 0) {
                v = i;
            } else {
                v = prev[j - 1] + (a[i - 1] != b[j - 1] ? 1 : 0);
                v = std::min(v, prev[j] + 1);
                if (j > lo) v = std::min(v, cur[j - 1] + 1);
            }
            cur[j] = std::min(v, big);
            best_row = std::min(best_row, cur[j]);
        }
        if (best_row > k) return -1;
        if (lo > 0) cur[lo - 1] = big;
        if (hi < m) cur[hi + 1] = big;
        prev.swap(cur);
    }
    return prev[m] <= k ? (long)prev[m] : -1;
}

inline std::vector<Record> sort_and_deduplicate(std::vector<Record> recs, const Params &p) {   // paths.rs:381-428
    std::stable_sort(recs.begin(), recs.end(), [](const Record &a, const Record &b) {
        const double sa = a.score.composite(), sb = b.score.composite();
        if (sa != sb) return sa > sb;
        return a.seq < b.seq;
    });
    std::vector<Record> kept;
    for (auto &r : recs) {
        bool dup = false;
        for (auto &q : kept)
            if (bounded_levenshtein(r.seq, q.seq, p.dedup_edit_threshold) >= 0) {
                dup = true;
                break;
            }
        if (!dup) kept.push_back(r);
    }
    if (kept.size() > MAX_NUM_AMPLICONS) kept.resize(MAX_NUM_AMPLICONS);
    return kept;
}

// ---- the pipeline -------------------------------------------------------------------------------
struct Outcome {   // pcr/mod.rs:286-291
    std::vector<Record> records;
    std::string failure_reason;   // empty when there are records
    size_t nodes = 0, edges = 0, lookup_calls = 0;
};

inline Outcome do_pcr(KmerCounts &table, const std::string &sample, const Params &p,
                      size_t max_num_nodes = DEFAULT_MAX_NUM_NODES, uint32_t view_min_count = 2) {   // mod.rs:434-795
    const size_t k = table.get_k();
    Outcome out;
    auto primers = get_primer_kmers(table, p);
    const PrimerKmers &fwd = primers.first, &rev = primers.second;
    if (fwd.empty() || rev.empty()) {
        out.failure_reason = std::string(fwd.empty() && rev.empty() ? "forward and reverse primers"
                                         : fwd.empty()              ? "forward primer"
                                                                    : "reverse primer") + " not found";
        return out;
    }
    auto seed = create_seed_graph(fwd, rev, k);
    const uint32_t max_primer = std::min(primer_counts_max_median(fwd).first, primer_counts_max_median(rev).first);
    std::string failure = "no path found";
    bool found_signal = false;
    DiGraph current = seed.first;
    for (uint32_t min_count : compute_coverage_thresholds(max_primer, p.min_count)) {
        DiGraph fresh = seed.first;
        std::unordered_map<uint64_t, uint32_t> lookup;
        for (uint32_t n : fresh.node_indices()) lookup[fresh.nodes[n].sub_kmer] = n;
        const bool found = extend_graph(fresh, lookup, table, view_min_count, min_count, p, max_num_nodes, k, &out.lookup_calls);
        current = std::move(fresh);
        if (found) {
            found_signal = true;
            break;
        }
    }
    if (current.n_nodes >= max_num_nodes) failure = "node budget exceeded";
    out.nodes = current.n_nodes;
    out.edges = current.n_edges;
    std::vector<Record> records;
    if (found_signal) {
        DiGraph pruned = current;
        remove_low_coverage_tips(pruned, k, p.tip_coverage_fraction);
        reachability_pruning(pruned);
        annotate_coverage_ratios(pruned);
        auto paths = get_assembly_paths(pruned, k, p);
        if (!paths.empty()) records = generate_sequences_from_paths(pruned, paths, k, sample, p, 0);
    }
    if (records.empty()) {
        out.failure_reason = failure;
        return out;
    }
    size_t i = 0;
    for (auto &r : sort_and_deduplicate(std::move(records), p)) {   // renumber after dedup (mod.rs:759-790)
        Record q = r;
        q.id = sample + "_" + p.gene_name + "_" + std::to_string(i);
        std::string desc, field;
        size_t pos = 0;
        while (pos <= r.desc.size()) {
            const size_t sp = r.desc.find(' ', pos);
            field = r.desc.substr(pos, sp == std::string::npos ? std::string::npos : sp - pos);
            if (!field.empty()) {
                if (field.rfind("product=", 0) == 0) field = "product=" + std::to_string(i);
                desc += (desc.empty() ? "" : " ") + field;
            }
            if (sp == std::string::npos) break;
            pos = sp + 1;
        }
        q.desc = desc;
        out.records.push_back(std::move(q));
        i++;
    }
    return out;
}

inline void write_fasta(FILE *f, const std::vector<Record> &records) {   // io.rs:144-158
    for (auto &r : records) {
        std::fprintf(f, ">%s %s\n", r.id.c_str(), r.desc.c_str());
        for (size_t i = 0; i < r.seq.size(); i += FASTA_LINE_WIDTH) std::fprintf(f, "%s\n", r.seq.substr(i, FASTA_LINE_WIDTH).c_str());
    }
}

struct GeneResult {   // stats.rs:12-23
    std::string gene_name, status, failure_reason;
    std::vector<size_t> product_lengths;
};

// stats.rs:49-155: one {directory}{sample}_{gene}.fasta per gene with products
inline std::vector<GeneResult> run_pcr(KmerCounts &table, std::vector<Params> runs, const std::string &sample,
                                       const std::string &directory, uint32_t min_kmer_count = 2,
                                       size_t max_nodes = DEFAULT_MAX_NUM_NODES) {
    std::vector<GeneResult> results;
    for (auto &p : runs) {
        if (p.min_count < min_kmer_count) p.min_count = min_kmer_count;   // cli.rs:556-570
        Outcome o = do_pcr(table, sample, p, max_nodes, min_kmer_count);
        GeneResult r;
        r.gene_name = p.gene_name;
        if (!o.records.empty()) {
            const std::string path = directory + sample + "_" + p.gene_name + ".fasta";
            FILE *f = std::fopen(path.c_str(), "w");
            if (!f) throw Error(SKM_ERR_INVALID_ARG, "Failed to create FASTA file: " + path);
            write_fasta(f, o.records);
            std::fclose(f);
            r.status = "success";
            for (auto &rec : o.records) r.product_lengths.push_back(rec.seq.size());
        } else {
            r.status = "fail";
            r.failure_reason = o.failure_reason.empty() ? "unknown (no reason reported by PCR pipeline)" : o.failure_reason;
        }
        results.push_back(std::move(r));
    }
    return results;
}

// ---- --pcr-primers "key=value,..." (cli.rs:12-140) and validation (pcr/mod.rs:296-401) -----------
inline Params parse_pcr_primers_string(const std::string &s) {
    if (s.empty()) throw Error(SKM_ERR_INVALID_ARG, "Invalid empty primer specification");
    Params p;
    p.gene_name.clear();
    std::set<std::string> seen;
    auto upper = [](std::string v) {
        for (char &c : v) c = (char)std::toupper((unsigned char)c);
        return v;
    };
    auto number = [](const std::string &key, const std::string &v) -> unsigned long long {
        if (v.empty() || v.find_first_not_of("0123456789") != std::string::npos)
            throw Error(SKM_ERR_INVALID_ARG, "Invalid value for " + key + ": " + v);
        return std::strtoull(v.c_str(), nullptr, 10);
    };
    size_t pos = 0;
    for (;;) {
        const size_t comma = s.find(',', pos);
        const std::string item = s.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
        const size_t eq = item.find('=');
        if (eq == std::string::npos)
            throw Error(SKM_ERR_INVALID_ARG, "Invalid parameter (should be key=value): '" + item +
                                                 "'\nCommas are not allowed in field values. Use --pcr-panel-file with a YAML panel for complex metadata.");
        std::string key = item.substr(0, eq);
        for (char &c : key) c = (char)std::tolower((unsigned char)c);
        const std::string value = item.substr(eq + 1);
        if (!seen.insert(key).second)
            throw Error(SKM_ERR_INVALID_ARG, "Duplicate parameter '" + key + "' in primer specification '" + s +
                                                 "'. Each key may appear at most once.");
        if (key == "name") p.gene_name = value;
        else if (key == "forward") p.forward_seq = upper(value);
        else if (key == "reverse") p.reverse_seq = upper(value);
        else if (key == "max-length") p.max_length = number(key, value);
        else if (key == "min-length") p.min_length = number(key, value);
        else if (key == "min-count") p.min_count = (uint32_t)number(key, value);
        else if (key == "mismatches") p.mismatches = number(key, value);
        else if (key == "trim") p.trim = number(key, value);
        else if (key == "citation" || key == "notes") {
        } else if (key == "dedup-edit-threshold") p.dedup_edit_threshold = (uint32_t)number(key, value);
        else throw Error(SKM_ERR_INVALID_ARG, "Unexpected parameter: " + key);
        if (comma == std::string::npos) break;
        pos = comma + 1;
    }
    return p;
}

inline std::vector<std::pair<std::string, std::string>> validate_pcr_params(const Params &p) {
    std::vector<std::pair<std::string, std::string>> e;
    if (p.forward_seq.size() < 2) e.emplace_back("Forward primer sequence is too short: '" + p.forward_seq + "'", "Primer sequences must be at least 2 bases");
    if (p.reverse_seq.size() < 2) e.emplace_back("Reverse primer sequence is too short: '" + p.reverse_seq + "'", "Primer sequences must be at least 2 bases");
    for (int which = 0; which < 2; which++) {
        const std::string &seq = which ? p.reverse_seq : p.forward_seq;
        if (seq.size() < 2) continue;
        std::string bad;
        for (char c : seq)
            if (!iupac_bases(c)) bad += (bad.empty() ? "" : ", ") + std::string(1, c);
        if (!bad.empty())
            e.emplace_back("Invalid nucleotide(s) " + bad + " in " + (which ? "reverse" : "forward") + " primer " + seq,
                           "Valid characters: A C G T R Y W S M K B D H V N");
    }
    if (p.min_length > p.max_length)
        e.emplace_back("min-length (" + std::to_string(p.min_length) + ") is greater than max-length (" + std::to_string(p.max_length) + ")",
                       "Swap the values or adjust the range");
    if (p.min_count < 2) e.emplace_back("min-count is " + std::to_string(p.min_count) + ", must be at least 2", "Set min-count to at least 2");
    if (p.max_length == 0) e.emplace_back("max-length is 0", "Set max-length to a positive value");
    if (p.gene_name.empty()) e.emplace_back("Gene name is empty", "Provide a unique name for the primer pair via the 'name' field");
    if (p.forward_seq == p.reverse_seq && p.forward_seq.size() >= 2)
        e.emplace_back("Forward and reverse primers are identical: " + p.forward_seq, "Check that forward and reverse sequences are not swapped");
    return e;
}

}  // namespace pcr
}  // namespace skm
