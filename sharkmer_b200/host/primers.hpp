// primers.hpp — primer preprocessing and primer k-mer discovery for sPCR over the device table:
// what get_primer_kmers computes (caseywdunn/sharkmer v3.1.0, src/pcr/primers.rs:448-478), with
// its full-table scans (find_oligos_in_kmers, :163-226, one per primer direction and mismatch
// level) running on the GPU behind skm_scan_oligos.  C++ twin of sharkmer_b200/primers.py.
//
// Primer variants are 2-bit packed integers (the form the device takes), kept as sorted unique
// vectors: IUPAC expansion is a product over per-position base lists, a mismatch level is the
// Hamming-1 neighbourhood of everything seen so far minus what was seen.
// Reference behaviour kept: trim >= k clamps to k-1 and a longer primer keeps its 3' end
// (:236-263); more than 10 000 resolved variants is an error (:268-277); mismatches clamp to the
// primer length (:279-280); levels are disjoint (:283-299); per level the matches not found at a
// lower level fill what is left of max_primer_kmers, by count descending then k-mer ascending
// (:376-446); defaults trim 15, mismatches 2, min_count 2 (cli.rs:22-24), cap 40 (pcr/mod.rs:281).
#pragma once

#include <algorithm>
#include <iterator>
#include <string>
#include <utility>
#include <vector>

#include "kmer.hpp"

namespace skm {

struct PCRParams {
    std::string forward_seq, reverse_seq, gene_name = "gene";
    uint32_t min_count = 2;
    size_t mismatches = 2, trim = 15, max_primer_kmers = 40;
};

struct Oligo {
    size_t length;
    uint64_t kmer;
};

constexpr size_t MAX_RESOLVED_VARIANTS = 10000;

inline Oligo string_to_oligo(const std::string &seq) {  // primers.rs:33-54
    if (seq.size() > 32)
        throw Error(SKM_ERR_INVALID_ARG, "Oligo sequence length " + std::to_string(seq.size()) + " exceeds maximum of 32 bases");
    uint64_t v = 0;
    for (char c : seq) {
        uint64_t b;
        switch (c) {
        case 'A': b = 0; break;
        case 'C': b = 1; break;
        case 'G': b = 2; break;
        case 'T': b = 3; break;
        default: throw Error(SKM_ERR_INVALID_ARG, std::string("Invalid nucleotide ") + c + " in " + seq);
        }
        v = (v << 2) | b;
    }
    return Oligo{seq.size(), v};
}

inline const char *iupac_bases(char c) {  // primers.rs:11-30, 63-76; nullptr = not a nucleotide code
    switch (c) {
    case 'A': return "A";
    case 'C': return "C";
    case 'G': return "G";
    case 'T': return "T";
    case 'R': return "AG";
    case 'Y': return "CT";
    case 'S': return "GC";
    case 'W': return "AT";
    case 'K': return "GT";
    case 'M': return "AC";
    case 'B': return "CGT";
    case 'D': return "AGT";
    case 'H': return "ACT";
    case 'V': return "ACG";
    case 'N': return "ACGT";
    default: return nullptr;
    }
}

inline std::string trimmed_primer(const PCRParams &p, bool reverse, size_t k) {
    const std::string &primer = reverse ? p.reverse_seq : p.forward_seq;
    const size_t trim = std::min(p.trim, k - 1);
    return primer.size() > trim ? primer.substr(primer.size() - trim) : primer;
}

inline void sort_unique(std::vector<uint64_t> &v) {
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
}

// every ambiguity-free reading of `primer`, packed, ascending
inline std::vector<uint64_t> resolve_primer(const std::string &primer) {
    if (primer.size() > 32)
        throw Error(SKM_ERR_INVALID_ARG, "Oligo sequence length " + std::to_string(primer.size()) + " exceeds maximum of 32 bases");
    if (primer.empty()) return {};
    size_t n = 1;
    for (char c : primer) {
        const char *b = iupac_bases(c);
        if (!b) throw Error(SKM_ERR_INVALID_ARG, std::string("Invalid nucleotide ") + c + " in " + primer);
        n *= std::char_traits<char>::length(b);
        if (n > (size_t(1) << 40)) n = size_t(1) << 40;  // no overflow; far past the limit anyway
    }
    if (n > MAX_RESOLVED_VARIANTS)
        throw Error(SKM_ERR_INVALID_ARG, "Primer " + primer + " has too many ambiguous bases: " + std::to_string(n) +
                                             " resolved variants exceeds limit of " + std::to_string(MAX_RESOLVED_VARIANTS) +
                                             ". Reduce ambiguity or use a more specific primer.");
    std::vector<uint64_t> out{0};
    for (char c : primer) {
        const std::string bases = iupac_bases(c);
        std::vector<uint64_t> next;
        next.reserve(out.size() * bases.size());
        for (uint64_t v : out)
            for (char b : bases) next.push_back((v << 2) | string_to_oligo(std::string(1, b)).kmer);
        out.swap(next);
    }
    sort_unique(out);
    return out;
}

// everything within one substitution of any element (the elements included)
inline std::vector<uint64_t> hamming1(const std::vector<uint64_t> &variants, size_t length) {
    std::vector<uint64_t> out;
    out.reserve(variants.size() * (3 * length + 1));
    for (uint64_t v : variants) {
        out.push_back(v);
        for (size_t p = 0; p < length; p++) {
            const uint64_t cleared = v & ~(uint64_t(3) << (2 * p));
            for (uint64_t b = 0; b < 4; b++) out.push_back(cleared | (b << (2 * p)));
        }
    }
    sort_unique(out);
    return out;
}

struct PrimerLevels {
    size_t length = 0;                           // bases of the trimmed primer
    std::vector<std::vector<uint64_t>> levels;   // levels[m]: variants first reached with m mismatches
};

inline PrimerLevels preprocess_primer_by_mismatch(const PCRParams &p, bool reverse, size_t k) {  // primers.rs:231-307
    const std::string primer = trimmed_primer(p, reverse, k);
    PrimerLevels r;
    r.length = primer.size();
    std::vector<uint64_t> seen = resolve_primer(primer);
    r.levels.push_back(seen);
    for (size_t m = 0; m < std::min(p.mismatches, primer.size()); m++) {
        std::vector<uint64_t> ball = hamming1(seen, primer.size()), fresh;
        std::set_difference(ball.begin(), ball.end(), seen.begin(), seen.end(), std::back_inserter(fresh));
        r.levels.push_back(std::move(fresh));
        seen.swap(ball);
    }
    return r;
}

using PrimerKmers = std::vector<std::pair<uint64_t, uint32_t>>;  // (k-mer, count), ascending k-mer

// discover_primer_kmers_by_round, primers.rs:376-446
inline PrimerKmers discover_primer_kmers(KmerCounts &table, const PrimerLevels &pl, uint32_t min_count, size_t cap) {
    PrimerKmers result;
    for (const auto &oligos : pl.levels) {
        if (result.size() >= cap) break;
        if (oligos.empty()) continue;
        auto found = table.scan_oligos(oligos, (uint32_t)pl.length, min_count);
        PrimerKmers fresh;
        for (size_t i = 0; i < found.first.size(); i++) {
            const uint64_t km = found.first[i];
            const bool have = std::any_of(result.begin(), result.end(), [&](const auto &e) { return e.first == km; });
            if (!have) fresh.emplace_back(km, found.second[i]);
        }
        std::sort(fresh.begin(), fresh.end(), [](const auto &a, const auto &b) {
            return a.second != b.second ? a.second > b.second : a.first < b.first;
        });
        const size_t take = std::min(fresh.size(), cap - result.size());
        result.insert(result.end(), fresh.begin(), fresh.begin() + take);
    }
    std::sort(result.begin(), result.end());
    return result;
}

// get_primer_kmers, primers.rs:448-478.  The scan threshold is params.min_count alone: the reference's
// FilteredKmerCounts::iter() yields every entry whatever the view's threshold (counting.rs:343-349).
inline std::pair<PrimerKmers, PrimerKmers> get_primer_kmers(KmerCounts &table, const PCRParams &p) {
    const size_t k = table.get_k();
    PrimerKmers fwd = discover_primer_kmers(table, preprocess_primer_by_mismatch(p, false, k), p.min_count, p.max_primer_kmers);
    PrimerKmers rev = discover_primer_kmers(table, preprocess_primer_by_mismatch(p, true, k), p.min_count, p.max_primer_kmers);
    return {std::move(fwd), std::move(rev)};
}

}  // namespace skm
