// kmer.hpp — C++ host mirror of the reference's k-mer API over the C ABI.
//
// The reference (caseywdunn/sharkmer v3.1.0) is Rust; Rust is not available in
// this build environment, so the host side above include/sharkmer_b200.h is
// C++.  Class and method names follow src/kmer/mod.rs:10-17 so that
// src/io.rs / src/stats.rs / src/pcr call sites map one to one:
//
//   Chunk::ingest_seq            src/kmer/chunk.rs:25-30      (batched: Batcher below)
//   KmerCounts::{insert, extend, get_count, get_canonical, get_canonical_count,
//                len, get_n_kmers, get_n_unique_kmers, filtered_view}
//                                src/kmer/counting.rs:113-312
//   FilteredKmerCounts           src/kmer/counting.rs:316-350
//   Histogram::get_vector        src/kmer/histogram.rs:125-134
//
// Errors are thrown as skm::Error (the reference returns anyhow::Result and
// aborts the run on the first error, src/io.rs:357).
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sharkmer_b200.h"

namespace skm {

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string &m) : std::runtime_error(m), code(c) {}
};

class Engine {
  public:
    Engine(uint32_t k, uint32_t chunks, uint64_t histo_max = 10000, uint64_t capacity_hint = 0,
           int32_t device = -1, uint32_t insert_mode = SKM_INSERT_AUTO) {
        skm_params p{};
        p.struct_size = sizeof p;
        p.k = k;
        p.chunks = chunks;
        p.insert_mode = insert_mode;
        p.histo_max = histo_max;
        p.capacity_hint = capacity_hint;
        p.device = device;
        p.n_ranks = 1;
        int32_t rc = skm_create(&p, &h_);
        if (rc) {
            std::string m = h_ ? skm_last_error(h_) : "skm_create failed";
            if (h_) skm_destroy(h_);
            h_ = nullptr;
            throw Error(rc, m);
        }
        k_ = k;
        chunks_ = chunks;
        histo_max_ = histo_max;
    }
    ~Engine() {
        if (h_) skm_destroy(h_);
    }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;

    skm_ctx *raw() { return h_; }
    uint32_t k() const { return k_; }
    uint32_t chunks() const { return chunks_; }
    uint32_t n_chunks() const { return chunks_ == 0 ? 1 : chunks_; }
    uint64_t histo_max() const { return histo_max_; }

    void check(int32_t rc) const {
        if (rc) throw Error(rc, skm_last_error(h_));
    }
    void *pinned_alloc(size_t n) {
        void *p = nullptr;
        check(skm_pinned_alloc(h_, n, &p));
        return p;
    }
    void pinned_free(void *p) { skm_pinned_free(h_, p); }
    void ingest_batch(uint32_t chunk, const uint8_t *seqs, uint64_t n, uint32_t flags = 0) {
        check(skm_ingest_batch(h_, chunk, seqs, n, flags));
    }
    void finalize() { check(skm_finalize(h_)); }
    std::vector<uint64_t> histogram(uint32_t chunk_i) {
        std::vector<uint64_t> v(histo_max_ + 2);
        check(skm_histogram(h_, chunk_i, v.data(), v.size()));
        return v;
    }
    skm_totals totals() {
        skm_totals t{};
        check(skm_totals_get(h_, &t));
        return t;
    }
    skm_totals chunk_totals(uint32_t c) {
        skm_totals t{};
        check(skm_chunk_totals(h_, c, &t));
        return t;
    }
    skm_stage_ms stage_times() {
        skm_stage_ms t{};
        check(skm_stage_times(h_, &t));
        return t;
    }

  private:
    skm_ctx *h_ = nullptr;
    uint32_t k_ = 0, chunks_ = 0;
    uint64_t histo_max_ = 0;
};

class FilteredKmerCounts;

// src/kmer/counting.rs:113-312 over the device table of an Engine.
class KmerCounts {
  public:
    explicit KmerCounts(Engine &e) : e_(e) {}
    uint32_t get_k() const { return e_.k(); }
    void insert(uint64_t kmer, uint32_t count) { e_.check(skm_insert_counts(e_.raw(), &kmer, &count, 1)); }
    void extend(KmerCounts &other) {
        if (other.get_k() != get_k()) throw Error(SKM_ERR_K_MISMATCH, "Cannot extend KmerCounts with different k");
        auto kv = other.iter();
        e_.check(skm_insert_counts(e_.raw(), kv.first.data(), kv.second.data(), kv.first.size()));
    }
    uint64_t len() {
        uint64_t n = 0;
        e_.check(skm_table_len(e_.raw(), &n));
        return n;
    }
    bool is_empty() { return len() == 0; }
    uint64_t get_n_kmers() { return e_.totals().n_kmers; }
    uint64_t get_n_unique_kmers() { return len(); }
    uint32_t get_count(uint64_t kmer) { return lookup1(kmer, 0, SKM_LOOKUP_EXACT).first; }
    bool contains(uint64_t kmer) { return lookup1(kmer, 0, SKM_LOOKUP_EXACT).second; }
    uint32_t get_canonical_count(uint64_t kmer) { return lookup1(kmer, 0, SKM_LOOKUP_CANONICAL).first; }
    // Some(count) / None as (count, found)
    std::pair<uint32_t, bool> get_canonical(uint64_t kmer) { return lookup1(kmer, 0, SKM_LOOKUP_EITHER); }
    // (keys, counts); sorted => ascending k-mer order
    std::pair<std::vector<uint64_t>, std::vector<uint32_t>> iter(bool sorted = false) {
        uint64_t n = len(), got = 0;
        std::vector<uint64_t> keys(n);
        std::vector<uint32_t> counts(n);
        e_.check(skm_export(e_.raw(), keys.data(), counts.data(), n, sorted ? 1 : 0, &got));
        keys.resize(got);
        counts.resize(got);
        return {std::move(keys), std::move(counts)};
    }
    std::vector<uint32_t> lookup(const std::vector<uint64_t> &kmers, uint32_t min_count, int32_t mode) {
        std::vector<uint32_t> c(kmers.size());
        e_.check(skm_lookup_batch(e_.raw(), kmers.data(), kmers.size(), min_count, mode, c.data(), nullptr));
        return c;
    }
    // find_oligos_in_kmers (src/pcr/primers.rs:163-226) as one table pass on the device: table
    // k-mers with count >= min_count that start with one of the (unshifted, 2-bit) oligos, or whose
    // reverse complement does (then the reverse complement is reported); ascending k-mer order
    std::pair<std::vector<uint64_t>, std::vector<uint32_t>> scan_oligos(const std::vector<uint64_t> &oligos,
                                                                        uint32_t oligo_length, uint32_t min_count) {
        uint64_t n = 0;
        e_.check(skm_scan_oligos(e_.raw(), oligos.data(), oligos.size(), oligo_length, min_count, nullptr, nullptr, 0, &n));
        std::vector<uint64_t> keys(n);
        std::vector<uint32_t> counts(n);
        if (n) e_.check(skm_scan_oligos(e_.raw(), oligos.data(), oligos.size(), oligo_length, min_count, keys.data(),
                                        counts.data(), n, &n));
        keys.resize(n);
        counts.resize(n);
        return {std::move(keys), std::move(counts)};
    }
    FilteredKmerCounts filtered_view(uint32_t min_count);
    Engine &engine() { return e_; }

  private:
    friend class FilteredKmerCounts;
    std::pair<uint32_t, bool> lookup1(uint64_t kmer, uint32_t min_count, int32_t mode) {
        uint32_t c = 0;
        uint8_t f = 0;
        e_.check(skm_lookup_batch(e_.raw(), &kmer, 1, min_count, mode, &c, &f));
        return {c, f != 0};
    }
    Engine &e_;
};

// src/kmer/counting.rs:316-350
class FilteredKmerCounts {
  public:
    FilteredKmerCounts(KmerCounts &inner, uint32_t min_count) : inner_(inner), min_(min_count) {}
    uint32_t get_k() const { return inner_.get_k(); }
    std::pair<uint32_t, bool> get_canonical(uint64_t kmer) { return inner_.lookup1(kmer, min_, SKM_LOOKUP_EITHER); }
    uint32_t get_canonical_count(uint64_t kmer) { return inner_.lookup1(kmer, min_, SKM_LOOKUP_CANONICAL).first; }
    // batched form for graph extension (src/pcr/graph.rs:419-430 does 4 of these per node)
    std::vector<uint32_t> get_canonical_counts(const std::vector<uint64_t> &kmers) {
        return inner_.lookup(kmers, min_, SKM_LOOKUP_CANONICAL);
    }
    std::pair<std::vector<uint64_t>, std::vector<uint32_t>> iter() { return inner_.iter(); }

  private:
    KmerCounts &inner_;
    uint32_t min_;
};

inline FilteredKmerCounts KmerCounts::filtered_view(uint32_t min_count) { return FilteredKmerCounts(*this, min_count); }

// free functions of src/kmer/encoding.rs used by the consumers
inline uint64_t revcomp_kmer(uint64_t kmer, uint32_t k) {  // encoding.rs:235-262
    uint64_t x = ~kmer;
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
    x = (x >> 32) | (x << 32);
    return x >> (64u - 2u * k);
}
inline std::string kmer_to_seq(uint64_t kmer, uint32_t k) {  // encoding.rs:311-325
    std::string s(k, 'A');
    for (uint32_t i = 0; i < k; i++) s[i] = "ACGT"[(kmer >> (2 * (k - i - 1))) & 3];
    return s;
}
inline char kmer_last_base(uint64_t kmer) { return "ACGT"[kmer & 3]; }  // encoding.rs:301-309

}  // namespace skm
