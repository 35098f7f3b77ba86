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

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sharkmer_b200.h"
#include "../../include/skm_common.h"

namespace skm {

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string &m) : std::runtime_error(m), code(c) {}
};

// One table, on one GPU or sharded over several (src/main.rs:112-131 drives a single table; with
// n_gpus > 1 the table is split by k-mer hash over the GPUs, include/sharkmer_b200.h skm_group_*).
// Batches are dealt to the GPUs in turn; every result (histogram columns, totals, lookups, scans)
// is the one of the whole table, whatever the number of GPUs.
class Engine {
  public:
    Engine(uint32_t k, uint32_t chunks, uint64_t histo_max = 10000, uint64_t capacity_hint = 0,
           int32_t device = -1, uint32_t insert_mode = SKM_INSERT_AUTO, uint32_t n_gpus = 1,
           uint64_t arena_bytes_per_gpu = 0, const std::vector<int32_t> &devices = {}) {
        skm_params p{};
        p.struct_size = sizeof p;
        p.k = k;
        p.chunks = chunks;
        p.insert_mode = insert_mode;
        p.histo_max = histo_max;
        p.device = device;
        p.n_ranks = 1;
        k_ = k;
        chunks_ = chunks;
        histo_max_ = histo_max;
        if (n_gpus <= 1) {
            p.capacity_hint = capacity_hint;
            skm_ctx *h = nullptr;
            int32_t rc = skm_create(&p, &h);
            if (rc) {
                std::string m = h ? skm_last_error(h) : "skm_create failed";
                if (h) skm_destroy(h);
                throw Error(rc, m);
            }
            ctx_.push_back(h);
            return;
        }
        p.capacity_hint = capacity_hint ? capacity_hint / n_gpus + capacity_hint / (16 * n_gpus) + 1024 : 0;
        std::vector<int32_t> devs = devices;
        if (devs.empty())
            for (uint32_t r = 0; r < n_gpus; r++) devs.push_back((int32_t)r);
        if (devs.size() != n_gpus) throw Error(SKM_ERR_INVALID_ARG, "one device per rank");
        int32_t rc = skm_group_create(&p, n_gpus, devs.data(), arena_bytes_per_gpu ? arena_bytes_per_gpu : (8ull << 30), &group_);
        if (rc) {
            std::string m = group_ ? skm_group_last_error(group_) : "skm_group_create failed";
            if (group_) skm_group_destroy(group_);
            group_ = nullptr;
            throw Error(rc, m);
        }
        for (uint32_t r = 0; r < n_gpus; r++) ctx_.push_back(skm_group_ctx(group_, r));
    }
    ~Engine() {
        if (group_) skm_group_destroy(group_);
        else if (!ctx_.empty()) skm_destroy(ctx_[0]);
    }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;

    skm_ctx *raw() { return ctx_[0]; }
    size_t n_gpus() const { return ctx_.size(); }
    uint32_t k() const { return k_; }
    uint32_t chunks() const { return chunks_; }
    uint32_t n_chunks() const { return chunks_ == 0 ? 1 : chunks_; }
    uint64_t histo_max() const { return histo_max_; }

    void check(int32_t rc, skm_ctx *c = nullptr) const {
        if (rc) throw Error(rc, skm_last_error(c ? c : ctx_[0]));
    }
    void *pinned_alloc(size_t n) {
        void *p = nullptr;
        check(skm_pinned_alloc(ctx_[0], n, &p));
        return p;
    }
    void pinned_free(void *p) { skm_pinned_free(ctx_[0], p); }
    // a batch goes to one GPU; the batches of a run are dealt to the GPUs in turn
    void ingest_batch(uint32_t chunk, const uint8_t *seqs, uint64_t n, uint32_t flags = 0) {
        skm_ctx *c = ctx_[next_++ % ctx_.size()];
        check(skm_ingest_batch(c, chunk, seqs, n, flags), c);
    }
    void sync() {
        for (auto c : ctx_) check(skm_sync(c), c);
    }
    void finalize() {
        if (!group_) {
            check(skm_finalize(ctx_[0]));
            return;
        }
        const int32_t rc = skm_group_finalize(group_);
        if (rc) throw Error(rc, skm_group_last_error(group_));
    }
    // (after a sharded finalize every member holds the columns summed over all GPUs)
    std::vector<uint64_t> histogram(uint32_t chunk_i) {
        std::vector<uint64_t> v(histo_max_ + 2);
        check(skm_histogram(ctx_[0], chunk_i, v.data(), v.size()));
        return v;
    }
    skm_totals totals() {
        skm_totals t{};
        for (size_t r = 0; r < ctx_.size(); r++) {
            skm_totals x{};
            check(skm_totals_get(ctx_[r], &x), ctx_[r]);
            t.n_reads += x.n_reads;
            t.n_bases += x.n_bases;
            t.n_bases_read += x.n_bases_read;
            t.n_kmers += x.n_kmers;
            t.n_unique += x.n_unique;
            t.n_saturated += x.n_saturated;
            if (r == 0) t.n_singletons = x.n_singletons;  // from the (global) last column
        }
        return t;
    }
    skm_totals chunk_totals(uint32_t ch) {
        skm_totals t{};
        for (auto c : ctx_) {
            skm_totals x{};
            check(skm_chunk_totals(c, ch, &x), c);
            t.n_reads += x.n_reads;
            t.n_bases += x.n_bases;
            t.n_bases_read += x.n_bases_read;
            t.n_kmers += x.n_kmers;
        }
        return t;
    }
    skm_stage_ms stage_times() {
        skm_stage_ms t{};
        check(skm_stage_times(ctx_[0], &t));
        return t;
    }

    // ---- the table's read side, over all partitions ----
    uint64_t table_len() {
        uint64_t n = 0;
        for (auto c : ctx_) {
            uint64_t x = 0;
            check(skm_table_len(c, &x), c);
            n += x;
        }
        return n;
    }
    void insert_counts(const uint64_t *keys, const uint32_t *counts, uint64_t n) {
        if (ctx_.size() != 1) throw Error(SKM_ERR_STATE, "insert_counts on a sharded table is not supported");
        check(skm_insert_counts(ctx_[0], keys, counts, n));
    }
    // a k-mer lives in exactly one partition: the answer is the one partition's that has it
    void lookup_batch(const uint64_t *kmers, uint64_t n, uint32_t min_count, int32_t mode, uint32_t *counts, uint8_t *found) {
        if (ctx_.size() == 1) {
            check(skm_lookup_batch(ctx_[0], kmers, n, min_count, mode, counts, found));
            return;
        }
        std::vector<uint32_t> c(n);
        std::vector<uint8_t> f(n);
        for (size_t r = 0; r < ctx_.size(); r++) {
            check(skm_lookup_batch(ctx_[r], kmers, n, min_count, mode, c.data(), f.data()), ctx_[r]);
            for (uint64_t i = 0; i < n; i++) {
                if (counts && (r == 0 || c[i] > counts[i])) counts[i] = c[i];
                if (found && (r == 0 || f[i])) found[i] = (r == 0) ? f[i] : (uint8_t)(found[i] | f[i]);
            }
        }
    }
    std::pair<std::vector<uint64_t>, std::vector<uint32_t>> scan_oligos(const std::vector<uint64_t> &oligos, uint32_t oligo_length,
                                                                        uint32_t min_count) {
        std::vector<std::pair<uint64_t, uint32_t>> all;
        for (auto c : ctx_) {
            uint64_t n = 0, cap = 1 << 16;   // one table pass; a second only if the matches did not fit
            std::vector<uint64_t> keys(cap);
            std::vector<uint32_t> counts(cap);
            int32_t rc = skm_scan_oligos(c, oligos.data(), oligos.size(), oligo_length, min_count, keys.data(), counts.data(), cap, &n);
            if (rc == SKM_ERR_INVALID_ARG && n > cap) {
                keys.resize(n);
                counts.resize(n);
                rc = skm_scan_oligos(c, oligos.data(), oligos.size(), oligo_length, min_count, keys.data(), counts.data(), n, &n);
            }
            check(rc, c);
            for (uint64_t i = 0; i < n; i++) all.emplace_back(keys[i], counts[i]);
        }
        if (ctx_.size() > 1) std::sort(all.begin(), all.end());
        std::pair<std::vector<uint64_t>, std::vector<uint32_t>> out;
        for (auto &kv : all) {
            out.first.push_back(kv.first);
            out.second.push_back(kv.second);
        }
        return out;
    }
    std::pair<std::vector<uint64_t>, std::vector<uint32_t>> export_all(bool sorted) {
        std::vector<uint64_t> keys;
        std::vector<uint32_t> counts;
        for (auto c : ctx_) {
            uint64_t n = 0, got = 0;
            check(skm_table_len(c, &n), c);
            const size_t at = keys.size();
            keys.resize(at + n);
            counts.resize(at + n);
            check(skm_export(c, keys.data() + at, counts.data() + at, n, sorted && ctx_.size() == 1 ? 1 : 0, &got), c);
            keys.resize(at + got);
            counts.resize(at + got);
        }
        if (sorted && ctx_.size() > 1) {
            std::vector<size_t> idx(keys.size());
            for (size_t i = 0; i < idx.size(); i++) idx[i] = i;
            std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
            std::vector<uint64_t> k2(keys.size());
            std::vector<uint32_t> c2(keys.size());
            for (size_t i = 0; i < idx.size(); i++) {
                k2[i] = keys[idx[i]];
                c2[i] = counts[idx[i]];
            }
            keys.swap(k2);
            counts.swap(c2);
        }
        return {std::move(keys), std::move(counts)};
    }

  private:
    std::vector<skm_ctx *> ctx_;
    skm_group *group_ = nullptr;
    size_t next_ = 0;
    uint32_t k_ = 0, chunks_ = 0;
    uint64_t histo_max_ = 0;
};

class FilteredKmerCounts;

// src/kmer/counting.rs:113-312 over the device table of an Engine.
//
// Point lookups.  src/pcr asks the table one k-mer at a time (FilteredKmerCounts::get_canonical,
// four times per graph node: src/pcr/graph.rs:419-430) — as single device calls that is a kernel
// launch and a stream synchronisation per probe.  Two ways out, both exact:
//   * batch the probes (FilteredKmerCounts::get_canonical_counts / pcr.hpp's lookup waves):
//     needs the caller to ask for many k-mers at once;
//   * mirror_to_host(): ONE export of the finished table into a host open-addressing table with the
//     same hash (include/skm_common.h); get / get_canonical / get_canonical_count are then plain
//     host probes, so src/pcr runs UNCHANGED over this class.  Costs 12 B per distinct k-mer of
//     host memory and one device-to-host copy; right for sPCR runs (a few genes, millions of
//     probes), wrong for tables that do not fit the host.
class KmerCounts {
  public:
    explicit KmerCounts(Engine &e) : e_(e) {}
    uint32_t get_k() const { return e_.k(); }
    void insert(uint64_t kmer, uint32_t count) {
        e_.insert_counts(&kmer, &count, 1);
        mirrored_ = false;
    }
    void extend(KmerCounts &other) {
        if (other.get_k() != get_k()) throw Error(SKM_ERR_K_MISMATCH, "Cannot extend KmerCounts with different k");
        auto kv = other.iter();
        e_.insert_counts(kv.first.data(), kv.second.data(), kv.first.size());
        mirrored_ = false;
    }
    uint64_t len() { return mirrored_ ? m_n_ : e_.table_len(); }
    bool is_empty() { return len() == 0; }
    uint64_t get_n_kmers() { return e_.totals().n_kmers; }
    uint64_t get_n_unique_kmers() { return len(); }
    uint32_t get_count(uint64_t kmer) { return lookup1(kmer, 0, SKM_LOOKUP_EXACT).first; }
    bool contains(uint64_t kmer) { return lookup1(kmer, 0, SKM_LOOKUP_EXACT).second; }
    uint32_t get_canonical_count(uint64_t kmer) { return lookup1(kmer, 0, SKM_LOOKUP_CANONICAL).first; }
    // Some(count) / None as (count, found)
    std::pair<uint32_t, bool> get_canonical(uint64_t kmer) { return lookup1(kmer, 0, SKM_LOOKUP_EITHER); }
    // (keys, counts); sorted => ascending k-mer order
    std::pair<std::vector<uint64_t>, std::vector<uint32_t>> iter(bool sorted = false) { return e_.export_all(sorted); }
    std::vector<uint32_t> lookup(const std::vector<uint64_t> &kmers, uint32_t min_count, int32_t mode) {
        std::vector<uint32_t> c(kmers.size());
        if (mirrored_) {
            for (size_t i = 0; i < kmers.size(); i++) c[i] = host_lookup(kmers[i], min_count, mode).first;
            return c;
        }
        if (!kmers.empty()) e_.lookup_batch(kmers.data(), kmers.size(), min_count, mode, c.data(), nullptr);
        return c;
    }
    void lookup_found(const std::vector<uint64_t> &kmers, uint32_t min_count, int32_t mode, std::vector<uint32_t> &counts,
                      std::vector<uint8_t> &found) {
        counts.assign(kmers.size(), 0);
        found.assign(kmers.size(), 0);
        if (mirrored_) {
            for (size_t i = 0; i < kmers.size(); i++) {
                auto r = host_lookup(kmers[i], min_count, mode);
                counts[i] = r.first;
                found[i] = r.second;
            }
            return;
        }
        if (!kmers.empty()) e_.lookup_batch(kmers.data(), kmers.size(), min_count, mode, counts.data(), found.data());
    }
    // find_oligos_in_kmers (src/pcr/primers.rs:163-226) as one table pass on the device: table
    // k-mers with count >= min_count that start with one of the (unshifted, 2-bit) oligos, or whose
    // reverse complement does (then the reverse complement is reported); ascending k-mer order
    std::pair<std::vector<uint64_t>, std::vector<uint32_t>> scan_oligos(const std::vector<uint64_t> &oligos,
                                                                        uint32_t oligo_length, uint32_t min_count) {
        return e_.scan_oligos(oligos, oligo_length, min_count);
    }
    FilteredKmerCounts filtered_view(uint32_t min_count);
    Engine &engine() { return e_; }

    // One export of the finished table into a host open-addressing table (linear probing, the shared
    // hash, load <= 0.5): point lookups become host probes.  Call again after the table changed.
    void mirror_to_host() {
        auto kv = e_.export_all(false);
        m_n_ = kv.first.size();
        uint32_t l2 = 4;
        while ((1ull << l2) < 2 * m_n_ + 16) l2++;
        m_log2_ = l2;
        m_keys_.assign((size_t)1 << l2, SKM_EMPTY_KEY);
        m_counts_.assign((size_t)1 << l2, 0);
        const uint64_t mask = ((uint64_t)1 << l2) - 1;
        for (size_t i = 0; i < kv.first.size(); i++) {
            uint64_t s = skm_home_slot(skm_hash_kmer(kv.first[i]), l2);
            while (m_keys_[s] != SKM_EMPTY_KEY) s = (s + 1) & mask;
            m_keys_[s] = kv.first[i];
            m_counts_[s] = kv.second[i];
        }
        mirrored_ = true;
    }
    bool mirrored() const { return mirrored_; }

  private:
    friend class FilteredKmerCounts;
    std::pair<uint32_t, bool> host_find(uint64_t kmer) const {
        const uint64_t mask = ((uint64_t)1 << m_log2_) - 1;
        uint64_t s = skm_home_slot(skm_hash_kmer(kmer), m_log2_);
        for (;;) {
            const uint64_t key = m_keys_[s];
            if (key == kmer) return {m_counts_[s], true};
            if (key == SKM_EMPTY_KEY) return {0u, false};
            s = (s + 1) & mask;
        }
    }
    // the three modes of skm_lookup_batch, on the mirror
    std::pair<uint32_t, bool> host_lookup(uint64_t kmer, uint32_t min_count, int32_t mode) const {
        const uint64_t rc = skm_revcomp_kmer(kmer, e_.k());
        std::pair<uint32_t, bool> r;
        if (mode == SKM_LOOKUP_CANONICAL) r = host_find(kmer < rc ? kmer : rc);
        else if (mode == SKM_LOOKUP_EXACT) r = host_find(kmer);
        else {
            r = host_find(kmer);
            if (!r.second) r = host_find(rc);
        }
        if (r.second && r.first < min_count) r = {0u, false};
        return r;
    }
    std::pair<uint32_t, bool> lookup1(uint64_t kmer, uint32_t min_count, int32_t mode) {
        if (mirrored_) return host_lookup(kmer, min_count, mode);
        uint32_t c = 0;
        uint8_t f = 0;
        e_.lookup_batch(&kmer, 1, min_count, mode, &c, &f);
        return {c, f != 0};
    }
    Engine &e_;
    bool mirrored_ = false;
    uint64_t m_n_ = 0;
    uint32_t m_log2_ = 0;
    std::vector<uint64_t> m_keys_;
    std::vector<uint32_t> m_counts_;
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
