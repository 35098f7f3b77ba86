// TEST-ONLY mock of the C ABI subset that sharkmer_b200/host/ingest.hpp uses, so
// that the host-side FASTQ framing / batching / chunk routing can be checked on a
// machine without a GPU.  It records what the host hands to skm_ingest_batch.
// Never linked into the product.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/sharkmer_b200.h"
#include "../../include/skm_common.h"

static bool g_discard = false;  // --bench: time the framing only, keep nothing
struct skm_ctx {
    skm_params p;
    std::vector<std::string> chunk_data;
    std::map<uint64_t, uint32_t> table;  // --primers mode: filled through skm_insert_counts
    std::string err;
};

extern "C" {
uint32_t skm_abi_version(void) { return SKM_ABI_VERSION; }
int32_t skm_create(const skm_params *p, skm_ctx **out) {
    skm_ctx *c = new skm_ctx();
    c->p = *p;
    *out = c;
    if (p->k < 1 || p->k >= 32) { c->err = "k must be less than 32 (and at least 1), got " + std::to_string(p->k); return SKM_ERR_INVALID_ARG; }
    if (p->k % 2 == 0) { c->err = "k must be odd, got " + std::to_string(p->k); return SKM_ERR_INVALID_ARG; }
    c->chunk_data.resize(p->chunks ? p->chunks : 1);
    return SKM_OK;
}
void skm_destroy(skm_ctx *c) { delete c; }
const char *skm_last_error(const skm_ctx *c) { return c->err.c_str(); }
int32_t skm_pinned_alloc(skm_ctx *, size_t n, void **out) { *out = std::malloc(n ? n : 1); return *out ? SKM_OK : SKM_ERR_OOM; }
int32_t skm_pinned_free(skm_ctx *, void *p) { std::free(p); return SKM_OK; }
int32_t skm_ingest_batch(skm_ctx *c, uint32_t chunk, const uint8_t *seqs, uint64_t n, uint32_t) {
    if (chunk >= c->chunk_data.size()) { c->err = "chunk out of range"; return SKM_ERR_INVALID_ARG; }
    if (n && seqs[n - 1] != '\n') { c->err = "batch must end with newline"; return SKM_ERR_INVALID_ARG; }
    if (!g_discard) c->chunk_data[chunk].append(reinterpret_cast<const char *>(seqs), n);
    return SKM_OK;
}
int32_t skm_sync(skm_ctx *) { return SKM_OK; }
int32_t skm_finalize(skm_ctx *) { return SKM_OK; }
int32_t skm_histogram(skm_ctx *, uint32_t, uint64_t *, uint64_t) { return SKM_ERR_STATE; }
int32_t skm_totals_get(skm_ctx *, skm_totals *t) { std::memset(t, 0, sizeof *t); return SKM_OK; }
int32_t skm_chunk_totals(skm_ctx *, uint32_t, skm_totals *t) { std::memset(t, 0, sizeof *t); return SKM_OK; }
int32_t skm_stage_times(skm_ctx *, skm_stage_ms *t) { std::memset(t, 0, sizeof *t); return SKM_OK; }
int32_t skm_table_len(skm_ctx *c, uint64_t *n) { *n = c->table.size(); return SKM_OK; }
int32_t skm_export(skm_ctx *c, uint64_t *keys, uint32_t *counts, uint64_t cap, int32_t, uint64_t *n) {
    *n = c->table.size();
    if (!keys || cap < *n) return keys ? SKM_ERR_INVALID_ARG : SKM_OK;
    uint64_t i = 0;
    for (auto &kv : c->table) {   // (ascending: std::map)
        keys[i] = kv.first;
        counts[i++] = kv.second;
    }
    return SKM_OK;
}
// the sharded entry points are not mocked: the host's group mode needs the real library
int32_t skm_group_create(const skm_params *, uint32_t, const int32_t *, uint64_t, skm_group **out) { *out = nullptr; return SKM_ERR_STATE; }
skm_ctx *skm_group_ctx(skm_group *, uint32_t) { return nullptr; }
int32_t skm_group_finalize(skm_group *) { return SKM_ERR_STATE; }
const char *skm_group_last_error(skm_group *) { return "not mocked"; }
void skm_group_destroy(skm_group *) {}
// the lookup contract of include/sharkmer_b200.h over the mock's std::map
int32_t skm_lookup_batch(skm_ctx *c, const uint64_t *kmers, uint64_t n, uint32_t min_count, int32_t mode,
                         uint32_t *counts, uint8_t *found) {
    const uint32_t k = c->p.k;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t rc = skm_revcomp_kmer(kmers[i], k);
        auto it = c->table.end();
        if (mode == SKM_LOOKUP_CANONICAL) it = c->table.find(kmers[i] < rc ? kmers[i] : rc);
        else if (mode == SKM_LOOKUP_EXACT) it = c->table.find(kmers[i]);
        else {
            it = c->table.find(kmers[i]);
            if (it == c->table.end()) it = c->table.find(rc);
        }
        const bool ok = it != c->table.end() && it->second >= min_count;
        counts[i] = ok ? it->second : 0;
        if (found) found[i] = ok ? 1 : 0;
    }
    return SKM_OK;
}
int32_t skm_insert_counts(skm_ctx *c, const uint64_t *keys, const uint32_t *counts, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) {
        uint64_t v = (uint64_t)c->table[keys[i]] + counts[i];
        c->table[keys[i]] = v > 0xffffffffull ? 0xffffffffu : (uint32_t)v;
    }
    return SKM_OK;
}
// brute-force statement of the scan's contract (include/sharkmer_b200.h), over the mock's std::map
int32_t skm_scan_oligos(skm_ctx *c, const uint64_t *oligos, uint64_t n_oligos, uint32_t len, uint32_t min_count,
                        uint64_t *keys, uint32_t *counts, uint64_t cap, uint64_t *n_out) {
    const uint32_t k = c->p.k;
    if (len == 0 || len >= k || n_oligos == 0) { c->err = "bad oligo length"; return SKM_ERR_INVALID_ARG; }
    std::map<uint64_t, uint32_t> out;
    for (auto &kv : c->table) {
        if (kv.second < min_count) continue;
        bool fwd = false, rev = false;
        for (uint64_t i = 0; i < n_oligos && !fwd; i++) fwd = (kv.first >> (2 * (k - len))) == oligos[i];
        if (!fwd) {
            const uint64_t rc = skm_revcomp_kmer(kv.first, k);
            for (uint64_t i = 0; i < n_oligos && !rev; i++) rev = (rc >> (2 * (k - len))) == oligos[i];
            if (rev) out[rc] = kv.second;
        } else {
            out[kv.first] = kv.second;
        }
    }
    *n_out = out.size();
    if (!keys) return SKM_OK;
    if (cap < out.size()) { c->err = "scan buffers too small"; return SKM_ERR_INVALID_ARG; }
    uint64_t i = 0;
    for (auto &kv : out) {
        if (i >= cap) break;
        keys[i] = kv.first;
        counts[i++] = kv.second;
    }
    return SKM_OK;
}
}

// harness: same flags as the CLI; dumps chunk_<c>.txt + counts.txt into --dump DIR
#include "../../sharkmer_b200/host/fastq_parallel.hpp"
#include "../../sharkmer_b200/host/primers.hpp"
#include "../../sharkmer_b200/host/pcr.hpp"

int main(int argc, char **argv) {
    uint32_t k = 21, chunks = 0;
    uint64_t max_reads = 0, validate_every = 0;
    size_t buffer_bytes = 0, window_bytes = size_t(128) << 20;
    unsigned threads = 0;
    bool paired = false, serial = false, mirror = false, mirror_check = false;
    skm::PCRParams pcr;
    std::string table_path, sample = "sample", outdir = "./";
    std::vector<std::string> pcr_specs;
    uint32_t min_kmer_count = 2;
    size_t max_nodes = skm::pcr::DEFAULT_MAX_NUM_NODES;
    std::string dump = ".";
    std::vector<std::string> inputs;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "-k") k = std::atoi(argv[++i]);
        else if (a == "--chunks") chunks = std::atoi(argv[++i]);
        else if (a == "-m") max_reads = std::strtoull(argv[++i], nullptr, 10);
        else if (a == "--validate-every") validate_every = std::strtoull(argv[++i], nullptr, 10);
        else if (a == "--buffer-bytes") buffer_bytes = std::strtoull(argv[++i], nullptr, 10);
        else if (a == "--paired") paired = true;
        else if (a == "--serial") serial = true;
        else if (a == "--bench") g_discard = true;
        else if (a == "--threads" || a == "-t") threads = (unsigned)std::atoi(argv[++i]);
        else if (a == "--window-bytes") window_bytes = std::strtoull(argv[++i], nullptr, 10);
        else if (a == "--dump") dump = argv[++i];
        else if (a == "--table") table_path = argv[++i];
        else if (a == "--host-mirror") mirror = true;
        else if (a == "--mirror-check") mirror_check = true;
        else if (a == "--pcr-primers") pcr_specs.push_back(argv[++i]);
        else if (a == "--sample") sample = argv[++i];
        else if (a == "--outdir") outdir = argv[++i];
        else if (a == "--min-kmer-count") min_kmer_count = (uint32_t)std::strtoul(argv[++i], nullptr, 10);
        else if (a == "--max-nodes") max_nodes = std::strtoull(argv[++i], nullptr, 10);
        else if (a == "--lev") {   // bounded_levenshtein a b k
            std::printf("%ld\n", skm::pcr::bounded_levenshtein(argv[i + 1], argv[i + 2], std::strtoull(argv[i + 3], nullptr, 10)));
            return 0;
        }
        else if (a == "--forward") pcr.forward_seq = argv[++i];
        else if (a == "--reverse") pcr.reverse_seq = argv[++i];
        else if (a == "--mismatches") pcr.mismatches = std::strtoull(argv[++i], nullptr, 10);
        else if (a == "--trim") pcr.trim = std::strtoull(argv[++i], nullptr, 10);
        else if (a == "--min-count") pcr.min_count = (uint32_t)std::strtoul(argv[++i], nullptr, 10);
        else if (a == "--cap") pcr.max_primer_kmers = std::strtoull(argv[++i], nullptr, 10);
        else inputs.push_back(a);
    }
    try {
        skm::Engine eng(k, chunks);
        if (!table_path.empty()) {
            // --primers mode: load "kmer count" lines, run get_primer_kmers, print "F|R kmer count"
            FILE *f = std::fopen(table_path.c_str(), "r");
            unsigned long long km, ct;
            skm::KmerCounts table(eng);
            while (f && std::fscanf(f, "%llu %llu", &km, &ct) == 2) table.insert((uint64_t)km, (uint32_t)ct);
            if (f) std::fclose(f);
            if (mirror) table.mirror_to_host();
            if (mirror_check) {
                // host mirror vs the ABI's lookups: every key, its reverse complement and absent k-mers, three modes
                skm::KmerCounts direct(eng);
                table.mirror_to_host();
                auto kv = direct.iter();
                std::vector<uint64_t> probe = kv.first;
                for (uint64_t x : kv.first) probe.push_back(skm::revcomp_kmer(x, k));
                for (uint64_t i = 0; i < 500; i++) probe.push_back((i * 0x9e3779b97f4a7c15ull) >> (64 - 2 * k));
                size_t bad = 0;
                for (int mode = 0; mode < 3; mode++)
                    for (uint32_t mc : {0u, 2u, 5u}) {
                        std::vector<uint32_t> c1, c2;
                        std::vector<uint8_t> f1, f2;
                        table.lookup_found(probe, mc, mode, c1, f1);
                        direct.lookup_found(probe, mc, mode, c2, f2);
                        for (size_t i = 0; i < probe.size(); i++) bad += c1[i] != c2[i] || f1[i] != f2[i];
                    }
                std::printf("mirror %zu keys %zu probes %zu mismatches\n", kv.first.size(), probe.size(), bad);
                return bad ? 1 : 0;
            }
            if (!pcr_specs.empty()) {
                // sPCR mode: one "gene status n_products lengths... | reason" line per primer pair
                std::vector<skm::pcr::Params> runs;
                for (auto &spec : pcr_specs) {
                    runs.push_back(skm::pcr::parse_pcr_primers_string(spec));
                    for (auto &e : skm::pcr::validate_pcr_params(runs.back()))
                        throw skm::Error(SKM_ERR_INVALID_ARG, e.first + " (" + e.second + ")");
                }
                for (auto &r : skm::pcr::run_pcr(table, runs, sample, outdir, min_kmer_count, max_nodes)) {
                    std::printf("%s %s %zu", r.gene_name.c_str(), r.status.c_str(), r.product_lengths.size());
                    for (size_t l : r.product_lengths) std::printf(" %zu", l);
                    std::printf(" | %s\n", r.failure_reason.c_str());
                }
                return 0;
            }
            auto res = skm::get_primer_kmers(table, pcr);
            for (auto &e : res.first) std::printf("F %llu %u\n", (unsigned long long)e.first, e.second);
            for (auto &e : res.second) std::printf("R %llu %u\n", (unsigned long long)e.first, e.second);
            return 0;
        }
        if (!serial) {
            skm::ParallelIngest st(eng, threads, window_bytes);
            if (paired) {
                if (max_reads > 0 && max_reads % 2 != 0) max_reads += 1;
                st.read_fastq_paired(inputs[0], inputs[1], max_reads, validate_every);
            } else {
                for (auto &p : inputs)
                    if (st.read_fastq(p, max_reads, validate_every)) break;
            }
            st.finish();
            FILE *f = std::fopen((dump + "/counts.txt").c_str(), "w");
            std::fprintf(f, "%llu %llu\n", (unsigned long long)st.n_reads_read, (unsigned long long)st.n_bases_read);
            std::fclose(f);
        } else {
            skm::Batcher st(eng, buffer_bytes);
            if (paired) {
                if (max_reads > 0 && max_reads % 2 != 0) max_reads += 1;
                skm::LineReader r1(inputs[0]), r2(inputs[1]);
                skm::read_fastq_paired(r1, r2, st, max_reads, validate_every);
            } else {
                for (auto &p : inputs) {
                    skm::LineReader r(p);
                    if (skm::read_fastq(r, st, max_reads, validate_every)) break;
                }
            }
            st.finish();
            FILE *f = std::fopen((dump + "/counts.txt").c_str(), "w");
            std::fprintf(f, "%llu %llu\n", (unsigned long long)st.n_reads_read, (unsigned long long)st.n_bases_read);
            std::fclose(f);
        }
        skm_ctx *c = eng.raw();
        for (size_t i = 0; i < c->chunk_data.size(); i++) {
            FILE *f = std::fopen((dump + "/chunk_" + std::to_string(i) + ".txt").c_str(), "w");
            std::fwrite(c->chunk_data[i].data(), 1, c->chunk_data[i].size(), f);
            std::fclose(f);
        }
    } catch (const skm::Error &e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    return 0;
}
