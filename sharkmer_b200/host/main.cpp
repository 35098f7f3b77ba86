// sharkmer_b200_cli — the counting half of `sharkmer` (src/main.rs:112-131) on a B200:
//   ingest_reads (src/io.rs:366-595)  ->  consolidate_and_histogram (src/io.rs:977-1161)  ->  stats
// Flags follow src/cli.rs:165-320 for the options this path uses:
//   -k <odd 1..31> (19)  --chunks <n> (0)  --histo-max <1..1e6> (10000)  -m/--max-reads <n>
//   -s/--sample <name> (sample)  -o/--outdir <dir> (./)  --paired  --validate-every <n>
//   --capacity-hint <distinct k-mers>  --insert-mode auto|direct|partitioned  --device <n>
//   --pcr-primers "forward=..,reverse=..,name=..[,max-length=..,min-length=..,min-count=..,mismatches=..,trim=..]"
//   (repeatable; src/cli.rs:12-140)  --min-kmer-count <n> (2)  --node-budget-global <n>: in silico PCR on the
//   device table (pcr.hpp; no read threading) -> {outdir}{sample}_{gene}.fasta
//   -t/--threads <n> (all cores): FASTQ framing threads (fastq_parallel.hpp); --serial: the
//   reference-shaped one-thread reader (ingest.hpp).  Both give the same batches, bit for bit.
//   --gpus <n> (1): shard the table by k-mer hash over n GPUs (skm_group_*; --devices a,b,.. to pick them,
//   repeats allowed); --arena-mb <MiB per GPU> (8192): receive arena of the exchange.  Same output files.
//   --host-mirror: copy the finished table once into a host hash table, so that the in silico PCR's
//   point lookups are host probes (the route that leaves src/pcr unchanged)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fastq_parallel.hpp"
#include "pcr.hpp"

int main(int argc, char **argv) {
    uint32_t k = 19, chunks = 0, insert_mode = SKM_INSERT_AUTO;
    uint64_t histo_max = 10000, max_reads = 0, validate_every = 0, capacity_hint = 0;
    int device = -1;
    unsigned threads = 0, n_gpus = 1;
    uint64_t arena_mb = 8192;
    std::vector<int32_t> devices;
    bool paired = false, serial = false, host_mirror = false;
    std::vector<std::string> pcr_specs;
    uint32_t min_kmer_count = 2;
    size_t node_budget = 0;
    std::string sample = "sample", outdir = "./", command;
    std::vector<std::string> inputs;
    for (int i = 0; i < argc; i++) command += (i ? " " : "") + std::string(argv[i]);
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() -> const char * {
            if (i + 1 >= argc) {
                std::fprintf(stderr, "error: %s needs a value\n", a.c_str());
                std::exit(2);
            }
            return argv[++i];
        };
        if (a == "-k") k = (uint32_t)std::strtoul(val(), nullptr, 10);
        else if (a == "--chunks") chunks = (uint32_t)std::strtoul(val(), nullptr, 10);
        else if (a == "--histo-max") histo_max = std::strtoull(val(), nullptr, 10);
        else if (a == "-m" || a == "--max-reads") max_reads = std::strtoull(val(), nullptr, 10);
        else if (a == "-s" || a == "--sample") sample = val();
        else if (a == "-o" || a == "--outdir") outdir = val();
        else if (a == "--validate-every") validate_every = std::strtoull(val(), nullptr, 10);
        else if (a == "--capacity-hint") capacity_hint = std::strtoull(val(), nullptr, 10);
        else if (a == "--device") device = std::atoi(val());
        else if (a == "--gpus") n_gpus = (unsigned)std::strtoul(val(), nullptr, 10);
        else if (a == "--arena-mb") arena_mb = std::strtoull(val(), nullptr, 10);
        else if (a == "--devices") {
            std::string v = val();
            for (size_t at = 0; at <= v.size();) {
                size_t e = v.find(',', at);
                if (e == std::string::npos) e = v.size();
                devices.push_back(std::atoi(v.substr(at, e - at).c_str()));
                at = e + 1;
            }
        } else if (a == "--host-mirror") host_mirror = true;
        else if (a == "--paired") paired = true;
        else if (a == "--serial") serial = true;
        else if (a == "--pcr-primers") pcr_specs.push_back(val());
        else if (a == "--min-kmer-count") min_kmer_count = (uint32_t)std::strtoul(val(), nullptr, 10);
        else if (a == "--node-budget-global") node_budget = std::strtoull(val(), nullptr, 10);
        else if (a == "-t" || a == "--threads") threads = (unsigned)std::strtoul(val(), nullptr, 10);
        else if (a == "--insert-mode") {
            std::string m = val();
            insert_mode = m == "direct" ? SKM_INSERT_DIRECT : m == "partitioned" ? SKM_INSERT_PARTITIONED : SKM_INSERT_AUTO;
        } else inputs.push_back(a);
    }
    if (inputs.empty() || (paired && inputs.size() != 2)) {
        std::fprintf(stderr, "usage: sharkmer_b200_cli -k K [--chunks N] [--histo-max H] [-m N] [-s SAMPLE] [-o DIR] [--paired] reads.fastq[.gz] ...\n");
        return 2;
    }
    std::string dir = outdir;
    if (!dir.empty() && dir.back() != '/') dir += '/';
    try {
        std::vector<skm::pcr::Params> pcr_runs;   // parsed and validated before any read is touched (cli.rs:491-570)
        for (auto &spec : pcr_specs) {
            pcr_runs.push_back(skm::pcr::parse_pcr_primers_string(spec));
            auto errs = skm::pcr::validate_pcr_params(pcr_runs.back());
            if (!errs.empty()) throw skm::Error(SKM_ERR_INVALID_ARG, errs[0].first + " (" + errs[0].second + ")");
        }
        skm::Engine eng(k, chunks, histo_max, capacity_hint, device, insert_mode, n_gpus, arena_mb << 20, devices);  // validates k, histo_max
        if (paired && max_reads > 0 && max_reads % 2 != 0) max_reads += 1;  // src/io.rs:483-485
        uint64_t n_reads_read = 0, n_bases_read = 0;
        if (serial) {
            skm::Batcher st(eng);
            if (paired) {
                skm::LineReader r1(inputs[0]), r2(inputs[1]);
                skm::read_fastq_paired(r1, r2, st, max_reads, validate_every);
            } else {
                for (auto &path : inputs) {
                    skm::LineReader r(path);
                    if (skm::read_fastq(r, st, max_reads, validate_every)) break;
                }
            }
            st.finish();
            n_reads_read = st.n_reads_read;
            n_bases_read = st.n_bases_read;
        } else {
            skm::ParallelIngest st(eng, threads);
            if (paired) {
                st.read_fastq_paired(inputs[0], inputs[1], max_reads, validate_every);
            } else {
                for (auto &path : inputs)
                    if (st.read_fastq(path, max_reads, validate_every)) break;
            }
            st.finish();
            n_reads_read = st.n_reads_read;
            n_bases_read = st.n_bases_read;
        }
        eng.finalize();
        skm::write_histo_files(eng, dir, sample);
        skm_totals t = eng.totals();
        std::vector<skm::pcr::GeneResult> pcr_results;
        if (!pcr_runs.empty()) {   // main.rs:146-177
            skm::KmerCounts table(eng);
            if (host_mirror) table.mirror_to_host();
            const size_t budget = node_budget ? node_budget : skm::pcr::compute_node_budget(t.n_bases);
            pcr_results = skm::pcr::run_pcr(table, pcr_runs, sample, dir, min_kmer_count, budget);
            for (auto &r : pcr_results) {
                if (r.status == "success") {
                    std::string lens;
                    for (size_t l : r.product_lengths) lens += (lens.empty() ? "" : ", ") + std::to_string(l);
                    std::fprintf(stderr, "  + %s (%zu product%s, %s bp)\n", r.gene_name.c_str(), r.product_lengths.size(),
                                 r.product_lengths.size() == 1 ? "" : "s", lens.c_str());
                } else {
                    std::fprintf(stderr, "  - %s (no products, %s)\n", r.gene_name.c_str(), r.failure_reason.c_str());
                }
            }
        }
        skm::write_stats_file(eng, n_reads_read, n_bases_read, dir, sample, command);
        if (!pcr_results.empty()) {   // stats.rs:12-45: the pcr_results list of the stats file
            FILE *f = std::fopen((dir + sample + ".stats.yaml").c_str(), "a");
            if (f) {
                std::fprintf(f, "pcr_results:\n");
                for (auto &r : pcr_results) {
                    std::fprintf(f, "- gene_name: %s\n  status: %s\n  n_products: %zu\n", r.gene_name.c_str(), r.status.c_str(),
                                 r.product_lengths.size());
                    if (!r.product_lengths.empty()) {
                        std::fprintf(f, "  product_lengths:\n");
                        for (size_t l : r.product_lengths) std::fprintf(f, "  - %zu\n", l);
                    }
                    if (!r.failure_reason.empty()) std::fprintf(f, "  failure_reason: %s\n", r.failure_reason.c_str());
                }
                std::fclose(f);
            }
        }
        skm_stage_ms ms = eng.stage_times();
        std::fprintf(stderr, "reads %llu bases %llu kmers %llu unique %llu | device ms: h2d %.2f pack %.2f insert %.2f histogram %.2f\n",
                     (unsigned long long)n_reads_read, (unsigned long long)n_bases_read,
                     (unsigned long long)t.n_kmers, (unsigned long long)t.n_unique, ms.h2d, ms.pack, ms.insert,
                     ms.histogram);
    } catch (const skm::Error &e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    return 0;
}
