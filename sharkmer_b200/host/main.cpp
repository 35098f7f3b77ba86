// sharkmer_b200_cli — the counting half of `sharkmer` (src/main.rs:112-131) on a B200:
//   ingest_reads (src/io.rs:366-595)  ->  consolidate_and_histogram (src/io.rs:977-1161)  ->  stats
// Flags follow src/cli.rs:165-320 for the options this path uses:
//   -k <odd 1..31> (19)  --chunks <n> (0)  --histo-max <1..1e6> (10000)  -m/--max-reads <n>
//   -s/--sample <name> (sample)  -o/--outdir <dir> (./)  --paired  --validate-every <n>
//   --capacity-hint <distinct k-mers>  --insert-mode auto|direct|partitioned  --device <n>
//   -t/--threads <n> (all cores): FASTQ framing threads (fastq_parallel.hpp); --serial: the
//   reference-shaped one-thread reader (ingest.hpp).  Both give the same batches, bit for bit.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fastq_parallel.hpp"

int main(int argc, char **argv) {
    uint32_t k = 19, chunks = 0, insert_mode = SKM_INSERT_AUTO;
    uint64_t histo_max = 10000, max_reads = 0, validate_every = 0, capacity_hint = 0;
    int device = -1;
    unsigned threads = 0;
    bool paired = false, serial = false;
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
        else if (a == "--paired") paired = true;
        else if (a == "--serial") serial = true;
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
        skm::Engine eng(k, chunks, histo_max, capacity_hint, device, insert_mode);  // validates k, histo_max
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
        skm::write_stats_file(eng, n_reads_read, n_bases_read, dir, sample, command);
        skm_totals t = eng.totals();
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
