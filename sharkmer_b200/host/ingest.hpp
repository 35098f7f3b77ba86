// ingest.hpp — host half of the hot path: FASTQ framing, 1000-read batching,
// round-robin chunk routing, and the histogram / stats writers.
//
// Mirrors, by behaviour, caseywdunn/sharkmer v3.1.0 src/io.rs:
//   open_fastq_reader   :598-625   gzip if name ends .gz/.gzip or magic 1f 8b; ONE gzip member
//   read_fastq          :271-352   4-line records, record 0 (and every validate_every-th) validated
//   read_fastq_paired   :630-697   R1,R2 interleave; the EOF probe of R2 ingests one extra record
//   drain_batch         :355-361   every 1000th read closes a batch; batch b -> chunk b mod n
//   ingest_reads        :366-595   final partial batch drained, "No reads were ingested" error
//   consolidate_and_histogram :977-1161  .histo / .final.histo formats
// The k-mer work itself happens on the GPU behind include/sharkmer_b200.h: this
// file only fills pinned buffers with sequence lines and hands them over.
#pragma once

#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "kmer.hpp"

namespace skm {

constexpr uint64_t N_READS_PER_BATCH = 1000;  // src/io.rs:15
constexpr const char *SHARKMER_VERSION = "3.1.0";  // CARGO_PKG_VERSION in the .histo comment line (io.rs:1009-1014)

// BufRead::lines() over a plain or single-member-gzip file.
class LineReader {
  public:
    explicit LineReader(const std::string &path) : path_(path) {
        f_ = std::fopen(path.c_str(), "rb");
        if (!f_) throw Error(SKM_ERR_INVALID_ARG, "Failed to open file: " + path);
        auto ends = [&](const char *suf) {
            size_t n = std::strlen(suf);
            return path.size() >= n && path.compare(path.size() - n, n, suf) == 0;
        };
        gz_ = ends(".gz") || ends(".gzip");
        if (!gz_) {
            int a = std::fgetc(f_), b = std::fgetc(f_);
            gz_ = (a == 0x1f && b == 0x8b);
            std::rewind(f_);
        }
        out_.resize(1 << 20);
        if (gz_) {
            in_.resize(1 << 20);
            std::memset(&zs_, 0, sizeof zs_);
            if (inflateInit2(&zs_, 15 + 16) != Z_OK) throw Error(SKM_ERR_INVALID_ARG, "zlib init failed");
        }
    }
    ~LineReader() {
        if (gz_) inflateEnd(&zs_);
        if (f_) std::fclose(f_);
    }
    LineReader(const LineReader &) = delete;
    LineReader &operator=(const LineReader &) = delete;

    // true + line (without "\n" or "\r\n"); false at EOF.  Throws on a broken stream.
    bool next(std::string &line) {
        line.clear();
        bool got = false;
        for (;;) {
            if (pos_ == fill_ && !refill()) break;
            const char *s = out_.data() + pos_;
            size_t avail = fill_ - pos_;
            const char *nl = static_cast<const char *>(std::memchr(s, '\n', avail));
            size_t take = nl ? size_t(nl - s) : avail;
            line.append(s, take);
            got = true;
            pos_ += take + (nl ? 1 : 0);
            if (nl) {
                if (!line.empty() && line.back() == '\r') line.pop_back();
                return true;
            }
        }
        return got;
    }
    const std::string &name() const { return path_; }

  private:
    bool refill() {
        pos_ = fill_ = 0;
        if (eof_) return false;
        if (!gz_) {
            fill_ = std::fread(out_.data(), 1, out_.size(), f_);
            if (!fill_) eof_ = true;
            return fill_ > 0;
        }
        while (!fill_ && !zend_) {
            if (zs_.avail_in == 0) {
                zs_.next_in = reinterpret_cast<Bytef *>(in_.data());
                zs_.avail_in = (uInt)std::fread(in_.data(), 1, in_.size(), f_);
                if (zs_.avail_in == 0) {
                    eof_ = true;
                    throw Error(SKM_ERR_INVALID_ARG, "Local read stream ended unexpectedly in " + path_ +
                                                         ". The file may be truncated or corrupted.");
                }
            }
            zs_.next_out = reinterpret_cast<Bytef *>(out_.data());
            zs_.avail_out = (uInt)out_.size();
            int rc = inflate(&zs_, Z_NO_FLUSH);
            fill_ = out_.size() - zs_.avail_out;
            if (rc == Z_STREAM_END)
                zend_ = true;  // GzDecoder, not MultiGzDecoder: stop after the first member
            else if (rc != Z_OK && rc != Z_BUF_ERROR) {
                eof_ = true;
                throw Error(SKM_ERR_INVALID_ARG, "corrupt gzip stream in " + path_);
            }
        }
        if (!fill_) eof_ = true;
        return fill_ > 0;
    }
    std::string path_;
    FILE *f_ = nullptr;
    bool gz_ = false, eof_ = false, zend_ = false;
    z_stream zs_;
    std::vector<char> in_, out_;
    size_t pos_ = 0, fill_ = 0;
};

// FastqReadState (src/io.rs:200-206) + drain_batch, with the Vec<String> replaced
// by one pinned buffer per chunk.  Reads are appended as "SEQUENCE\n"; when a
// 1000-read batch closes the writer moves to the next chunk; a chunk's buffer is
// handed to the GPU when it is full.
class Batcher {
  public:
    Batcher(Engine &e, size_t buffer_bytes = 0) : e_(e), n_chunks_(e.n_chunks()) {
        if (!buffer_bytes) {
            buffer_bytes = (size_t(512) << 20) / n_chunks_;
            if (buffer_bytes > (size_t(64) << 20)) buffer_bytes = size_t(64) << 20;
            if (buffer_bytes < (size_t(1) << 20)) buffer_bytes = size_t(1) << 20;
        }
        cap_ = buffer_bytes;
        bufs_.resize(n_chunks_);
        fill_.assign(n_chunks_, 0);
        for (auto &b : bufs_) b = static_cast<uint8_t *>(e_.pinned_alloc(cap_));
    }
    ~Batcher() {
        for (auto b : bufs_) e_.pinned_free(b);
    }
    // io.rs:334-337: count, keep the sequence for the current batch
    void push(const std::string &seq) {
        n_bases_read += seq.size();
        if (fill_[chunk_index] + seq.size() + 1 > cap_) flush(chunk_index);
        if (seq.size() + 1 > cap_) {  // a single enormous read: send it on its own
            std::vector<uint8_t> tmp(seq.size() + 1);
            std::memcpy(tmp.data(), seq.data(), seq.size());
            tmp[seq.size()] = '\n';
            e_.ingest_batch(chunk_index, tmp.data(), tmp.size());
        } else {
            uint8_t *dst = bufs_[chunk_index] + fill_[chunk_index];
            std::memcpy(dst, seq.data(), seq.size());
            dst[seq.size()] = '\n';
            fill_[chunk_index] += seq.size() + 1;
        }
        n_reads_read++;
        pending_++;
    }
    // io.rs:340-343: called after every record; closes the batch on every 1000th read
    void maybe_drain() {
        if (n_reads_read % N_READS_PER_BATCH == 0) drain_batch();
    }
    // io.rs:355-361
    void drain_batch() {
        pending_ = 0;
        chunk_index = (chunk_index + 1) % n_chunks_;
    }
    // io.rs:541-543: the partial last batch stays in the current chunk; everything goes to the GPU
    void finish() {
        drain_batch();
        for (uint32_t c = 0; c < n_chunks_; c++) flush(c);
    }
    uint64_t n_reads_read = 0, n_bases_read = 0;
    uint32_t chunk_index = 0;

  private:
    void flush(uint32_t c) {
        if (!fill_[c]) return;
        e_.ingest_batch(c, bufs_[c], fill_[c]);  // returns once the buffer may be reused
        fill_[c] = 0;
    }
    Engine &e_;
    uint32_t n_chunks_;
    size_t cap_;
    std::vector<uint8_t *> bufs_;
    std::vector<size_t> fill_;
    uint64_t pending_ = 0;
};

// io.rs:161-198
inline void validate_fastq_record(const std::string &header, const std::string &separator,
                                  const std::string &quality, size_t sequence_len, uint64_t record_num) {
    auto first = [](const std::string &s) { return s.empty() ? ' ' : s[0]; };
    if (!header.empty() && header[0] == '>')
        throw Error(SKM_ERR_INVALID_ARG, "Input appears to be FASTA format, not FASTQ (record " +
                                             std::to_string(record_num + 1) +
                                             " starts with '>'). sharkmer requires FASTQ input with quality scores.");
    if (header.empty() || header[0] != '@')
        throw Error(SKM_ERR_INVALID_ARG, "FASTQ record " + std::to_string(record_num + 1) +
                                             " has invalid header (expected '@', got '" + first(header) + "'): " + header);
    if (separator.empty() || separator[0] != '+')
        throw Error(SKM_ERR_INVALID_ARG, "FASTQ record " + std::to_string(record_num + 1) +
                                             " has invalid separator line (expected '+', got '" + first(separator) +
                                             "'): " + separator);
    if (quality.size() != sequence_len)
        throw Error(SKM_ERR_INVALID_ARG, "FASTQ record " + std::to_string(record_num + 1) +
                                             " has mismatched sequence (" + std::to_string(sequence_len) +
                                             ") and quality (" + std::to_string(quality.size()) + ") lengths");
}

// io.rs:701-765.  Returns true at EOF (no record read).
inline bool read_one_fastq_record(LineReader &r, Batcher &st, uint64_t validate_every) {
    std::string header, sequence, separator, quality;
    if (!r.next(header)) return true;
    auto need = [&](std::string &dst, const char *role) {
        if (!r.next(dst))
            throw Error(SKM_ERR_INVALID_ARG, "Truncated FASTQ record at record " + std::to_string(st.n_reads_read + 1) +
                                                 " in " + r.name() + ": missing " + role + " line");
    };
    need(sequence, "sequence");
    need(separator, "separator");
    need(quality, "quality");
    const bool should_validate =
        st.n_reads_read == 0 || (validate_every > 0 && st.n_reads_read % validate_every == 0);
    if (should_validate) validate_fastq_record(header, separator, quality, sequence.size(), st.n_reads_read);
    st.push(sequence);
    return false;
}

// io.rs:271-352.  Returns true if max_reads was reached.
inline bool read_fastq(LineReader &r, Batcher &st, uint64_t max_reads, uint64_t validate_every) {
    for (;;) {
        if (read_one_fastq_record(r, st, validate_every)) return false;
        st.maybe_drain();
        if (max_reads > 0 && st.n_reads_read >= max_reads) return true;
    }
}

// io.rs:630-697
inline bool read_fastq_paired(LineReader &r1, LineReader &r2, Batcher &st, uint64_t max_reads,
                              uint64_t validate_every) {
    for (;;) {
        if (read_one_fastq_record(r1, st, validate_every)) {
            read_one_fastq_record(r2, st, validate_every);  // EOF probe; an extra R2 record IS ingested (:653-657)
            return false;
        }
        st.maybe_drain();
        if (max_reads > 0 && st.n_reads_read >= max_reads) return true;
        if (read_one_fastq_record(r2, st, validate_every)) return false;
        st.maybe_drain();
        if (max_reads > 0 && st.n_reads_read >= max_reads) return true;
    }
}

// io.rs:1049-1094: {dir}{sample}.histo and {dir}{sample}.final.histo
inline void write_histo_files(Engine &e, const std::string &directory, const std::string &sample) {
    if (e.chunks() == 0) return;  // no histogram mode
    const uint32_t n = e.n_chunks();
    std::vector<std::vector<uint64_t>> cols;
    for (uint32_t c = 0; c < n; c++) cols.push_back(e.histogram(c));
    const std::string comment =
        std::string("# sharkmer ") + SHARKMER_VERSION + " k=" + std::to_string(e.k()) + " chunks=" + std::to_string(e.chunks());
    {
        std::string path = directory + sample + ".histo";
        FILE *f = std::fopen(path.c_str(), "w");
        if (!f) throw Error(SKM_ERR_INVALID_ARG, "Failed to create histogram file");
        std::fprintf(f, "%s\n", comment.c_str());
        std::fprintf(f, "count");
        for (uint32_t c = 1; c <= n; c++) std::fprintf(f, "\tchunk_%u", c);
        std::fprintf(f, "\n");
        for (uint64_t i = 1; i < e.histo_max() + 2; i++) {
            std::fprintf(f, "%llu", (unsigned long long)i);
            for (uint32_t c = 0; c < n; c++) std::fprintf(f, "\t%llu", (unsigned long long)cols[c][i]);
            std::fprintf(f, "\n");
        }
        std::fclose(f);
    }
    {
        std::string path = directory + sample + ".final.histo";
        FILE *f = std::fopen(path.c_str(), "w");
        if (!f) throw Error(SKM_ERR_INVALID_ARG, "Failed to create final histogram file");
        std::fprintf(f, "%s\n", comment.c_str());
        std::fprintf(f, "count\tfrequency\n");
        for (uint64_t i = 1; i < e.histo_max() + 2; i++)
            std::fprintf(f, "%llu\t%llu\n", (unsigned long long)i, (unsigned long long)cols[n - 1][i]);
        std::fclose(f);
    }
}

// main.rs:182-197 / stats.rs:26-45 (scalar fields)
inline void write_stats_file(Engine &e, uint64_t n_reads_read, uint64_t n_bases_read, const std::string &directory,
                             const std::string &sample, const std::string &command) {
    skm_totals t = e.totals();
    std::string path = directory + sample + ".stats.yaml";
    FILE *f = std::fopen(path.c_str(), "w");
    if (!f) throw Error(SKM_ERR_INVALID_ARG, "Failed to create stats file");
    std::fprintf(f, "sharkmer_version: %s\n", SHARKMER_VERSION);
    std::fprintf(f, "command: %s\n", command.c_str());
    std::fprintf(f, "sample: %s\n", sample.c_str());
    std::fprintf(f, "kmer_length: %u\n", e.k());
    std::fprintf(f, "chunks: %u\n", e.chunks());
    std::fprintf(f, "n_reads_read: %llu\n", (unsigned long long)n_reads_read);
    std::fprintf(f, "n_bases_read: %llu\n", (unsigned long long)n_bases_read);
    std::fprintf(f, "n_subreads_ingested: %llu\n", (unsigned long long)t.n_reads);
    std::fprintf(f, "n_bases_ingested: %llu\n", (unsigned long long)t.n_bases);
    std::fprintf(f, "n_kmers: %llu\n", (unsigned long long)t.n_kmers);
    if (e.chunks() > 0) {
        // main.rs:193: occurrences minus the DISTINCT singleton count, saturating
        const uint64_t multi = t.n_kmers > t.n_singletons ? t.n_kmers - t.n_singletons : 0;
        std::fprintf(f, "n_multi_kmers: %llu\n", (unsigned long long)multi);
        std::fprintf(f, "n_singleton_kmers: %llu\n", (unsigned long long)t.n_singletons);
    }
    std::fprintf(f, "peak_memory_bytes: 0\n");
    std::fclose(f);
}

}  // namespace skm
