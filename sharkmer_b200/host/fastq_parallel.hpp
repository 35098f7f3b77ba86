// fastq_parallel.hpp — multi-threaded FASTQ framing into pinned staging buffers: "the step before
// the path" (SURVEY.md §8 f3).  Same observable behaviour as the serial reader in ingest.hpp, which
// follows caseywdunn/sharkmer v3.1.0 src/io.rs line by line:
//   read_fastq / read_one_fastq_record  :271-352, :701-765   a record is FOUR LINES, whatever they hold
//   validate_fastq_record               :161-198             record 0 and every validate_every-th
//   drain_batch                         :355-361             read g goes to chunk (g / 1000) mod n
//   read_fastq_paired                   :630-697             R1[j] -> read 2j, R2[j] -> read 2j+1, and the
//                                                            longer file contributes ONE extra record
//   open_fastq_reader                   :598-625             gzip by suffix or magic, one member only
// but built for throughput rather than after the reference's control flow:
//   * the text is handled in WINDOWS (default 128 MiB) that start on a record boundary; a plain file
//     is mapped, so a window is just a span of the page cache (no read() copy: framing is memory
//     bound); a gzip file is inflated by a producer thread while the previous window is being framed;
//   * every worker finds the newlines of its 1 MiB block with SSE2 compares; since a record is four
//     lines, the line table IS the record table (no per-line strings, no state machine);
//   * 1000-read batches are sized in parallel, given their place in the chunk-major staging buffer
//     by a short serial scan, and copied in parallel; each chunk's part of the staging buffer goes
//     to the GPU with one asynchronous skm_ingest_batch, and framing of the next window overlaps
//     those copies (two staging buffers).
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <exception>
#include <functional>
#include <mutex>
#include <thread>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "ingest.hpp"

namespace skm {

// Fork-join pool: parallel_for(n, fn) runs fn(0..n-1) on the workers and the calling thread.
class WorkerPool {
  public:
    explicit WorkerPool(unsigned n_threads) {
        if (n_threads == 0) n_threads = std::max(1u, std::thread::hardware_concurrency());
        n_ = n_threads;
        for (unsigned i = 1; i < n_; i++) workers_.emplace_back([this] { worker(); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            generation_++;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    WorkerPool(const WorkerPool &) = delete;
    WorkerPool &operator=(const WorkerPool &) = delete;
    unsigned size() const { return n_; }

    void parallel_for(size_t n, const std::function<void(size_t)> &fn) {
        if (n == 0) return;
        if (n == 1 || n_ == 1) {
            for (size_t i = 0; i < n; i++) fn(i);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn;
            total_ = n;
            next_.store(0);
            running_ = (unsigned)workers_.size();
            error_ = nullptr;
            generation_++;
        }
        cv_.notify_all();
        drain();
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return running_ == 0; });
        fn_ = nullptr;
        if (error_) std::rethrow_exception(error_);
    }

  private:
    void drain() {
        for (;;) {
            size_t i = next_.fetch_add(1);
            if (i >= total_) return;
            try {
                (*fn_)(i);
            } catch (...) {
                std::lock_guard<std::mutex> lk(mu_);
                if (!error_) error_ = std::current_exception();
            }
        }
    }
    void worker() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_) return;
            }
            drain();
            {
                std::lock_guard<std::mutex> lk(mu_);
                running_--;
            }
            done_cv_.notify_one();
        }
    }
    unsigned n_ = 1;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(size_t)> *fn_ = nullptr;
    size_t total_ = 0;
    std::atomic<size_t> next_{0};
    unsigned running_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
    std::exception_ptr error_;
};

// Offsets of every '\n' in [p + lo, p + hi), appended to `out` (offsets relative to p).
inline void find_newlines(const char *p, size_t lo, size_t hi, std::vector<uint32_t> &out) {
    size_t i = lo;
#if defined(__SSE2__)
    const __m128i nl = _mm_set1_epi8('\n');
    for (; i + 16 <= hi; i += 16) {
        __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(p + i));
        unsigned m = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(v, nl));
        while (m) {
            out.push_back((uint32_t)(i + (unsigned)__builtin_ctz(m)));
            m &= m - 1;
        }
    }
#endif
    for (; i < hi; i++)
        if (p[i] == '\n') out.push_back((uint32_t)i);
}

// The line table of one window of FASTQ text that starts on a record boundary.
struct WindowIndex {
    const char *base = nullptr;
    size_t size = 0;
    std::vector<uint32_t> line_end;  // offset of the '\n' that ends line i (or `size` for an unterminated last line)
    size_t n_rec = 0;                // complete four-line records
    size_t consumed = 0;             // bytes those records cover
    unsigned leftover_lines = 0;     // lines after the last complete record (a truncated record if at EOF)
    size_t next_rec = 0;             // cursor of the consumer

    size_t line_begin(size_t i) const { return i ? size_t(line_end[i - 1]) + 1 : 0; }
    // BufRead::lines(): the line without "\n" or "\r\n"
    std::pair<const char *, size_t> line(size_t i) const {
        size_t b = line_begin(i), e = line_end[i];
        if (e > b && base[e - 1] == '\r') e--;
        return {base + b, e - b};
    }
    std::pair<const char *, size_t> sequence(size_t rec) const { return line(4 * rec + 1); }
    size_t available() const { return n_rec - next_rec; }
};

inline void index_window(WorkerPool &pool, const char *p, size_t n, bool eof, WindowIndex &ix,
                         size_t block_bytes = size_t(1) << 20) {
    ix.base = p;
    ix.size = n;
    ix.next_rec = 0;
    const size_t nb = (n + block_bytes - 1) / block_bytes;
    std::vector<std::vector<uint32_t>> per_block(nb);
    pool.parallel_for(nb, [&](size_t b) {
        per_block[b].reserve(block_bytes / 64);
        find_newlines(p, b * block_bytes, std::min(n, (b + 1) * block_bytes), per_block[b]);
    });
    std::vector<size_t> first(nb + 1, 0);
    for (size_t b = 0; b < nb; b++) first[b + 1] = first[b] + per_block[b].size();
    const bool open_tail = eof && n > 0 && p[n - 1] != '\n';  // lines() still yields an unterminated last line
    ix.line_end.resize(first[nb] + (open_tail ? 1 : 0));
    pool.parallel_for(nb, [&](size_t b) {
        std::copy(per_block[b].begin(), per_block[b].end(), ix.line_end.begin() + first[b]);
    });
    if (open_tail) ix.line_end.back() = (uint32_t)n;
    const size_t n_lines = ix.line_end.size();
    ix.n_rec = n_lines / 4;
    ix.leftover_lines = (unsigned)(n_lines % 4);
    ix.consumed = ix.n_rec ? std::min(n, size_t(ix.line_end[4 * ix.n_rec - 1]) + 1) : 0;
}

// Successive windows of one FASTQ file.  After each window the caller says how many bytes it used;
// the rest (a partial record) is carried to the front of the next window.
class TextSource {
  public:
    TextSource(const std::string &path, WorkerPool &pool, size_t window_bytes)
        : path_(path), pool_(pool), window_(std::min(window_bytes, kMaxWindow)) {
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0) throw Error(SKM_ERR_INVALID_ARG, "Failed to open file: " + path);
        auto ends = [&](const char *suf) {
            size_t n = std::strlen(suf);
            return path.size() >= n && path.compare(path.size() - n, n, suf) == 0;
        };
        gz_ = ends(".gz") || ends(".gzip");  // io.rs:606-611
        if (!gz_) {                           // io.rs:612-617: gzip magic without the suffix
            unsigned char m[2] = {0, 0};
            ssize_t got = ::pread(fd_, m, 2, 0);
            gz_ = got == 2 && m[0] == 0x1f && m[1] == 0x8b;
        }
        struct stat st;
        regular_ = ::fstat(fd_, &st) == 0 && S_ISREG(st.st_mode);
        file_size_ = regular_ ? (size_t)st.st_size : 0;
        if (!gz_ && regular_ && file_size_ > 0) {
            void *m = ::mmap(nullptr, file_size_, PROT_READ, MAP_PRIVATE, fd_, 0);
            if (m != MAP_FAILED) {
                map_ = static_cast<const char *>(m);
                ::madvise(m, file_size_, MADV_SEQUENTIAL);
            }
        }
        if (gz_) producer_ = std::thread([this] { inflate_loop(); });
    }
    ~TextSource() {
        if (producer_.joinable()) {
            {
                std::lock_guard<std::mutex> lk(mu_);
                cancel_ = true;
            }
            cv_.notify_all();
            producer_.join();
        }
        if (map_) ::munmap(const_cast<char *>(map_), file_size_);
        if (fd_ >= 0) ::close(fd_);
    }
    TextSource(const TextSource &) = delete;
    TextSource &operator=(const TextSource &) = delete;

    // Load the next window, keeping buf[keep_from, size) of the current one in front of it.
    // Returns false when there is no text left at all.  `grow`: the last window held no complete
    // record, so read further instead of starting over.
    bool next(size_t keep_from) {
        if (map_) {
            // keep_from == 0 on a later call: the window held no complete record, so widen it
            span_ = (started_ && keep_from == 0 && !eof_) ? std::min(span_ * 2, kMaxWindow) : window_;
            started_ = true;
            file_pos_ += keep_from;
            const size_t left = file_size_ - file_pos_;
            data_ = map_ + file_pos_;
            size_ = std::min(span_, left);
            eof_ = size_ == left;
            return size_ > 0;
        }
        const size_t carry = size_ - keep_from;
        if (eof_ && carry == 0) {
            size_ = 0;
            return false;
        }
        if (eof_) {  // only the carried bytes remain (a truncated record)
            std::memmove(buf_.data(), buf_.data() + keep_from, carry);
            size_ = carry;
            data_ = buf_.data();
            return true;
        }
        size_t want = window_;
        if (carry + want > kMaxWindow) want = kMaxWindow - carry;
        if (want == 0) throw Error(SKM_ERR_INVALID_ARG, "FASTQ record longer than 3 GiB in " + path_);
        if (buf_.size() < carry + want) {
            std::vector<char> nb(carry + want);
            std::memcpy(nb.data(), buf_.data() + keep_from, carry);
            buf_.swap(nb);
        } else if (carry) {
            std::memmove(buf_.data(), buf_.data() + keep_from, carry);
        }
        size_t got = gz_ ? take_inflated(buf_.data() + carry, want) : read_plain(buf_.data() + carry, want);
        size_ = carry + got;
        data_ = buf_.data();
        return size_ > 0 || !eof_;
    }
    const char *data() const { return data_; }
    size_t size() const { return size_; }
    bool eof() const { return eof_; }
    const std::string &name() const { return path_; }

  private:
    static constexpr size_t kMaxWindow = size_t(3) << 30;  // line offsets are 32-bit

    size_t read_plain(char *dst, size_t want) {  // a pipe, a device or an unmappable file
        size_t n = 0;
        while (n < want) {
            ssize_t r = ::read(fd_, dst + n, want - n);
            if (r < 0) throw Error(SKM_ERR_INVALID_ARG, "Failed to read " + path_);
            if (r == 0) {
                eof_ = true;
                break;
            }
            n += (size_t)r;
        }
        return n;
    }

    // ---- gzip: a producer thread inflates ahead into two slots ---------------------------------
    struct Slot {
        std::vector<char> data;
        size_t len = 0, taken = 0;
        bool full = false, last = false;
    };
    void inflate_loop() {
        try {
            z_stream zs;
            std::memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, 15 + 16) != Z_OK) throw Error(SKM_ERR_INVALID_ARG, "zlib init failed");
            std::vector<unsigned char> in(size_t(1) << 20);
            const size_t slot_bytes = std::min(window_, size_t(64) << 20);
            bool zend = false;
            unsigned w = 0;
            while (!zend) {
                Slot &s = slots_[w & 1];
                {
                    std::unique_lock<std::mutex> lk(mu_);
                    cv_.wait(lk, [&] { return !s.full || cancel_; });
                    if (cancel_) break;
                }
                if (s.data.size() < slot_bytes) s.data.resize(slot_bytes);
                size_t fill = 0;
                while (fill < slot_bytes && !zend) {
                    if (zs.avail_in == 0) {
                        ssize_t r = ::read(fd_, in.data(), in.size());
                        if (r <= 0) {
                            inflateEnd(&zs);
                            throw Error(SKM_ERR_INVALID_ARG, "Local read stream ended unexpectedly in " + path_ +
                                                                 ". The file may be truncated or corrupted.");
                        }
                        zs.next_in = in.data();
                        zs.avail_in = (uInt)r;
                    }
                    zs.next_out = reinterpret_cast<Bytef *>(s.data.data() + fill);
                    zs.avail_out = (uInt)std::min<size_t>(slot_bytes - fill, size_t(1) << 30);
                    const size_t before = zs.avail_out;
                    int rc = inflate(&zs, Z_NO_FLUSH);
                    fill += before - zs.avail_out;
                    if (rc == Z_STREAM_END) {
                        zend = true;  // GzDecoder, not MultiGzDecoder (io.rs:619-621): one member
                    } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
                        inflateEnd(&zs);
                        throw Error(SKM_ERR_INVALID_ARG, "corrupt gzip stream in " + path_);
                    }
                }
                {
                    std::lock_guard<std::mutex> lk(mu_);
                    s.len = fill;
                    s.taken = 0;
                    s.last = zend;
                    s.full = true;
                }
                cv_.notify_all();
                w++;
            }
            inflateEnd(&zs);
        } catch (...) {
            std::lock_guard<std::mutex> lk(mu_);
            producer_error_ = std::current_exception();
            cv_.notify_all();
        }
    }
    size_t take_inflated(char *dst, size_t want) {
        size_t n = 0;
        while (n < want && !eof_) {
            Slot &s = slots_[r_ & 1];
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return s.full || producer_error_; });
                if (!s.full && producer_error_) std::rethrow_exception(producer_error_);
            }
            const size_t m = std::min(want - n, s.len - s.taken);
            std::memcpy(dst + n, s.data.data() + s.taken, m);
            n += m;
            s.taken += m;
            if (s.taken == s.len) {
                if (s.last) eof_ = true;
                {
                    std::lock_guard<std::mutex> lk(mu_);
                    s.full = false;
                }
                cv_.notify_all();
                r_++;
            }
        }
        return n;
    }

    std::string path_;
    WorkerPool &pool_;
    size_t window_;
    int fd_ = -1;
    bool gz_ = false, regular_ = false, eof_ = false;
    size_t file_size_ = 0, file_pos_ = 0;
    const char *map_ = nullptr;  // plain regular file: the whole file, mapped read-only
    size_t span_ = 0;
    bool started_ = false;
    std::vector<char> buf_;      // otherwise: carried partial record + freshly read / inflated text
    const char *data_ = nullptr;
    size_t size_ = 0;
    std::thread producer_;
    std::mutex mu_;
    std::condition_variable cv_;
    Slot slots_[2];
    unsigned r_ = 0;
    bool cancel_ = false;
    std::exception_ptr producer_error_;
};

// The parallel counterpart of Batcher + read_fastq / read_fastq_paired.
class ParallelIngest {
  public:
    ParallelIngest(Engine &e, unsigned threads = 0, size_t window_bytes = size_t(128) << 20)
        : e_(e), pool_(threads), n_chunks_(e.n_chunks()), window_(std::max<size_t>(window_bytes, 64)) {}
    ~ParallelIngest() {
        for (auto &s : stage_)
            if (s.p) e_.pinned_free(s.p);
    }
    ParallelIngest(const ParallelIngest &) = delete;
    ParallelIngest &operator=(const ParallelIngest &) = delete;

    uint64_t n_reads_read = 0, n_bases_read = 0;
    unsigned threads() const { return pool_.size(); }

    // io.rs:271-352 over one file; true when max_reads was reached (the caller stops opening files)
    bool read_fastq(const std::string &path, uint64_t max_reads, uint64_t validate_every) {
        TextSource src(path, pool_, window_);
        WindowIndex ix;
        size_t keep = 0;
        for (;;) {
            if (!src.next(keep)) return false;
            index_window(pool_, src.data(), src.size(), src.eof(), ix);
            if (ix.n_rec == 0) {
                if (src.eof()) {
                    if (ix.leftover_lines) truncated(src, ix.leftover_lines);
                    return false;
                }
                keep = 0;  // not even one record yet: read on (next() extends the carried text)
                continue;
            }
            size_t n = ix.n_rec;
            if (max_reads > 0) n = (size_t)std::min<uint64_t>(n, max_reads - n_reads_read);
            const uint64_t g0 = n_reads_read;
            validate(g0, n, validate_every, [&](uint64_t g) { return std::make_pair(&ix, size_t(g - g0)); });
            emit(g0, n, [&](uint64_t g) { return ix.sequence(size_t(g - g0)); });
            if (max_reads > 0 && n_reads_read >= max_reads) return true;
            if (src.eof() && ix.leftover_lines) truncated(src, ix.leftover_lines);
            keep = ix.consumed;
        }
    }

    // io.rs:630-697
    bool read_fastq_paired(const std::string &path1, const std::string &path2, uint64_t max_reads,
                           uint64_t validate_every) {
        TextSource s1(path1, pool_, window_), s2(path2, pool_, window_);
        WindowIndex i1, i2;
        bool more1 = true, more2 = true;  // false once the file has no complete record left
        auto refill = [&](TextSource &s, WindowIndex &ix, bool &more) {
            size_t keep = ix.base ? ix.consumed : 0;
            while (more && ix.available() == 0) {
                if (!s.next(keep)) {
                    ix.n_rec = ix.next_rec = 0;
                    ix.leftover_lines = 0;
                    more = false;
                    break;
                }
                index_window(pool_, s.data(), s.size(), s.eof(), ix);
                keep = 0;
                if (ix.n_rec == 0 && s.eof()) more = false;
            }
        };
        for (;;) {
            refill(s1, i1, more1);
            refill(s2, i2, more2);
            if (!more1) {
                // R1 is at EOF (or holds a truncated record, which errors before R2 is probed)
                if (i1.leftover_lines) truncated(s1, i1.leftover_lines);
                if (more2) {  // the EOF probe of R2 reads — and ingests — one more record (:653-657)
                    take(i2, 1, validate_every);
                } else if (i2.leftover_lines) {
                    truncated(s2, i2.leftover_lines);
                }
                return false;
            }
            if (!more2) {
                take(i1, 1, validate_every);  // R1's record is ingested before R2 reports EOF
                if (max_reads > 0 && n_reads_read >= max_reads) return true;
                if (i2.leftover_lines) truncated(s2, i2.leftover_lines);
                return false;
            }
            size_t m = std::min(i1.available(), i2.available());
            if (max_reads > 0) m = (size_t)std::min<uint64_t>(m, (max_reads - n_reads_read + 1) / 2);
            const uint64_t g0 = n_reads_read;
            const size_t a1 = i1.next_rec, a2 = i2.next_rec;
            uint64_t n = 2 * uint64_t(m);
            if (max_reads > 0 && g0 + n > max_reads) n = max_reads - g0;  // odd remainder: stop after R1's record
            auto where = [&](uint64_t g) {
                const uint64_t j = (g - g0) / 2;
                return ((g - g0) & 1) ? std::make_pair(&i2, a2 + size_t(j)) : std::make_pair(&i1, a1 + size_t(j));
            };
            validate(g0, n, validate_every, where);
            emit(g0, n, [&](uint64_t g) {
                auto w = where(g);
                return w.first->sequence(w.second);
            });
            i1.next_rec += size_t((n + 1) / 2);
            i2.next_rec += size_t(n / 2);
            if (max_reads > 0 && n_reads_read >= max_reads) return true;
        }
    }

    // everything handed over has been copied to the device
    void finish() {
        e_.sync();
        for (auto &s : stage_) s.in_flight = false;
    }

  private:
    struct Stage {
        uint8_t *p = nullptr;
        size_t cap = 0;
        bool in_flight = false;
    };

    [[noreturn]] void truncated(const TextSource &s, unsigned lines) const {
        static const char *role[4] = {"", "sequence", "separator", "quality"};
        throw Error(SKM_ERR_INVALID_ARG, "Truncated FASTQ record at record " + std::to_string(n_reads_read + 1) +
                                             " in " + s.name() + ": missing " + role[lines & 3] + " line");
    }

    void take(WindowIndex &ix, size_t n, uint64_t validate_every) {
        const uint64_t g0 = n_reads_read;
        const size_t a = ix.next_rec;
        validate(g0, n, validate_every, [&](uint64_t g) { return std::make_pair(&ix, a + size_t(g - g0)); });
        emit(g0, n, [&](uint64_t g) { return ix.sequence(a + size_t(g - g0)); });
        ix.next_rec += n;
    }

    // io.rs:321-332: read g is validated when g == 0 or g is a multiple of validate_every
    template <class Where>
    void validate(uint64_t g0, uint64_t n, uint64_t validate_every, Where where) {
        auto one = [&](uint64_t g) {
            auto w = where(g);
            const WindowIndex &ix = *w.first;
            auto h = ix.line(4 * w.second), s = ix.line(4 * w.second + 1), p = ix.line(4 * w.second + 2),
                 q = ix.line(4 * w.second + 3);
            validate_fastq_record(std::string(h.first, h.second), std::string(p.first, p.second),
                                  std::string(q.first, q.second), s.second, g);
        };
        if (n == 0) return;
        if (g0 == 0) one(0);
        if (validate_every == 0) return;
        uint64_t g = (g0 + validate_every - 1) / validate_every * validate_every;
        if (g == 0) g = validate_every;
        for (; g < g0 + n; g += validate_every) one(g);
    }

    // Copy reads [g0, g0+n) into the staging buffer, chunk-major, and hand every chunk's part over.
    template <class Seq>
    void emit(uint64_t g0, uint64_t n, Seq seq) {
        if (n == 0) return;
        const uint64_t g1 = g0 + n;
        const uint64_t b0 = g0 / N_READS_PER_BATCH, nb = (g1 - 1) / N_READS_PER_BATCH - b0 + 1;
        std::vector<uint64_t> bytes(nb), bases(nb), at(nb);
        auto range = [&](uint64_t b, uint64_t &lo, uint64_t &hi) {
            lo = std::max(g0, (b0 + b) * N_READS_PER_BATCH);
            hi = std::min(g1, (b0 + b + 1) * N_READS_PER_BATCH);
        };
        pool_.parallel_for(nb, [&](size_t b) {
            uint64_t lo, hi, s = 0;
            range(b, lo, hi);
            for (uint64_t g = lo; g < hi; g++) s += seq(g).second;
            bases[b] = s;
            bytes[b] = s + (hi - lo);  // every read is followed by '\n'
        });
        std::vector<uint64_t> chunk_bytes(n_chunks_, 0), chunk_base(n_chunks_ + 1, 0);
        for (uint64_t b = 0; b < nb; b++) chunk_bytes[(b0 + b) % n_chunks_] += bytes[b];
        for (uint32_t c = 0; c < n_chunks_; c++) chunk_base[c + 1] = chunk_base[c] + chunk_bytes[c];
        std::vector<uint64_t> fill(chunk_base.begin(), chunk_base.end() - 1);
        for (uint64_t b = 0; b < nb; b++) {
            const uint32_t c = (uint32_t)((b0 + b) % n_chunks_);
            at[b] = fill[c];
            fill[c] += bytes[b];
            n_bases_read += bases[b];
        }
        Stage &st = stage_[stage_next_++ & 1];
        if (st.in_flight) {  // its previous contents may still be on their way to the device
            e_.sync();
            for (auto &s : stage_) s.in_flight = false;
        }
        const size_t total = (size_t)chunk_base[n_chunks_];
        if (st.cap < total) {
            if (st.p) e_.pinned_free(st.p);
            st.p = nullptr;
            st.cap = 0;
            st.p = static_cast<uint8_t *>(e_.pinned_alloc(total + total / 8 + 4096));
            st.cap = total + total / 8 + 4096;
        }
        uint8_t *out = st.p;
        pool_.parallel_for(nb, [&](size_t b) {
            uint64_t lo, hi;
            range(b, lo, hi);
            uint8_t *dst = out + at[b];
            for (uint64_t g = lo; g < hi; g++) {
                auto s = seq(g);
                std::memcpy(dst, s.first, s.second);
                dst[s.second] = '\n';
                dst += s.second + 1;
            }
        });
        for (uint32_t c = 0; c < n_chunks_; c++)
            if (chunk_bytes[c]) e_.ingest_batch(c, out + chunk_base[c], chunk_bytes[c], SKM_INGEST_ASYNC);
        st.in_flight = true;
        n_reads_read += n;
    }

    Engine &e_;
    WorkerPool pool_;
    uint32_t n_chunks_;
    size_t window_;
    Stage stage_[2];
    unsigned stage_next_ = 0;
};

}  // namespace skm
