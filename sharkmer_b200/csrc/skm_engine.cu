// skm_engine.cu — host side of the C ABI (include/sharkmer_b200.h).
//
// One ctx = one GPU = one table partition.  Reads are staged in HBM as 2-bit
// codes + break mask per chunk (pack kernel runs at ingest time, in stream
// order); skm_finalize then walks the chunks in index order — extract+insert,
// histogram snapshot — which is what consolidate_and_histogram's merge loop
// computes (src/io.rs:1016-1047): column i of the histogram is the histogram of
// the table that holds every read of chunks 0..i.
//
// There is no CPU fallback anywhere in this file.
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/sharkmer_b200.h"
#include "skm_kernels.cuh"

using namespace skm;

namespace {

constexpr double kMaxLoad = 0.70;       // grow once the exact load passes this
constexpr double kHardLoad = 0.90;      // the host's upper bound on the load never passes this
constexpr double kTargetLoad = 0.60;    // load at capacity_hint
constexpr uint64_t kMinTile = 1ull << 22;  // k-mers; smallest tile worth a launch near the limit
constexpr uint32_t kMinLog2Cap = 16;

enum Stage { ST_H2D, ST_PACK, ST_COUNT, ST_PART, ST_INSERT, ST_HISTO, ST_GROW, ST_FINALIZE, ST_SCAN, ST_SORT, ST_N };

// Multi-GPU: the k-mers of one batch that belong to ONE owner rank, bucketed by that owner's kFineRegions
// table regions and tile-sorted — the same geometry a single GPU builds for itself.  The owner's own list
// stays here; the others are shipped whole into the owners' receive arenas.
struct OwnerList {
    unsigned long long *list = nullptr, *meta = nullptr;
    tile_off_t *tile_off = nullptr;
    uint32_t max_tiles = 0;
    uint64_t cap = 0;           // capped layout: cells per bucket (0 = exact layout)
    size_t cells = 0;
    uint64_t *h_off = nullptr;  // pinned: capped = kFineRegions totals + overflow total; exact = kFineRegions + 1 offsets
    bool shipped = false;
    size_t rec = 0;             // its record in skm_ctx::mg_sent
};

struct Segment {
    std::vector<OwnerList> owners;   // n_ranks > 1 only
    uint64_t *codes = nullptr;
    uint32_t *breaks = nullptr;
    uint64_t n_units = 0;
    uint64_t n_bytes = 0;
    // eager partition (single GPU): the segment's k-mers, already ordered by table region
    unsigned long long *list = nullptr;
    uint64_t *h_offsets = nullptr;  // pinned, n_buckets + 1 entries (valid once `ready` has fired)
    cudaEvent_t ready = nullptr;    // fires when the segment's pack (+ bucketing) has finished
    uint32_t n_buckets = 0;
    unsigned long long *d_counts = nullptr;  // per-bucket counts on the device (multi-GPU routing)
    // capped layout (CapLayout; single GPU): h_offsets then holds the n_buckets bucket totals + the
    // overflow total instead of offsets, and codes/breaks stay alive until the totals were checked
    uint64_t cap = 0, ovf_cap = 0;
    size_t list_cells = 0;  // cells allocated for `list`
    // tile-sorted form (tiled insert): where the buckets' tiles are + the sub-bucket offsets of every tile
    unsigned long long *meta = nullptr;   // ListMeta arrays (list_meta_words(n_buckets) words)
    tile_off_t *tile_off = nullptr;       // max_tiles * (2^g2 + 1)
    uint32_t max_tiles = 0;
    bool tiled = false;
    cudaEvent_t copied = nullptr;   // host-fed batch: fires when its host-to-device copy has finished
    bool shipped = false;       // multi-GPU: the other owners' slices are on their way (or there)
    size_t rec_first = 0;       // first of this list's n_ranks - 1 records in skm_ctx::mg_sent
};

struct ChunkState {
    std::vector<Segment> segs;
    uint64_t n_bytes = 0;       // staged bytes incl. separators
    uint64_t n_windows_host = 0;
    bool counted = false;
};

// One list slice shipped to another rank: where it lies inside the sender's sub-arena at the
// destination (byte offsets), and what it holds.  Plain data: exchanged by all-gather at finalize.
struct MgRecord {
    uint32_t src, dst, chunk, dead;   // dead: superseded (a capped list that overflowed and was rebuilt)
    uint32_t regions, n_tiles;
    uint64_t off_cells, n_cells;
    uint64_t off_tile_off, off_cell_begin, off_tile_begin;
    uint64_t n_kmers;
};

struct TimedSpan {
    cudaEvent_t a, b;
    int stage;
};

}  // namespace

struct skm_ctx {
    skm_params p{};
    int device = 0;
    uint32_t n_chunks = 1;
    uint32_t n_ranks = 1;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host->device copies of raw batches, nothing else: the copy
                                          // engine never queues behind a kernel that waits for an SM
    cudaStream_t dma_stream = nullptr;   // peer copies of routed k-mers (copy engines)
    std::vector<cudaStream_t> dma_peer;  // n_ranks > 1: one copy stream per destination rank.  (On ONE stream the seven
                                         // peers' copies of a batch ran one after the other on one copy engine: at N=8
                                         // 9.5 GB per rank and step took ~38 ms and the insert waited 11 ms for the tail.)
    cudaEvent_t ev_dma = nullptr;
    cudaStream_t part_stream = nullptr;  // bucketing of incoming batches (overlaps inserts on `stream`)
    cudaStream_t pack_stream = nullptr;  // pack kernels of incoming batches.  A stream of their own (high priority): on the
                                         // bucketing stream, pack(b+1) sat between pass A of b and of b+1, ran 12x slower
                                         // beside the tile sort of b, and so kept pass A(b+1) from overlapping that sort
    cudaStream_t sort_stream = nullptr;  // tile sort of a bucketed batch (overlaps the bucketing of the next one)
    cudaEvent_t ev_sort = nullptr;
    bool sort_overlap = true;            // SKM_SORT_OVERLAP=0: tile sort on the bucketing stream
    cudaStream_t work = nullptr;         // stream the bucketing helpers currently launch on
    cudaEvent_t ev_main = nullptr;
    uint32_t insert_ctas_per_sm = 5;      // persistent insert grid (SKM_INSERT_CTAS): 5 of the 6 CTAs that fit,
                                          // the rest of the SM is for the bucketing of the next chunk
    int sm_count = 148;

    Slot *table = nullptr;
    uint32_t log2cap = 0;
    uint64_t capacity = 0;
    uint64_t distinct_ub = 0;   // upper bound on occupied slots (host-side bookkeeping)

    ChunkCounters *d_cc = nullptr;  // n_chunks
    GlobalCounters *d_gc = nullptr;
    unsigned long long *d_hist = nullptr;   // running histogram, histo_max + 2 (chunks > 0)
    bool track_histo = false;
    int region_log2 = 19;       // partitioned mode: slots per table region, 2^19 = 8 MiB (SKM_REGION_LOG2)
    int pipe_depth = 1;         // probes in flight per thread (SKM_PIPE_DEPTH: 1, 2, 4, 8); measured best: 1
    unsigned long long *d_bins = nullptr;   // histo_max + 2 (scan histogram)
    HistoTotals *d_tot = nullptr;
    uint64_t *h_pinned = nullptr;           // small pinned scratch for read-backs
    size_t h_pinned_words = 0;

    std::vector<ChunkState> chunks;
    std::vector<std::vector<uint64_t>> histos;  // per chunk snapshot
    uint64_t *h_cols = nullptr;                 // pinned landing area for asynchronous column snapshots
    std::vector<bool> col_pending;
    std::vector<bool> have_histo;
    std::vector<ChunkCounters> h_cc;
    HistoTotals last_tot{};
    bool have_tot = false;
    uint64_t n_bases_read = 0;
    std::vector<uint64_t> chunk_bases_read;
    uint64_t pos_base = 0;  // running byte position for error reports
    bool finalized = false;
    bool sticky_error = false;

    // pinned arena for the per-segment bucket offsets
    std::vector<uint64_t *> off_blocks;
    size_t off_used = 0;  // entries used in the last block
    bool eager = true;    // SKM_EAGER=0 disables partitioning at ingest time
    uint32_t max_buckets = kMaxBuckets;  // SKM_MAX_BUCKETS (power of two <= kMaxBuckets; same on every rank)
    bool capped = true;   // SKM_CAPPED=0: single-GPU eager bucketing uses the exact two-pass layout
    uint32_t n_capped_fallbacks = 0;
    size_t mem_budget = 0, list_bytes = 0;
    cudaEvent_t ev_alloc = nullptr, ev_copy = nullptr;
    // ring of raw (ASCII) device buffers for host batches: the copy stream may run kRawRing-1 batches
    // ahead of the pack kernels (which may be waiting for the insert kernel to release the SMs)
    static constexpr uint32_t kRawRing = 4;
    uint8_t *raw_buf[kRawRing] = {};
    size_t raw_cap[kRawRing] = {};
    cudaEvent_t raw_copied[kRawRing] = {}, raw_packed[kRawRing] = {};
    bool raw_in_use[kRawRing] = {};
    uint32_t raw_next = 0;

    // multi-GPU exchange: this rank's receive arena (cut into one sub-arena per source rank) and
    // the peers' arenas as mapped here.  A sender bump-allocates inside ITS sub-arena of every peer.
    static constexpr uint32_t kMaxPeers = 16;
    uint8_t *mg_arena = nullptr;
    size_t mg_arena_bytes = 0, mg_sub_bytes = 0;
    uint8_t *mg_peer[kMaxPeers] = {};
    bool mg_peer_ipc[kMaxPeers] = {};
    size_t mg_cursor[kMaxPeers] = {};          // bytes this rank has used of its sub-arena at each peer
    std::vector<MgRecord> mg_sent;             // what this rank shipped, in order
    uint64_t mg_bytes_sent = 0;

    // asynchronous snapshots of the device's occupied-slot counter
    static constexpr uint32_t kSnapRing = 8;
    uint64_t *h_snap = nullptr;  // pinned, kSnapRing entries
    cudaEvent_t snap_event[kSnapRing] = {};
    uint64_t snap_launched[kSnapRing] = {};
    bool snap_pending[kSnapRing] = {};
    uint32_t snap_next = 0;
    uint64_t launched_total = 0;

    // routing / partition scratch
    unsigned long long *d_bucket_counts = nullptr, *d_bucket_offsets = nullptr, *d_bucket_cursors = nullptr;
    unsigned long long *d_list = nullptr;
    uint64_t list_cap = 0;

    float stage_ms[ST_N] = {0};
    uint32_t stage_launches[ST_N] = {0};
    uint64_t insert_kmers = 0, insert_bases = 0;
    bool own_stream = true;
    std::vector<TimedSpan> spans;
    std::vector<cudaEvent_t> event_pool;
    uint32_t launches = 0;
    uint32_t n_grows = 0;

    std::vector<cudaStream_t> read_streams;  // idle private streams of the read-side calls

    // Cache of the big device buffers (k-mer lists, tile offsets).  cudaMallocAsync blocked the host for
    // milliseconds per list once peers were mapped and lists were freed on another stream than they were
    // allocated on (new physical memory every batch: 50-400 ms of host time per step at N=2, profiles/
    // experiments_r02.md); a freed buffer is kept here with an event that marks its last use instead.
    struct CachedBuf {
        void *p;
        size_t bytes;
        cudaEvent_t done;
    };
    std::vector<CachedBuf> buf_cache;
    std::vector<std::pair<void *, size_t>> buf_live;

    // tiled insert (tile_insert_kernel)
    uint32_t g2 = 7;                         // sub-bucket bits of the lists built by this ctx
    uint32_t tile_log2 = kTileLog2;          // cells per tile of those lists: 2^13 (one CTA sorts a tile) .. 2^16 (a cluster
                                             // of 8 does), SKM_TILE_LOG2
    int sort_clusters[3] = {0, 0, 0};        // resident clusters of tile_sort_cluster_kernel<2 / 4 / 8> (persistent grid)
    static constexpr uint32_t kSortCounters = 64;
    unsigned int *d_sort_counters = nullptr; // tile counters of the cluster sorts in flight (a ring)
    uint32_t sort_counter_next = 0;
    bool mg_slices = true;                   // multi-GPU: an owner gets its SLICE of the sender's one list, sorted by a
                                             // cluster straight down to the owner's partitions; false: every owner's
                                             // slice is re-bucketed into a list of its own first (skm_create decides;
                                             // SKM_MG_SLICES=1/0 forces)
    bool table_fresh = true;                 // logically empty: no key was ever inserted since create / reset
    bool table_zombie = false;               // logically empty but NOT physically cleared (skm_reset defers the clear:
                                             // the tiled insert starts every partition empty and rewrites the whole table)
    unsigned long long *d_delta = nullptr;   // per-chunk histogram moves of one launch
    size_t delta_words = 0;
    unsigned long long *d_recount = nullptr; // histogram of the written-back partitions (histo_max + 2)
    uint32_t *d_fail = nullptr;              // failed partitions of the last launch
    uint32_t *d_part_ids = nullptr;          // partitions to retry
    static constexpr uint32_t kFailCap = 1u << 20;
    void *h_segs = nullptr, *d_segs = nullptr;   // SegDesc[] + chunk_first_seg[] of one launch (pinned / device)
    size_t segs_bytes = 0;
    uint32_t n_tiled_launches = 0, n_tiled_retries = 0;
    double host_ms[6] = {0, 0, 0, 0, 0, 0};  // SKM_DEBUG: host time in build / refine-alloc / refine-launch / ship / pack / other
    bool recount_valid = false;              // d_recount / last_tot describe the whole table as it is now

    std::string err;
    std::mutex mu;
};

namespace {

int32_t fail(skm_ctx *c, int32_t code, const char *fmt, ...) {
    if (c) {
        char buf[1024];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        c->err = buf;
    }
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(c, e_ == cudaErrorMemoryAllocation ? SKM_ERR_OOM : SKM_ERR_CUDA,           \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// SKM_SYNC_DEBUG=1: synchronise the device after the marked launches and report where a fault surfaced
#define DBG_SYNC(where)                                                                                  \
    do {                                                                                                 \
        static const bool dbg_ = getenv("SKM_SYNC_DEBUG") != nullptr;                                    \
        if (dbg_) {                                                                                      \
            cudaError_t e_ = cudaDeviceSynchronize();                                                    \
            if (e_ != cudaSuccess) return fail(c, SKM_ERR_CUDA, "fault after %s: %s", where, cudaGetErrorString(e_)); \
        }                                                                                                \
    } while (0)

struct HostTimer {
    double *acc;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit HostTimer(double *a) : acc(a) {}
    ~HostTimer() { *acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

cudaEvent_t get_event(skm_ctx *c) {
    if (!c->event_pool.empty()) {
        cudaEvent_t e = c->event_pool.back();
        c->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

size_t buf_round(size_t bytes) {
    size_t gran = 1 << 20;
    while (gran * 16 < bytes) gran <<= 1;   // <= 12.5 % above the request
    return (bytes + gran - 1) / gran * gran;
}

// A buffer of >= `bytes` bytes, usable on stream `st` (stream-ordered: a cached buffer's last use is waited for).
cudaError_t buf_alloc(skm_ctx *c, void **out, size_t bytes, cudaStream_t st) {
    bytes = buf_round(std::max<size_t>(bytes, 16));
    size_t best = SIZE_MAX;
    for (size_t i = 0; i < c->buf_cache.size(); i++)
        if (c->buf_cache[i].bytes >= bytes && c->buf_cache[i].bytes <= bytes + bytes / 4 &&
            (best == SIZE_MAX || c->buf_cache[i].bytes < c->buf_cache[best].bytes))
            best = i;
    if (best != SIZE_MAX) {
        skm_ctx::CachedBuf cb = c->buf_cache[best];
        c->buf_cache.erase(c->buf_cache.begin() + best);
        cudaStreamWaitEvent(st, cb.done, 0);
        c->event_pool.push_back(cb.done);
        c->buf_live.emplace_back(cb.p, cb.bytes);
        *out = cb.p;
        return cudaSuccess;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation) {   // give the cached buffers back and retry
        cudaGetLastError();
        cudaDeviceSynchronize();
        for (auto &cb : c->buf_cache) {
            cudaFree(cb.p);
            c->event_pool.push_back(cb.done);
        }
        c->buf_cache.clear();
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) return e;
    c->buf_live.emplace_back(p, bytes);
    *out = p;
    return cudaSuccess;
}

// Stream-ordered free: everything that uses `p` has been queued on `st` (or on streams `st` waits for).
void buf_free(skm_ctx *c, void *p, cudaStream_t st) {
    if (!p) return;
    size_t bytes = 0;
    for (size_t i = 0; i < c->buf_live.size(); i++)
        if (c->buf_live[i].first == p) {
            bytes = c->buf_live[i].second;
            c->buf_live[i] = c->buf_live.back();
            c->buf_live.pop_back();
            break;
        }
    if (!bytes) {   // not from the cache (should not happen): hand it to the stream-ordered pool
        cudaFreeAsync(p, st);
        return;
    }
    cudaEvent_t e = get_event(c);
    cudaEventRecord(e, st);
    c->buf_cache.push_back(skm_ctx::CachedBuf{p, bytes, e});
}

struct Span {
    skm_ctx *c;
    cudaStream_t s;
    TimedSpan t;
    Span(skm_ctx *c_, int stage, cudaStream_t s_) : c(c_), s(s_) {
        t.stage = stage;
        t.a = get_event(c);
        t.b = get_event(c);
        cudaEventRecord(t.a, s);
    }
    ~Span() {
        cudaEventRecord(t.b, s);
        c->spans.push_back(t);
    }
};

// Call only after the streams are synchronised.  SKM_TRACE=<file>: also append one line per span
// ("batch stage start_ms end_ms", times relative to the first span recorded since the last
// collection) — a poor man's timeline of the copy / pack / bucketing / insert streams.
void collect_spans(skm_ctx *c) {
    static const char *trace_path = getenv("SKM_TRACE");
    if (trace_path && !c->spans.empty()) {
        static const char *names[] = {"h2d", "pack", "count", "partition", "insert", "histogram", "grow", "other", "scan", "sort"};
        static int batch = 0;
        if (FILE *f = std::fopen(trace_path, "a")) {
            for (auto &t : c->spans) {
                float a = 0, b = 0;
                if (cudaEventElapsedTime(&a, c->spans[0].a, t.a) == cudaSuccess &&
                    cudaEventElapsedTime(&b, c->spans[0].a, t.b) == cudaSuccess)
                    std::fprintf(f, "%d %s %.3f %.3f\n", batch, names[t.stage < 10 ? t.stage : 7], a, b);
            }
            std::fclose(f);
        }
        batch++;
    }
    for (auto &t : c->spans) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) c->stage_ms[t.stage] += ms;
        c->event_pool.push_back(t.a);
        c->event_pool.push_back(t.b);
    }
    c->spans.clear();
}

// zero `bytes` (a multiple of 8) of device memory with a kernel, never a copy engine
void zero_async(skm_ctx *c, void *p, size_t bytes, cudaStream_t st) {
    const uint64_t words = bytes / 8;
    const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((words + 255) / 256, 1024));
    zero_kernel<<<grid, 256, 0, st>>>((unsigned long long *)p, words);
    c->launches++;
}

TableRef tref(const skm_ctx *c) { return TableRef{c->table, c->log2cap, c->n_ranks}; }

uint32_t ceil_log2(uint64_t v) {
    uint32_t l = 0;
    while ((1ull << l) < v) l++;
    return l;
}

inline uint32_t grid_for(uint64_t n, uint32_t block) { return (uint32_t)((n + block - 1) / block); }

int32_t alloc_table(skm_ctx *c, uint32_t log2cap, Slot **out, bool clear = true) {
    Slot *t = nullptr;
    const uint64_t cap = 1ull << log2cap;
    CU(cudaMallocAsync((void **)&t, cap * sizeof(Slot), c->stream));
    if (clear) {
        table_clear_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(t, cap);
        c->launches++;
        CU(cudaGetLastError());
    }
    *out = t;
    return SKM_OK;
}

// skm_reset (and a re-sized empty table) leave the table "zombie": logically empty, physically
// stale.  The tiled insert never reads it (every partition starts empty in shared memory and the
// whole table is rewritten); anything else that touches the table clears it first.
int32_t ensure_physical(skm_ctx *c) {
    if (!c->table_zombie) return SKM_OK;
    table_clear_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->table, c->capacity);
    c->launches++;
    CU(cudaGetLastError());
    c->table_zombie = false;
    return SKM_OK;
}

int32_t read_distinct(skm_ctx *c, uint64_t *out) {
    CU(cudaMemcpyAsync(c->h_pinned, &c->d_gc->n_distinct, sizeof(uint64_t), cudaMemcpyDeviceToHost,
                       c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *out = c->h_pinned[0];
    return SKM_OK;
}

int32_t grow_table(skm_ctx *c, uint32_t new_log2cap) {
    Span sp(c, ST_GROW, c->stream);
    Slot *nt = nullptr;
    if (c->table_fresh) {  // nothing to move; the new table is cleared lazily (ensure_physical)
        CU(cudaFreeAsync(c->table, c->stream));
        c->table = nullptr;
        int32_t rc0 = alloc_table(c, new_log2cap, &nt, false);
        if (rc0) return rc0;
        c->table_zombie = true;
        c->table = nt;
        c->log2cap = new_log2cap;
        c->capacity = 1ull << new_log2cap;
        c->n_grows++;
        return SKM_OK;
    }
    int32_t rc = alloc_table(c, new_log2cap, &nt);
    if (rc) return rc;
    rehash_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->table, c->capacity,
                                                          TableRef{nt, new_log2cap, c->n_ranks}, c->d_gc);
    c->launches++;
    c->stage_launches[ST_GROW]++;
    CU(cudaGetLastError());
    CU(cudaFreeAsync(c->table, c->stream));
    c->table = nt;
    c->log2cap = new_log2cap;
    c->capacity = 1ull << new_log2cap;
    c->n_grows++;
    return SKM_OK;
}

// Occupancy bookkeeping without stalling the host.  The exact number of occupied slots lives on
// the device; the host keeps an UPPER BOUND: the last count it has seen plus every k-mer launched
// since.  After each insert launch a copy of the device counter is queued into a small pinned ring
// (note_inserted); reserve_headroom picks up whichever copies have completed (event query, no
// wait) before it falls back to a blocking read.
void note_inserted(skm_ctx *c, uint64_t n) {
    c->distinct_ub += n;
    const uint32_t slot = c->snap_next++ % skm_ctx::kSnapRing;
    if (!c->snap_event[slot]) cudaEventCreateWithFlags(&c->snap_event[slot], cudaEventDisableTiming);
    cudaMemcpyAsync(&c->h_snap[slot], &c->d_gc->n_distinct, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream);
    cudaEventRecord(c->snap_event[slot], c->stream);
    c->snap_launched[slot] = c->launched_total += n;
    c->snap_pending[slot] = true;
}

void absorb_snapshots(skm_ctx *c) {
    for (uint32_t i = 0; i < skm_ctx::kSnapRing; i++) {
        if (!c->snap_pending[i] || cudaEventQuery(c->snap_event[i]) != cudaSuccess) continue;
        c->snap_pending[i] = false;
        // occupied <= (count when this snapshot was taken) + (k-mers launched after it)
        const uint64_t ub = c->h_snap[i] + (c->launched_total - c->snap_launched[i]);
        if (ub < c->distinct_ub) c->distinct_ub = ub;
    }
}

// Make room for `want` more k-mers; returns how many may be inserted now.  The table is sized for
// load <= kTargetLoad and grown when the EXACT count passes kMaxLoad; the upper bound only has to
// stay below kHardLoad (linear probing needs free slots to terminate).
int32_t reserve_headroom(skm_ctx *c, uint64_t want, uint64_t *granted) {
    auto headroom = [&](double load) -> uint64_t {
        const uint64_t limit = (uint64_t)(load * (double)c->capacity);
        return limit > c->distinct_ub ? limit - c->distinct_ub : 0;
    };
    const uint64_t need = std::min(want, kMinTile);
    if (headroom(kHardLoad) < want) absorb_snapshots(c);
    if (headroom(kHardLoad) < want) {
        // the bound is too loose: get the real count (blocks until the queued inserts have finished)
        uint64_t d = 0;
        int32_t rc = read_distinct(c, &d);
        if (rc) return rc;
        c->distinct_ub = d;
        for (auto &p : c->snap_pending) p = false;
        if (headroom(kMaxLoad) < need) {
            uint32_t nl = c->log2cap + 1;
            while ((uint64_t)(kTargetLoad * (double)(1ull << nl)) < d + need * 4) nl++;
            rc = grow_table(c, nl);
            if (rc) return rc;
        }
        *granted = std::min(want, std::max(headroom(kMaxLoad), std::min(need, headroom(kHardLoad))));
        return SKM_OK;
    }
    *granted = want;
    return SKM_OK;
}

// ---- direct mode: fused extract + insert over one segment -------------------
int32_t insert_segment_direct(skm_ctx *c, const Segment &sg, uint32_t chunk) {
    c->recount_valid = false;
    uint64_t u = 0;
    while (u < sg.n_units) {
        uint64_t granted = 0;
        int32_t rc = reserve_headroom(c, (sg.n_units - u) * 32, &granted);
        if (rc) return rc;
        rc = ensure_physical(c);  // (after the headroom check: an empty table may just have been re-sized)
        if (rc) return rc;
        uint64_t tile_units = std::max<uint64_t>(granted / 32, 1);
        tile_units = std::min(tile_units, sg.n_units - u);
        {
            Span sp(c, ST_INSERT, c->stream);
#define SKM_LAUNCH_EI(D, H)                                                                              \
    extract_insert_kernel<D, H><<<grid_for(tile_units, 256), 256, 0, c->stream>>>(                      \
        sg.codes, sg.breaks, u, u + tile_units, c->p.k, tref(c), &c->d_cc[chunk], c->d_gc,              \
        c->d_hist, c->p.histo_max)
            const bool h = c->track_histo;
            static const bool coop = getenv("SKM_WARP_COOP") != nullptr;   // A/B: warp-cooperative probing
            if (coop) {
                if (h) extract_insert_coop_kernel<true><<<grid_for(tile_units, 256), 256, 0, c->stream>>>(
                    sg.codes, sg.breaks, u, u + tile_units, c->p.k, tref(c), &c->d_cc[chunk], c->d_gc, c->d_hist, c->p.histo_max);
                else extract_insert_coop_kernel<false><<<grid_for(tile_units, 256), 256, 0, c->stream>>>(
                    sg.codes, sg.breaks, u, u + tile_units, c->p.k, tref(c), &c->d_cc[chunk], c->d_gc, c->d_hist, c->p.histo_max);
            } else
            switch (c->pipe_depth) {
            case 1: if (h) SKM_LAUNCH_EI(1, true); else SKM_LAUNCH_EI(1, false); break;
            case 2: if (h) SKM_LAUNCH_EI(2, true); else SKM_LAUNCH_EI(2, false); break;
            case 8: if (h) SKM_LAUNCH_EI(8, true); else SKM_LAUNCH_EI(8, false); break;
            default: if (h) SKM_LAUNCH_EI(4, true); else SKM_LAUNCH_EI(4, false); break;
            }
#undef SKM_LAUNCH_EI
            c->launches++;
            c->stage_launches[ST_INSERT]++;
            c->insert_bases += std::min(tile_units * 32, sg.n_bytes - u * 32);
        }
        CU(cudaGetLastError());
        c->table_fresh = false;
        note_inserted(c, tile_units * 32);
        u += tile_units;
    }
    return SKM_OK;
}

// regions per owner used by the router: the largest power of two with n_ranks * regions <= kMaxBuckets
uint32_t route_log2_regions(const skm_ctx *c) {
    uint32_t l = 0;
    while (((uint64_t)c->n_ranks << (l + 1)) <= c->max_buckets) l++;
    return l;
}

struct WorkStream {  // selects the stream the bucketing helpers launch on, for the current scope
    skm_ctx *c;
    cudaStream_t prev;
    WorkStream(skm_ctx *c_, cudaStream_t s) : c(c_), prev(c_->work) { c->work = s; }
    ~WorkStream() { c->work = prev; }
};

int32_t sync_dma(skm_ctx *c) {
    CU(cudaStreamSynchronize(c->dma_stream));
    for (cudaStream_t s : c->dma_peer) CU(cudaStreamSynchronize(s));
    return SKM_OK;
}
cudaStream_t dma_stream_for(const skm_ctx *c, uint32_t dst) { return dst < c->dma_peer.size() ? c->dma_peer[dst] : c->dma_stream; }

int32_t sync_all(skm_ctx *c) {
    CU(cudaStreamSynchronize(c->copy_stream));
    CU(cudaStreamSynchronize(c->pack_stream));
    CU(cudaStreamSynchronize(c->part_stream));
    CU(cudaStreamSynchronize(c->sort_stream));
    {
        int32_t rc = sync_dma(c);
        if (rc) return rc;
    }
    CU(cudaStreamSynchronize(c->stream));
    return SKM_OK;
}

// Pass 1 of bucketing segments [s0, s1) of a chunk: per-bucket counts + offsets/cursors on the
// device; the total (and optionally the counts) on the host.
int32_t bucket_count(skm_ctx *c, uint32_t chunk, size_t s0, size_t s1, BucketFn fn, uint32_t n_buckets,
                     uint64_t *total_out, uint64_t *h_counts) {
    const ChunkState &cs = c->chunks[chunk];
    zero_async(c, c->d_bucket_counts, (n_buckets + 1) * sizeof(uint64_t), c->work);
    {
        Span sp(c, ST_COUNT, c->work);
        for (size_t s = s0; s < s1; s++) {
            const Segment &sg = cs.segs[s];
            if (!sg.n_units || !sg.codes) continue;
            bucket_count_kernel<<<grid_for(sg.n_units, 256 * kBucketUnits), 256, n_buckets * sizeof(uint32_t), c->work>>>(
                sg.codes, sg.breaks, 0, sg.n_units, c->p.k, fn, n_buckets, c->d_bucket_counts,
                &c->d_cc[chunk]);
            c->launches++;
            c->stage_launches[ST_COUNT]++;
        }
        bucket_scan_kernel<<<1, kScanThreads, 0, c->work>>>(c->d_bucket_counts, n_buckets, c->d_bucket_offsets,
                                                       c->d_bucket_cursors);
        c->launches++;
    }
    CU(cudaGetLastError());
    if (!total_out) return SKM_OK;  // caller does not need the numbers on the host: stay asynchronous
    CU(cudaMemcpyAsync(c->h_pinned, c->d_bucket_offsets + n_buckets, sizeof(uint64_t),
                       cudaMemcpyDeviceToHost, c->work));
    if (h_counts)
        CU(cudaMemcpyAsync(c->h_pinned + 1, c->d_bucket_counts, n_buckets * sizeof(uint64_t),
                           cudaMemcpyDeviceToHost, c->work));
    CU(cudaStreamSynchronize(c->work));
    *total_out = c->h_pinned[0];
    if (h_counts) memcpy(h_counts, c->h_pinned + 1, n_buckets * sizeof(uint64_t));
    return SKM_OK;
}

// Pass 2: scatter the k-mers into `d_out`, grouped by bucket (uses the cursors of pass 1).
int32_t bucket_scatter(skm_ctx *c, uint32_t chunk, size_t s0, size_t s1, BucketFn fn, uint32_t n_buckets,
                       unsigned long long *d_out) {
    const ChunkState &cs = c->chunks[chunk];
    Span sp(c, ST_PART, c->work);
    const size_t smem = scatter_smem_bytes(n_buckets);
    for (size_t s = s0; s < s1; s++) {
        const Segment &sg = cs.segs[s];
        if (!sg.n_units || !sg.codes) continue;
        bucket_scatter_kernel<false><<<grid_for(sg.n_units, kScatterThreads), kScatterThreads, smem, c->work>>>(
            sg.codes, sg.breaks, 0, sg.n_units, c->p.k, fn, n_buckets, c->d_bucket_cursors, d_out, CapLayout{},
            nullptr);
        c->launches++;
        c->stage_launches[ST_PART]++;
    }
    CU(cudaGetLastError());
    return SKM_OK;
}

// Forget a capped list (waits for its bucketing): frees it on `st` and takes back the windows its
// scatter added to the chunk's counter, so that the exact path can count them again.
void release_list(skm_ctx *c, Segment &sg, cudaStream_t st);
int32_t drop_capped_list(skm_ctx *c, uint32_t chunk, Segment &sg, cudaStream_t st) {
    CU(cudaEventSynchronize(sg.ready));
    uint64_t counted = 0;   // windows the bucketing of this batch added to the chunk's counter
    if (sg.cap) for (uint32_t r = 0; r < sg.n_buckets; r++) counted += sg.h_offsets[r];
    else counted = sg.h_offsets[sg.n_buckets];
    adjust_counter_kernel<<<1, 1, 0, st>>>(&c->d_cc[chunk].n_windows, 0ull - counted);
    c->launches++;
    release_list(c, sg, st);
    return SKM_OK;
}

// One-pass bucketing of one segment into the capped layout (no counting pass): see CapLayout.
int32_t bucket_scatter_capped(skm_ctx *c, uint32_t chunk, const Segment &sg, BucketFn fn, uint32_t n_buckets,
                              unsigned long long *d_out, CapLayout lay) {
    zero_async(c, c->d_bucket_cursors, (n_buckets + 1) * sizeof(uint64_t), c->work);
    Span sp(c, ST_PART, c->work);
    bucket_scatter_kernel<true><<<grid_for(sg.n_units, kScatterThreads), kScatterThreads,
                                  scatter_smem_bytes(n_buckets, true), c->work>>>(
        sg.codes, sg.breaks, 0, sg.n_units, c->p.k, fn, n_buckets, c->d_bucket_cursors, d_out, lay, &c->d_cc[chunk]);
    c->launches++;
    c->stage_launches[ST_PART]++;
    CU(cudaGetLastError());
    return SKM_OK;
}

#define SKM_LAUNCH_RUNS(D, H)                                                                     \
    insert_runs_kernel<D, H><<<grid, 256, 0, c->stream>>>(nullptr, 0, single, n_dev, n_tiles, counter, \
                                                          tref(c), c->d_gc, c->d_hist, c->p.histo_max)
// One flat list (skm_insert_counts, skm_insert_kmers_device): global-memory atomics.
void launch_insert_runs(skm_ctx *c, uint64_t n_tiles, RunDesc single, const unsigned long long *n_dev) {
    const bool h = c->track_histo;
    uint64_t want = (n_tiles + 7) / 8;
    const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)c->sm_count * c->insert_ctas_per_sm));
    unsigned long long *counter = &c->d_gc->scratch[0];
    zero_async(c, counter, sizeof(unsigned long long), c->stream);
    switch (c->pipe_depth) {
    case 2: if (h) SKM_LAUNCH_RUNS(2, true); else SKM_LAUNCH_RUNS(2, false); break;
    case 4: if (h) SKM_LAUNCH_RUNS(4, true); else SKM_LAUNCH_RUNS(4, false); break;
    case 8: if (h) SKM_LAUNCH_RUNS(8, true); else SKM_LAUNCH_RUNS(8, false); break;
    default: if (h) SKM_LAUNCH_RUNS(1, true); else SKM_LAUNCH_RUNS(1, false); break;
    }
    c->launches++;
    c->stage_launches[ST_INSERT]++;
}
#undef SKM_LAUNCH_RUNS

// Insert one flat list (optionally with counts).  n_dev: exact length in device memory (then n
// is an upper bound and the whole list goes in one launch).
int32_t insert_list(skm_ctx *c, const unsigned long long *d_kmers, const uint32_t *d_counts, uint64_t n,
                    const unsigned long long *n_dev = nullptr) {
    c->recount_valid = false;
    uint64_t i = 0;
    while (i < n) {
        uint64_t granted = 0;
        int32_t rc = reserve_headroom(c, n - i, &granted);
        if (rc) return rc;
        rc = ensure_physical(c);  // (after the headroom check: an empty table may just have been re-sized)
        if (rc) return rc;
        granted = std::max<uint64_t>(std::min(granted, n - i), 1);
        {
            Span sp(c, ST_INSERT, c->stream);
            RunDesc single{d_kmers + i, d_counts ? d_counts + i : nullptr, granted, 0};
            launch_insert_runs(c, (granted + kWarpTile - 1) / kWarpTile, single,
                               (i == 0 && granted == n) ? n_dev : nullptr);
            c->insert_kmers += granted;
        }
        CU(cudaGetLastError());
        c->table_fresh = false;
        note_inserted(c, granted);
        i += granted;
    }
    return SKM_OK;
}

// One streaming pass over the table: histogram + totals (+ digest).  Synchronous.
int32_t scan_table(skm_ctx *c, bool want_digest, std::vector<uint64_t> *bins_out) {
    const uint64_t nb = c->p.histo_max + 2;
    {
        int32_t rc0 = ensure_physical(c);
        if (rc0) return rc0;
    }
    zero_async(c, c->d_bins, nb * sizeof(uint64_t), c->stream);
    zero_async(c, c->d_tot, sizeof(HistoTotals), c->stream);
    const uint32_t n_smem_bins = (uint32_t)std::min<uint64_t>(nb, 12288);
    const size_t smem = (kLowBins * 32 + n_smem_bins) * sizeof(uint32_t);
    {
        Span sp(c, ST_HISTO, c->stream);
        histogram_kernel<<<c->sm_count * 4, 512, smem, c->stream>>>(c->table, c->capacity, c->p.histo_max,
                                                                    n_smem_bins, c->d_bins, c->d_tot,
                                                                    want_digest ? 1 : 0);
        c->launches++;
        c->stage_launches[ST_HISTO]++;
    }
    CU(cudaGetLastError());
    if (bins_out) {
        bins_out->assign(nb, 0);
        CU(cudaMemcpyAsync(bins_out->data(), c->d_bins, nb * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaMemcpyAsync(&c->last_tot, c->d_tot, sizeof(HistoTotals), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->have_tot = true;
    return SKM_OK;
}

// Column `chunk_i` of the incremental histogram: the running histogram kept by the
// insert kernels (a stream-ordered copy), or a table scan when tracking is off.
int32_t snapshot_histogram(skm_ctx *c, uint32_t chunk_i) {
    const uint64_t nb = c->p.histo_max + 2;
    if (chunk_i >= c->n_chunks) return fail(c, SKM_ERR_INVALID_ARG, "chunk_index out of range");
    if (!c->track_histo) {
        int32_t rc = scan_table(c, false, &c->histos[chunk_i]);
        if (rc) return rc;
    } else if (c->h_cols) {
        // pinned landing area: the copy is truly asynchronous (a pageable destination would block the
        // host until the inserts queued before it have finished)
        CU(cudaMemcpyAsync(c->h_cols + (size_t)chunk_i * nb, c->d_hist, nb * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                           c->stream));
        c->col_pending[chunk_i] = true;
    } else {
        c->histos[chunk_i].assign(nb, 0);
        CU(cudaMemcpyAsync(c->histos[chunk_i].data(), c->d_hist, nb * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                           c->stream));
    }
    c->have_histo[chunk_i] = true;
    return SKM_OK;
}

// After the main stream has been synchronised: move landed columns into c->histos.
void materialize_cols(skm_ctx *c) {
    const uint64_t nb = c->p.histo_max + 2;
    for (uint32_t i = 0; i < c->n_chunks; i++)
        if (c->col_pending[i]) {
            c->histos[i].assign(c->h_cols + (size_t)i * nb, c->h_cols + (size_t)(i + 1) * nb);
            c->col_pending[i] = false;
        }
}

int32_t check_sticky(skm_ctx *c) {
    // requires stream sync done by the caller
    GlobalCounters gc;
    CU(cudaMemcpy(&gc, c->d_gc, sizeof gc, cudaMemcpyDeviceToHost));
    if (gc.part_full) {
        c->sticky_error = true;
        return fail(c, SKM_ERR_CAPACITY, "a table partition (%u slots) filled up: keys were dropped; raise capacity_hint", kPartSlots);
    }
    if (gc.first_bad != ~0ull) {
        c->sticky_error = true;
        const unsigned ch = (unsigned)(gc.first_bad & 0xFF);
        const unsigned long long pos = gc.first_bad >> 8;
        char shown[8];
        if (ch >= 0x20 && ch < 0x7F)
            snprintf(shown, sizeof shown, "%c", (char)ch);
        else
            snprintf(shown, sizeof shown, "\\x%02x", ch);
        // text of src/kmer/encoding.rs:353-356, plus where it was found
        return fail(c, SKM_ERR_INVALID_BASE,
                    "Invalid character '%s' in sequence. Only ACGTN allowed. (byte %llu of the ingested stream)",
                    shown, pos);
    }
    return SKM_OK;
}

int32_t refresh_chunk_counters(skm_ctx *c) {
    CU(cudaMemcpy(c->h_cc.data(), c->d_cc, c->n_chunks * sizeof(ChunkCounters), cudaMemcpyDeviceToHost));
    return SKM_OK;
}

BucketFn route_fn(const skm_ctx *c) {
    BucketFn fn;
    fn.n_ranks = c->n_ranks;
    fn.log2_regions = route_log2_regions(c);
    return fn;
}
constexpr size_t kOffBlockEntries = 64 * (kMaxBuckets + 1);

uint64_t *alloc_offsets(skm_ctx *c, uint32_t n) {
    if (c->off_blocks.empty() || c->off_used + n > kOffBlockEntries) {
        uint64_t *b = nullptr;
        if (cudaMallocHost((void **)&b, kOffBlockEntries * sizeof(uint64_t)) != cudaSuccess) return nullptr;
        c->off_blocks.push_back(b);
        c->off_used = 0;
    }
    uint64_t *p = c->off_blocks.back() + c->off_used;
    c->off_used += n;
    return p;
}

bool want_partitioned_cap(const skm_ctx *c, uint64_t n_bytes, uint64_t capacity) {
    if (c->p.insert_mode == SKM_INSERT_DIRECT) return false;
    if (c->p.insert_mode == SKM_INSERT_PARTITIONED) return true;
    return capacity * sizeof(Slot) >= (96ull << 20) && n_bytes >= (1ull << 23);
}

bool want_partitioned(const skm_ctx *c, uint64_t n_bytes) {
    if (c->p.insert_mode == SKM_INSERT_DIRECT) return false;
    if (c->p.insert_mode == SKM_INSERT_PARTITIONED) return true;
    // AUTO: partitioned wins once the table no longer fits in L2 and there is enough work to
    // amortise the two bucketing passes (measured crossover); else direct.
    return c->capacity * sizeof(Slot) >= (96ull << 20) && n_bytes >= (1ull << 23);
}

void release_list(skm_ctx *c, Segment &sg, cudaStream_t st) {
    if (sg.list) {
        buf_free(c, sg.list, st);
        c->list_bytes -= std::min<size_t>(c->list_bytes, sg.list_cells * sizeof(uint64_t));
    }
    if (sg.meta) buf_free(c, sg.meta, st);
    if (sg.tile_off) {
        buf_free(c, sg.tile_off, st);
        c->list_bytes -= std::min<size_t>(c->list_bytes, (size_t)sg.max_tiles * ((1u << c->g2) + 1) * sizeof(tile_off_t));
    }
    if (sg.d_counts) cudaFreeAsync(sg.d_counts, st);
    for (auto &ol : sg.owners) {
        if (ol.list) {
            buf_free(c, ol.list, st);
            c->list_bytes -= std::min<size_t>(c->list_bytes, ol.cells * sizeof(uint64_t));
        }
        buf_free(c, ol.meta, st);
        buf_free(c, ol.tile_off, st);
    }
    sg.owners.clear();
    sg.list = nullptr;
    sg.meta = nullptr;
    sg.tile_off = nullptr;
    sg.d_counts = nullptr;
    sg.tiled = false;
    sg.cap = 0;
}

// geometry of the lists the insert kernel reads: on one GPU the buckets of pass A are the table regions; across
// GPUs every owner's slice is re-bucketed into kFineRegions regions of that owner (tile_rebucket_kernel)
uint32_t list_log2_regions(const skm_ctx *c) { return c->n_ranks > 1 && !c->mg_slices ? kFineLog2 : route_log2_regions(c); }
uint32_t list_tile_log2(const skm_ctx *c) { return c->n_ranks > 1 && !c->mg_slices ? kTileLog2 : c->tile_log2; }
ListGeom list_geom(const skm_ctx *c) { return ListGeom{c->n_ranks, list_log2_regions(c), c->g2}; }

// Pass B over `tiles` tile slots of 2^tile_log2 cells: one CTA per tile, or a cluster of 2^(tile_log2 - 13) CTAs
template <int C, uint32_t SP>
cudaError_t launch_cluster_sort(skm_ctx *c, uint32_t tiles, cudaStream_t st, unsigned long long *list, ListMeta m, uint32_t nb, ListGeom geom,
                                tile_off_t *tile_off) {
    // persistent clusters: as many as the device holds at once (a power-of-two index into the cache)
    int &resident = c->sort_clusters[C == 2 ? 0 : C == 4 ? 1 : 2];
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((uint32_t)c->sm_count * 2u / (uint32_t)C * (uint32_t)C);
    cfg.blockDim = dim3(kSortThreads);
    cfg.dynamicSmemBytes = tile_sort_cluster_smem_bytes();
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (resident == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, tile_sort_cluster_kernel<C, SP>, &cfg) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = c->sm_count * 2 / C;
        }
        resident = n;
    }
    cfg.gridDim = dim3((uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(tiles, (uint64_t)resident)) * (uint32_t)C);
    // the tile counter of this launch: a ring of words, so that launches queued behind each other do not share one
    unsigned int *counter = c->d_sort_counters + (c->sort_counter_next++ % skm_ctx::kSortCounters);
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, tile_sort_cluster_kernel<C, SP>, list, m, nb, geom, tile_off, counter);
}

cudaError_t launch_tile_sort(skm_ctx *c, uint32_t tile_log2, uint32_t g2, uint32_t tiles, cudaStream_t st, unsigned long long *list, ListMeta m,
                             uint32_t nb, ListGeom geom, tile_off_t *tile_off) {
    if (tiles == 0) return cudaSuccess;
    switch (tile_log2) {
        // (two sub-buckets per thread in the kernels' scans up to 2^10 sub-buckets, four for 2^11)
        case kTileLog2:
            if (g2 <= 10) tile_sort_kernel<2><<<tiles, kSortThreads, tile_sort_smem_bytes(g2), st>>>(list, m, nb, geom, tile_off);
            else tile_sort_kernel<4><<<tiles, kSortThreads, tile_sort_smem_bytes(g2), st>>>(list, m, nb, geom, tile_off);
            return cudaGetLastError();
        case kTileLog2 + 1:
            return g2 <= 10 ? launch_cluster_sort<2, 2>(c, tiles, st, list, m, nb, geom, tile_off)
                            : launch_cluster_sort<2, 4>(c, tiles, st, list, m, nb, geom, tile_off);
        case kTileLog2 + 2:
            return g2 <= 10 ? launch_cluster_sort<4, 2>(c, tiles, st, list, m, nb, geom, tile_off)
                            : launch_cluster_sort<4, 4>(c, tiles, st, list, m, nb, geom, tile_off);
        case kTileLog2 + 3:
            return g2 <= 10 ? launch_cluster_sort<8, 2>(c, tiles, st, list, m, nb, geom, tile_off)
                            : launch_cluster_sort<8, 4>(c, tiles, st, list, m, nb, geom, tile_off);
        default: return cudaErrorInvalidValue;
    }
}

// Passes A and B for one packed segment, on c->work: bucket its k-mers by (owner, table region)
// into a list — capped one-pass layout, or exact two-pass layout — then sort every tile of every
// bucket by sub-bucket in place (tile_sort_kernel).  `must`: fail instead of skipping when memory
// is short.  Leaves sg.list == nullptr when skipped.
int32_t build_list(skm_ctx *c, uint32_t chunk, size_t seg_index, uint64_t *h_off, bool exact, bool must, int reserve_tables = 3,
                   bool sort = true) {
    Segment &sg = c->chunks[chunk].segs[seg_index];
    const BucketFn fn = route_fn(c);
    const uint32_t nb = c->n_ranks << fn.log2_regions;
    const uint32_t F = 1u << c->g2;
    CapLayout lay{};
    size_t cells;
    uint32_t tpb = 0, max_tiles;
    const uint32_t tl = list_tile_log2(c);
    const uint64_t T = 1ull << tl;
    if (!exact) {
        lay.cap = ((sg.n_bytes / nb + sg.n_bytes / (16ull * nb) + 1024) + 15) & ~15ull;
        lay.ovf_base = lay.cap * nb;
        lay.ovf_cap = 4096;  // any k-mer that lands here sends the segment to the exact path
        cells = (size_t)(lay.ovf_base + lay.ovf_cap);
        tpb = (uint32_t)((lay.cap + T - 1) / T);
        max_tiles = nb * tpb;
    } else {
        cells = sg.n_bytes;
        max_tiles = (uint32_t)(sg.n_bytes / T) + nb + 1;
    }
    const size_t off_bytes = sort ? (size_t)max_tiles * (F + 1) * sizeof(tile_off_t) : 16;
    const size_t need = cells * 8 + off_bytes;
    // budget = device memory that was free when the ctx was created (no cudaMemGetInfo here: it would serialise
    // the ingest path); room is kept for `reserve_tables` tables of the current size (growth) plus slack
    if (!must && c->list_bytes + need + (size_t)reserve_tables * c->capacity * sizeof(Slot) + (4ull << 30) > c->mem_budget) return SKM_OK;
    CU(cudaStreamWaitEvent(c->work, sg.ready, 0));
    unsigned long long *list = nullptr, *meta = nullptr;
    tile_off_t *tile_off = nullptr;
    HostTimer t_malloc(&c->host_ms[4]);
    if (buf_alloc(c, (void **)&list, cells * sizeof(uint64_t), c->work) != cudaSuccess ||
        buf_alloc(c, (void **)&tile_off, off_bytes, c->work) != cudaSuccess ||
        buf_alloc(c, (void **)&meta, list_meta_words(nb) * sizeof(uint64_t), c->work) != cudaSuccess) {
        cudaGetLastError();
        if (list) buf_free(c, list, c->work);
        if (tile_off) buf_free(c, tile_off, c->work);
        return must ? fail(c, SKM_ERR_OOM, "device allocation failed (k-mer list of %llu cells)", (unsigned long long)cells) : SKM_OK;
    }
    t_malloc.~HostTimer();
    t_malloc.acc = &c->host_ms[5];   // (the rest of the function: launches)
    t_malloc.t0 = std::chrono::steady_clock::now();
    sg.list = list;
    sg.meta = meta;
    sg.tile_off = tile_off;
    sg.max_tiles = max_tiles;
    sg.n_buckets = nb;
    sg.h_offsets = h_off;
    sg.list_cells = cells;
    c->list_bytes += need;
    const ListMeta m = list_meta_at(meta, nb);
    int32_t rc;
    if (!exact) {
        rc = bucket_scatter_capped(c, chunk, sg, fn, nb, sg.list, lay);
        if (rc) return rc;
        copy_words_kernel<<<4, 256, 0, c->work>>>(c->d_bucket_cursors, (unsigned long long *)h_off, nb + 1);
        tile_plan_kernel<<<1, 1024, 0, c->work>>>(c->d_bucket_cursors, nb, lay.cap, tpb, m, tl);
        c->launches += 2;
        sg.cap = lay.cap;
        sg.ovf_cap = 0;  // capped lists have no usable overflow run: h_offsets[nb] > 0 => rebuild exactly
    } else {
        rc = bucket_count(c, chunk, seg_index, seg_index + 1, fn, nb, nullptr, nullptr);
        if (rc) return rc;
        rc = bucket_scatter(c, chunk, seg_index, seg_index + 1, fn, nb, sg.list);
        if (rc) return rc;
        copy_words_kernel<<<4, 256, 0, c->work>>>(c->d_bucket_offsets, (unsigned long long *)h_off, nb + 1);
        tile_plan_kernel<<<1, 1024, 0, c->work>>>(c->d_bucket_offsets, nb, 0ull, 0u, m, tl);
        c->launches += 2;
        sg.cap = 0;
        CU(cudaFreeAsync(sg.codes, c->work));
        CU(cudaFreeAsync(sg.breaks, c->work));
        sg.codes = nullptr;
        sg.breaks = nullptr;
    }
    if (!sort) {   // multi-GPU: the coarse list is refined per owner first (refine_owner_lists)
        CU(cudaGetLastError());
        DBG_SYNC("coarse bucketing");
        CU(cudaEventRecord(sg.ready, c->work));
        return SKM_OK;
    }
    // The tile sort runs on its own stream: it is bound by HBM bandwidth, the bucketing of the NEXT
    // batch (same `work` stream otherwise) by instruction issue, so the two share the chip well.
    cudaStream_t sort_st = c->sort_overlap ? c->sort_stream : c->work;
    if (sort_st != c->work) {
        CU(cudaEventRecord(c->ev_sort, c->work));
        CU(cudaStreamWaitEvent(sort_st, c->ev_sort, 0));
    }
    {
        Span sp(c, ST_SORT, sort_st);
        CU(launch_tile_sort(c, tl, c->g2, max_tiles, sort_st, sg.list, m, nb, list_geom(c), sg.tile_off));
        c->launches++;
        c->stage_launches[ST_SORT]++;
    }
    CU(cudaGetLastError());
    sg.tiled = true;
    CU(cudaEventRecord(sg.ready, sort_st));  // `ready` now also covers the list, the tile offsets and the totals
    return SKM_OK;
}

// ---- tiled insert ---------------------------------------------------------------------------------
SegDesc local_desc(const skm_ctx *c, const Segment &sg, uint32_t chunk_rel) {
    const ListMeta m = list_meta_at(sg.meta, sg.n_buckets);
    return SegDesc{sg.list, sg.tile_off, m.tile_begin, m.cell_begin, c->p.rank << route_log2_regions(c), chunk_rel, 0u, 0u};
}

// dynamic shared memory one insert CTA may use so that kInsCtasPerSm of them fit an SM (227 KiB, 1 KiB reserved per CTA)
constexpr size_t kTileInsertSmemBudget = (227 * 1024 - kInsCtasPerSm * 1024) / kInsCtasPerSm - 1024;

uint32_t tiled_k_low(const skm_ctx *c, uint32_t n_chunks_l, bool histo) {
    // as many low bins in shared memory as fit beside the partition with 3 CTAs per SM
    uint32_t k = 1024;
    while (k > 16 && tile_insert_smem_bytes(n_chunks_l, k, histo) > kTileInsertSmemBudget) k >>= 1;
    const uint64_t want = c->p.histo_max + 2;   // no point in more bins than the histogram has
    while (k > 16 && (k >> 1) >= want) k >>= 1;
    return k;
}
constexpr uint32_t kMaxChunksPerLaunch = 64;

int32_t ensure_tiled_buffers(skm_ctx *c, uint32_t n_chunks_l, size_t seg_bytes) {
    const size_t nbins = c->p.histo_max + 2;
    if (c->delta_words < (size_t)n_chunks_l * nbins) {
        if (c->d_delta) CU(cudaFreeAsync(c->d_delta, c->stream));
        c->d_delta = nullptr;
        CU(cudaMallocAsync((void **)&c->d_delta, (size_t)n_chunks_l * nbins * sizeof(uint64_t), c->stream));
        c->delta_words = (size_t)n_chunks_l * nbins;
    }
    if (!c->d_recount) CU(cudaMalloc((void **)&c->d_recount, nbins * sizeof(uint64_t)));
    if (!c->d_fail) CU(cudaMalloc((void **)&c->d_fail, (size_t)skm_ctx::kFailCap * sizeof(uint32_t)));
    if (c->segs_bytes < seg_bytes) {
        CU(cudaStreamSynchronize(c->stream));
        if (c->h_segs) CU(cudaFreeHost(c->h_segs));
        if (c->d_segs) CU(cudaFree(c->d_segs));
        c->h_segs = c->d_segs = nullptr;
        const size_t cap = std::max<size_t>(seg_bytes * 2, 64 * 1024);
        CU(cudaMallocHost(&c->h_segs, cap));
        CU(cudaMalloc(&c->d_segs, cap));
        c->segs_bytes = cap;
    }
    return SKM_OK;
}

// One launch (plus retries after growth) of tile_insert_kernel over the given lists, which hold
// chunks [chunk0, chunk0 + n_chunks_l) in chunk order.  Histogram columns of those chunks are
// produced when the ctx tracks the histogram.  Synchronises the main stream.
constexpr int32_t kSpecAborted = -1;   // launch_tiled (internal; the SKM_ERR_* codes are positive): the speculative launch
                                       // was called off on the device, nothing changed

int32_t launch_tiled(skm_ctx *c, const std::vector<SegDesc> &segs, uint32_t chunk0, uint32_t n_chunks_l,
                     const unsigned long long *abort_if = nullptr) {
    const bool histo = c->track_histo;
    const size_t nbins = c->p.histo_max + 2;
    const uint32_t n = (uint32_t)segs.size();
    const size_t seg_bytes = (size_t)n * sizeof(SegDesc) + ((size_t)n_chunks_l + 2) * sizeof(uint32_t) + 16;
    int32_t rc = ensure_tiled_buffers(c, n_chunks_l, seg_bytes);
    if (rc) return rc;
    // descriptors -> pinned -> device (by a kernel: never behind the read batches on a copy engine)
    SegDesc *hd = (SegDesc *)c->h_segs;
    uint32_t *hfirst = (uint32_t *)(hd + n);
    const ListGeom geom = list_geom(c);
    uint32_t at = 0;
    for (uint32_t ch = 0; ch <= n_chunks_l; ch++) {
        while (at < n && segs[at].chunk < ch) at++;
        hfirst[ch] = at;
    }
    hfirst[n_chunks_l] = n;
    for (uint32_t i = 0; i < n; i++) hd[i] = segs[i];
    {
        unsigned long long *h_dev = nullptr;
        CU(cudaHostGetDevicePointer((void **)&h_dev, c->h_segs, 0));
        const uint32_t words = (uint32_t)((seg_bytes + 7) / 8);
        copy_words_kernel<<<8, 256, 0, c->stream>>>(h_dev, (unsigned long long *)c->d_segs, words);
        c->launches++;
    }
    InsertLaunch L{};
    L.n_ranks = c->n_ranks;
    L.g1 = geom.g1;
    L.g2 = geom.g2;
    L.tile_log2 = list_tile_log2(c);
    L.segs = (const SegDesc *)c->d_segs;
    L.chunk_first_seg = (const uint32_t *)((const SegDesc *)c->d_segs + n);
    L.n_segs = n;
    L.n_chunks = n_chunks_l;
    L.k_low = tiled_k_low(c, n_chunks_l, histo);
    L.max_occupied = (uint32_t)(kHardLoad * kPartSlots);
    L.g_delta = c->d_delta;
    L.g_recount = c->d_recount;
    L.histo_max = c->p.histo_max;
    L.gc = c->d_gc;
    L.tot = c->d_tot;
    L.part_counter = &c->d_gc->scratch[0];
    L.fail_list = c->d_fail;
    L.fail_cap = skm_ctx::kFailCap;
    L.abort_if = abort_if;
    const size_t smem = tile_insert_smem_bytes(n_chunks_l, L.k_low, histo);
    if (histo) zero_async(c, c->d_delta, (size_t)n_chunks_l * nbins * sizeof(uint64_t), c->stream);
    zero_async(c, c->d_recount, nbins * sizeof(uint64_t), c->stream);
    zero_async(c, c->d_tot, sizeof(HistoTotals), c->stream);
    std::vector<uint32_t> part_ids;  // empty: all partitions
    for (uint32_t attempt = 0;; attempt++) {
        L.table = c->table;
        L.log2cap = c->log2cap;
        L.fresh = c->table_fresh ? 1 : 0;
        const uint64_t n_parts_all = c->capacity >> kPartLog2;
        if (part_ids.empty()) {
            L.part_ids = nullptr;
            L.n_parts = n_parts_all;
        } else {
            if (c->d_part_ids) CU(cudaFreeAsync(c->d_part_ids, c->stream));
            CU(cudaMallocAsync((void **)&c->d_part_ids, part_ids.size() * sizeof(uint32_t), c->stream));
            CU(cudaMemcpyAsync(c->d_part_ids, part_ids.data(), part_ids.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
            L.part_ids = c->d_part_ids;
            L.n_parts = part_ids.size();
        }
        // how many (list, bucket) pairs one partition reads: the kernel stages them in shared memory
        const uint32_t pbits = c->log2cap - kPartLog2;
        const uint64_t nbr = pbits >= geom.g1 ? 1ull : (1ull << (geom.g1 - pbits));
        if ((uint64_t)n * nbr > kMaxVseg)
            return fail(c, SKM_ERR_STATE, "internal: %u lists x %llu buckets per partition exceed one launch", n, (unsigned long long)nbr);
        zero_async(c, L.part_counter, sizeof(unsigned long long), c->stream);
        zero_async(c, &c->d_gc->n_failed, 2 * sizeof(unsigned long long), c->stream);  // n_failed, fatal
        int occ = 0;
        if (histo) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tile_insert_kernel<true>, kInsThreads, smem));
        else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tile_insert_kernel<false>, kInsThreads, smem));
        const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(L.n_parts, (uint64_t)c->sm_count * std::max(occ, 1)));
        DBG_SYNC("descriptor upload before tile_insert");
        if (getenv("SKM_DEBUG"))
            std::fprintf(stderr, "[skm] tile_insert: grid %u (occ %d) smem %zu k_low %u chunks %u lists %u parts %llu g1 %u g2 %u log2cap %u fresh %d attempt %u\n",
                         grid, occ, smem, L.k_low, n_chunks_l, n, (unsigned long long)L.n_parts, L.g1, L.g2, L.log2cap, L.fresh, attempt);
        {
            Span sp(c, ST_INSERT, c->stream);
            if (histo) tile_insert_kernel<true><<<grid, kInsThreads, smem, c->stream>>>(L);
            else tile_insert_kernel<false><<<grid, kInsThreads, smem, c->stream>>>(L);
            c->launches++;
            c->stage_launches[ST_INSERT]++;
            c->n_tiled_launches++;
        }
        CU(cudaGetLastError());
        const bool was_fresh = c->table_fresh;
        GlobalCounters *hgc = (GlobalCounters *)c->h_pinned;
        CU(cudaMemcpyAsync(hgc, c->d_gc, sizeof(GlobalCounters), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (abort_if && hgc->pad) {   // (abort_if == &d_gc->pad) called off on the device: no partition was touched
            c->n_tiled_launches--;
            return kSpecAborted;
        }
        c->table_fresh = false;
        c->table_zombie = false;  // every partition was written (failed ones: written empty)
        c->distinct_ub = hgc->n_distinct;
        for (auto &pnd : c->snap_pending) pnd = false;
        const uint64_t n_failed = hgc->n_failed;
        if (n_failed == 0) break;
        if (hgc->fatal)
            return fail(c, SKM_ERR_CAPACITY,
                        "a table partition filled up after publishing histogram moves (capacity_hint %llu far too small for this input)",
                        (unsigned long long)c->p.capacity_hint);
        if (attempt >= 24) return fail(c, SKM_ERR_CAPACITY, "table growth did not converge");
        // grow (the partitions that succeeded keep their keys: rehash), then retry the children of the failed ones
        std::vector<uint32_t> failed(std::min<uint64_t>(n_failed, skm_ctx::kFailCap));
        const bool all = n_failed > skm_ctx::kFailCap;
        if (!all) CU(cudaMemcpy(failed.data(), c->d_fail, failed.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        // size the new table from what was seen: failed partitions hold > 0.9 * kPartSlots keys each
        uint32_t nl = c->log2cap + 1;
        if (n_failed * 2 > n_parts_all && was_fresh) nl = c->log2cap + 2;
        {
            size_t free_b = 0, total_b = 0;
            CU(cudaMemGetInfo(&free_b, &total_b));
            if (((size_t)sizeof(Slot) << nl) + (1ull << 30) > free_b) nl = c->log2cap + 1;
            if (((size_t)sizeof(Slot) << nl) + (1ull << 30) > free_b)
                return fail(c, SKM_ERR_CAPACITY, "out of device memory growing the table to 2^%u slots", nl);
        }
        const uint32_t shift = nl - c->log2cap;
        rc = grow_table(c, nl);
        if (rc) return rc;
        c->n_tiled_retries++;
        std::vector<uint32_t> next;
        if (all) {
            // too many to list: every partition that is still empty of this launch's k-mers is unknown
            return fail(c, SKM_ERR_CAPACITY, "more than %u table partitions overflowed; raise capacity_hint", skm_ctx::kFailCap);
        }
        next.reserve(failed.size() << shift);
        for (uint32_t q : failed)
            for (uint32_t j = 0; j < (1u << shift); j++) next.push_back((q << shift) | j);
        part_ids.swap(next);
    }
    if (c->d_part_ids) {
        CU(cudaFreeAsync(c->d_part_ids, c->stream));
        c->d_part_ids = nullptr;
    }
    if (histo) {
        // columns of chunks chunk0 .. chunk0 + n_chunks_l - 1 (in place over the moves), running histogram updated
        hist_columns_kernel<<<std::min<uint32_t>(grid_for(nbins, 256), 1024), 256, 0, c->stream>>>(c->d_delta, n_chunks_l, nbins, c->d_hist, c->d_delta);
        c->launches++;
        CU(cudaGetLastError());
        for (uint32_t i = 0; i < n_chunks_l; i++) {
            const uint32_t ch = chunk0 + i;
            if (c->h_cols) {
                CU(cudaMemcpyAsync(c->h_cols + (size_t)ch * nbins, c->d_delta + (size_t)i * nbins, nbins * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
                c->col_pending[ch] = true;
            } else {
                c->histos[ch].assign(nbins, 0);
                CU(cudaMemcpyAsync(c->histos[ch].data(), c->d_delta + (size_t)i * nbins, nbins * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
            }
            c->have_histo[ch] = true;
        }
    }
    c->recount_valid = true;
    return SKM_OK;
}

int32_t ship_segment(skm_ctx *c, uint32_t chunk, size_t seg_index, bool sizes_on_host);
int32_t ship_slices(skm_ctx *c, uint32_t chunk, size_t seg_index, bool sizes_on_host);
int32_t refine_owner_lists(skm_ctx *c, uint32_t chunk, size_t seg_index, bool exact);

// Bucket a freshly packed segment by table region right away (single GPU).  The work is queued
// behind the pack kernel on the ctx's stream, so it overlaps the next batch's host-to-device copy,
// and skm_finalize only has the inserts left.  Costs 8 B per position of HBM instead of 0.375 B;
// skipped when memory is short (the segment then stays packed and is bucketed at finalize).
int32_t eager_partition(skm_ctx *c, uint32_t chunk, size_t seg_index, bool force = false) {
    Segment &sg = c->chunks[chunk].segs[seg_index];
    if (!sg.codes || sg.list) return SKM_OK;
    if (!force) {
        // multi-GPU: routing needs the buckets anyway; single GPU: only when partitioned insert pays off
        if (!c->eager || (c->n_ranks == 1 && !want_partitioned(c, sg.n_bytes))) return SKM_OK;
    }
    const BucketFn fn = route_fn(c);  // (owner, region); a single GPU is the n_ranks == 1 case
    const uint32_t nb = c->n_ranks << fn.log2_regions;
    uint64_t *h_off = alloc_offsets(c, nb + 1);
    if (!h_off) return force ? fail(c, SKM_ERR_OOM, "pinned allocation failed") : SKM_OK;
    // bucket by (owner, table region) and tile-sort the buckets (tiled insert).  One pass (capped
    // layout) unless the caller needs the exact layout: `force` = a capped list overflowed.
    // The capped layout pays a fixed slack per bucket (1024 cells): batches too small to amortise it, and
    // batches whose bucket counts could overflow 32 bits, take the exact two-pass layout.
    const bool exact = force || !c->capped || sg.n_bytes >= (3ull << 30) || sg.n_bytes / nb < 4096;
    // (a table sized from a capacity_hint is not expected to grow: keep room for it alone)
    int32_t rc;
    {
        HostTimer t(&c->host_ms[0]);
        rc = build_list(c, chunk, seg_index, h_off, exact, force, c->p.capacity_hint ? 1 : 3, /*sort=*/c->n_ranks == 1 || c->mg_slices);
    }
    if (rc || c->n_ranks == 1 || !sg.list) return rc;
    if (c->mg_slices) {
        HostTimer t(&c->host_ms[3]);
        return ship_slices(c, chunk, seg_index, /*sizes_on_host=*/false);  // the exchange starts at ingest time (capped lists)
    }
    {
        HostTimer t(&c->host_ms[2]);
        rc = refine_owner_lists(c, chunk, seg_index, exact);
    }
    if (rc) return rc;
    HostTimer t(&c->host_ms[3]);
    return ship_segment(c, chunk, seg_index, /*sizes_on_host=*/false);  // the exchange starts at ingest time (capped lists)
}

// Multi-GPU, after pass A (coarse list of segment `sg`, bucketed by (owner, coarse region)): one OwnerList per
// rank — re-bucketed into that owner's kFineRegions table regions (tile_rebucket_kernel), tile-sorted.  The
// coarse list is freed.  `exact`: exact two-pass layouts (small batches, or the retry of an overflow).
int32_t refine_owner_lists(skm_ctx *c, uint32_t chunk, size_t seg_index, bool exact) {
    Segment &sg = c->chunks[chunk].segs[seg_index];
    const uint32_t N = c->n_ranks, g1c = route_log2_regions(c), nb_in = N << g1c;
    const uint32_t F = 1u << c->g2;
    const ListMeta m_in = list_meta_at(sg.meta, nb_in);
    const uint64_t n_exp = sg.n_bytes / N + sg.n_bytes / (8ull * N) + 4096;   // an owner's share, with room for imbalance
    exact = exact || n_exp / kFineRegions < 4096;
    sg.owners.assign(N, OwnerList{});
    const size_t smem = tile_rebucket_smem_bytes();
    // every tile slot of the coarse list (grid size; the kernel skips the empty ones)
    const uint32_t grid_in = sg.cap ? nb_in * (uint32_t)((sg.cap + kTile - 1) / kTile) : (uint32_t)(sg.n_bytes / kTile) + nb_in + 1;
    const uint32_t W = kFineRegions + 1;   // words per owner in the scratch arrays and in h_off
    uint64_t *h_off_all = alloc_offsets(c, N * W);
    if (!h_off_all) return fail(c, SKM_ERR_OOM, "pinned allocation failed");
    CapLayout lay{};
    if (!exact) lay.cap = ((n_exp / kFineRegions + n_exp / (16ull * kFineRegions) + 1024) + 15) & ~15ull;
    OwnerArrays oa{};
    for (uint32_t o = 0; o < N; o++) {
        OwnerList &ol = sg.owners[o];
        ol.h_off = h_off_all + (size_t)o * W;
        if (!exact) {
            ol.cells = (size_t)lay.cap * kFineRegions;
            ol.max_tiles = kFineRegions * (uint32_t)((lay.cap + kTile - 1) / kTile);
        } else {
            ol.cells = (size_t)sg.n_bytes + 16;   // (an owner cannot get more than the batch holds)
            ol.max_tiles = (uint32_t)(ol.cells / kTile) + kFineRegions + 1;
        }
        ol.cap = lay.cap;
        const size_t off_bytes = (size_t)ol.max_tiles * (F + 1) * sizeof(tile_off_t);
        {
            HostTimer t_alloc(&c->host_ms[1]);
            if (buf_alloc(c, (void **)&ol.list, ol.cells * sizeof(uint64_t), c->work) != cudaSuccess ||
                buf_alloc(c, (void **)&ol.tile_off, off_bytes, c->work) != cudaSuccess ||
                buf_alloc(c, (void **)&ol.meta, list_meta_words(kFineRegions) * sizeof(uint64_t), c->work) != cudaSuccess) {
                cudaGetLastError();
                return fail(c, SKM_ERR_OOM, "device allocation failed (owner list of %zu cells)", ol.cells);
            }
        }
        c->list_bytes += ol.cells * sizeof(uint64_t);
        oa.list[o] = ol.list;
        oa.meta[o] = ol.meta;
        oa.tile_off[o] = ol.tile_off;
    }
    // one launch of each kernel serves all owners (owner = coarse bucket >> g1c)
    {
        Span sp(c, ST_PART, c->work);
        if (!exact) {
            zero_async(c, c->d_bucket_cursors, (size_t)N * W * sizeof(uint64_t), c->work);
            tile_rebucket_kernel<0><<<grid_in, kSortThreads, smem, c->work>>>(sg.list, m_in, nb_in, g1c, N, c->d_bucket_cursors, oa, lay.cap);
            copy_words_kernel<<<8, 256, 0, c->work>>>(c->d_bucket_cursors, (unsigned long long *)h_off_all, N * W);
            tile_plan_owners_kernel<<<N, 1024, 0, c->work>>>(c->d_bucket_cursors, kFineRegions, lay.cap,
                                                             (uint32_t)((lay.cap + kTile - 1) / kTile), oa);
            c->launches += 3;
        } else {
            zero_async(c, c->d_bucket_counts, (size_t)N * W * sizeof(uint64_t), c->work);
            tile_rebucket_kernel<1><<<grid_in, kSortThreads, smem, c->work>>>(sg.list, m_in, nb_in, g1c, N, c->d_bucket_counts, oa, 0ull);
            bucket_scan_kernel<<<N, kScanThreads, 0, c->work>>>(c->d_bucket_counts, kFineRegions, c->d_bucket_offsets, c->d_bucket_cursors);
            tile_rebucket_kernel<2><<<grid_in, kSortThreads, smem, c->work>>>(sg.list, m_in, nb_in, g1c, N, c->d_bucket_cursors, oa, 0ull);
            copy_words_kernel<<<8, 256, 0, c->work>>>(c->d_bucket_offsets, (unsigned long long *)h_off_all, N * W);
            tile_plan_owners_kernel<<<N, 1024, 0, c->work>>>(c->d_bucket_offsets, kFineRegions, 0ull, 0u, oa);
            c->launches += 5;
        }
        c->stage_launches[ST_PART]++;
    }
    CU(cudaGetLastError());
    DBG_SYNC(exact ? "tile_rebucket (exact)" : "tile_rebucket (capped)");
    {
        cudaStream_t sort_st = c->sort_overlap ? c->sort_stream : c->work;
        if (sort_st != c->work) {
            CU(cudaEventRecord(c->ev_sort, c->work));
            CU(cudaStreamWaitEvent(sort_st, c->ev_sort, 0));
        }
        Span sp(c, ST_SORT, sort_st);
        tile_sort_owners_kernel<<<dim3(sg.owners[0].max_tiles, N), kSortThreads, tile_sort_smem_bytes(c->g2), sort_st>>>(
            oa, kFineRegions, list_geom(c));
        c->launches++;
        c->stage_launches[ST_SORT]++;
    }
    CU(cudaGetLastError());
    DBG_SYNC("tile_sort of the owner lists");
    // the coarse list has been consumed (on `work`); the packed form stays for the overflow retry
    buf_free(c, sg.list, c->work);
    buf_free(c, sg.meta, c->work);
    buf_free(c, sg.tile_off, c->work);
    c->list_bytes -= std::min<size_t>(c->list_bytes, sg.list_cells * sizeof(uint64_t) + 16);
    sg.list = nullptr;
    sg.meta = nullptr;
    sg.tile_off = nullptr;
    sg.tiled = true;
    cudaStream_t last = c->sort_overlap ? c->sort_stream : c->work;
    if (last != c->work) {   // everything queued on `work` so far happens before `ready` fires
        CU(cudaEventRecord(c->ev_sort, c->work));
        CU(cudaStreamWaitEvent(last, c->ev_sort, 0));
    }
    CU(cudaEventRecord(sg.ready, last));
    return SKM_OK;
}

// Multi-GPU: push the other owners' lists of a batch into their receive arenas (peer copies by the copy
// engines, on the dma stream, ordered after the batch's `ready` event).  A capped list has a fixed geometry,
// so nothing here waits for the device; exact lists are shipped from skm_mg_finalize, where their sizes are
// known on the host.
int32_t ship_segment(skm_ctx *c, uint32_t chunk, size_t seg_index, bool sizes_on_host) {
    Segment &sg = c->chunks[chunk].segs[seg_index];
    if (sg.owners.empty()) return SKM_OK;
    const uint32_t me = c->p.rank, N = c->n_ranks;
    for (uint32_t o = 0; o < N; o++)
        if (o != me && !c->mg_peer[o]) return SKM_OK;  // arenas not wired yet: shipped at finalize
    const uint32_t F = 1u << c->g2;
    auto take = [&](uint32_t o, size_t bytes, uint64_t *off) -> bool {
        const size_t at = (c->mg_cursor[o] + 255) & ~(size_t)255;
        if (at + bytes > c->mg_sub_bytes) return false;
        *off = at;
        c->mg_cursor[o] = at + bytes;
        return true;
    };
    for (uint32_t i = 1; i < N; i++) {
        const uint32_t o = (me + i) % N;  // stagger the destinations across ranks
        OwnerList &ol = sg.owners[o];
        if (ol.shipped || !ol.list) continue;
        if (!ol.cap && !sizes_on_host) continue;   // exact layout: its offsets are still on their way to the host
        uint64_t n_cells, n_tiles;
        if (ol.cap) {
            n_cells = ol.cells;
            n_tiles = ol.max_tiles;
        } else {
            // exact layout: sizes from the offsets (the caller has waited for `ready`)
            n_cells = ol.h_off[kFineRegions];
            n_tiles = 0;
            for (uint32_t b = 0; b < kFineRegions; b++) n_tiles += (ol.h_off[b + 1] - ol.h_off[b] + kTile - 1) / kTile;
        }
        cudaStream_t ds = dma_stream_for(c, o);
        CU(cudaStreamWaitEvent(ds, sg.ready, 0));
        MgRecord rec{};
        rec.src = me;
        rec.dst = o;
        rec.chunk = chunk;
        rec.regions = kFineRegions;
        rec.n_tiles = (uint32_t)n_tiles;
        rec.n_cells = n_cells;
        const size_t b_cells = n_cells * 8, b_off = n_tiles * (F + 1) * sizeof(tile_off_t);
        const size_t b_cb = (size_t)kFineRegions * 8, b_tb = (size_t)(kFineRegions + 1) * 4;
        if (!take(o, b_cells, &rec.off_cells) || !take(o, b_off, &rec.off_tile_off) || !take(o, b_cb, &rec.off_cell_begin) ||
            !take(o, b_tb, &rec.off_tile_begin))
            return fail(c, SKM_ERR_OOM, "receive arena of rank %u is too small for rank %u's k-mers (%zu bytes per source): create larger arenas",
                        o, me, c->mg_sub_bytes);
        const ListMeta m = list_meta_at(ol.meta, kFineRegions);
        uint8_t *base = c->mg_peer[o] + (size_t)me * c->mg_sub_bytes;
        if (b_cells) CU(cudaMemcpyAsync(base + rec.off_cells, ol.list, b_cells, cudaMemcpyDefault, ds));
        if (b_off) CU(cudaMemcpyAsync(base + rec.off_tile_off, ol.tile_off, b_off, cudaMemcpyDefault, ds));
        CU(cudaMemcpyAsync(base + rec.off_cell_begin, m.cell_begin, b_cb, cudaMemcpyDefault, ds));
        CU(cudaMemcpyAsync(base + rec.off_tile_begin, m.tile_begin, b_tb, cudaMemcpyDefault, ds));
        c->mg_bytes_sent += b_cells + b_off + b_cb + b_tb;
        ol.rec = c->mg_sent.size();
        ol.shipped = true;
        c->mg_sent.push_back(rec);
    }
    return SKM_OK;
}

// Multi-GPU, slices: the batch has ONE list, bucketed by (owner, coarse region) and tile-sorted down to the owners'
// partitions; owner o's k-mers are the buckets [o << g1, (o + 1) << g1) — contiguous cells, contiguous tile
// slots, contiguous metadata.  Each of the four pieces goes to the owner's receive arena as one peer copy; the
// receiver reads them with SegDesc::rel = 1 (indices relative to the slice's first entries).
int32_t ship_slices(skm_ctx *c, uint32_t chunk, size_t seg_index, bool sizes_on_host) {
    Segment &sg = c->chunks[chunk].segs[seg_index];
    if (!sg.list || !sg.tiled || sg.shipped) return SKM_OK;
    const uint32_t me = c->p.rank, N = c->n_ranks;
    for (uint32_t o = 0; o < N; o++)
        if (o != me && !c->mg_peer[o]) return SKM_OK;  // arenas not wired yet: shipped at finalize
    if (!sg.cap && !sizes_on_host) return SKM_OK;      // exact layout: its offsets are still on their way to the host
    const uint32_t g1c = route_log2_regions(c), R = 1u << g1c, F = 1u << c->g2, tl = list_tile_log2(c);
    const uint64_t T = 1ull << tl;
    // first cell / first tile slot of every owner's slice
    std::vector<uint64_t> cell0(N + 1), tile0(N + 1);
    if (sg.cap) {
        const uint64_t tpb = (sg.cap + T - 1) / T;
        for (uint32_t o = 0; o <= N; o++) {
            cell0[o] = (uint64_t)o * R * sg.cap;
            tile0[o] = (uint64_t)o * R * tpb;
        }
    } else {
        uint64_t tiles = 0;
        for (uint32_t b = 0; b < sg.n_buckets; b++) {
            if (b % R == 0) {
                cell0[b / R] = sg.h_offsets[b];
                tile0[b / R] = tiles;
            }
            tiles += (sg.h_offsets[b + 1] - sg.h_offsets[b] + T - 1) / T;
        }
        cell0[N] = sg.h_offsets[sg.n_buckets];
        tile0[N] = tiles;
    }
    auto take = [&](uint32_t o, size_t bytes, uint64_t *off) -> bool {
        const size_t at = (c->mg_cursor[o] + 255) & ~(size_t)255;
        if (at + bytes > c->mg_sub_bytes) return false;
        *off = at;
        c->mg_cursor[o] = at + bytes;
        return true;
    };
    const ListMeta m = list_meta_at(sg.meta, sg.n_buckets);
    sg.rec_first = c->mg_sent.size();
    for (uint32_t i = 1; i < N; i++) {
        const uint32_t o = (me + i) % N;  // stagger the destinations across ranks
        MgRecord rec{};
        rec.src = me;
        rec.dst = o;
        rec.chunk = chunk;
        rec.regions = R;
        rec.n_tiles = (uint32_t)(tile0[o + 1] - tile0[o]);
        rec.n_cells = cell0[o + 1] - cell0[o];
        const size_t b_cells = rec.n_cells * 8, b_off = (size_t)rec.n_tiles * (F + 1) * sizeof(tile_off_t);
        const size_t b_cb = (size_t)R * 8, b_tb = (size_t)(R + 1) * 4;
        if (!take(o, b_cells, &rec.off_cells) || !take(o, b_off, &rec.off_tile_off) || !take(o, b_cb, &rec.off_cell_begin) ||
            !take(o, b_tb, &rec.off_tile_begin))
            return fail(c, SKM_ERR_OOM, "receive arena of rank %u is too small for rank %u's k-mers (%zu bytes per source): create larger arenas",
                        o, me, c->mg_sub_bytes);
        uint8_t *base = c->mg_peer[o] + (size_t)me * c->mg_sub_bytes;
        cudaStream_t ds = dma_stream_for(c, o);   // one copy stream per destination: the peers' copies run side by side
        CU(cudaStreamWaitEvent(ds, sg.ready, 0));
        if (b_cells) CU(cudaMemcpyAsync(base + rec.off_cells, sg.list + cell0[o], b_cells, cudaMemcpyDefault, ds));
        if (b_off) CU(cudaMemcpyAsync(base + rec.off_tile_off, sg.tile_off + tile0[o] * (F + 1), b_off, cudaMemcpyDefault, ds));
        CU(cudaMemcpyAsync(base + rec.off_cell_begin, m.cell_begin + (size_t)o * R, b_cb, cudaMemcpyDefault, ds));
        CU(cudaMemcpyAsync(base + rec.off_tile_begin, m.tile_begin + (size_t)o * R, b_tb, cudaMemcpyDefault, ds));
        c->mg_bytes_sent += b_cells + b_off + b_cb + b_tb;
        c->mg_sent.push_back(rec);
    }
    sg.shipped = true;
    return SKM_OK;
}

// Stage one batch that is already in device memory: pack it on `pack_stream` (default c->work;
// `packed` fires once the raw bytes are no longer needed), then bucket it eagerly on c->work.
int32_t stage_device(skm_ctx *c, uint32_t chunk, const uint8_t *d_seqs, uint64_t n_bytes,
                     cudaStream_t pack_stream = nullptr, cudaEvent_t packed = nullptr) {
    if (n_bytes == 0) return SKM_OK;
    if (!pack_stream) pack_stream = c->work;
    Segment sg;
    sg.n_bytes = n_bytes;
    sg.n_units = (n_bytes + 31) / 32;
    CU(cudaMallocAsync((void **)&sg.codes, sg.n_units * sizeof(uint64_t), pack_stream));
    CU(cudaMallocAsync((void **)&sg.breaks, sg.n_units * sizeof(uint32_t), pack_stream));
    {
        Span sp(c, ST_PACK, pack_stream);
        pack_kernel<<<grid_for(sg.n_units, 256 * kPackUnits), 256, 0, pack_stream>>>(
            d_seqs, n_bytes, c->pos_base, sg.codes, sg.breaks, sg.n_units, &c->d_cc[chunk], c->d_gc);
        c->launches++;
        c->stage_launches[ST_PACK]++;
    }
    CU(cudaGetLastError());
    sg.ready = get_event(c);
    CU(cudaEventRecord(sg.ready, pack_stream));  // packed: the raw buffer may be overwritten, the codes read
    if (packed) CU(cudaEventRecord(packed, pack_stream));
    c->pos_base += n_bytes;
    c->chunks[chunk].segs.push_back(sg);
    c->chunks[chunk].n_bytes += n_bytes;
    const size_t idx = c->chunks[chunk].segs.size() - 1;
    int32_t rc = eager_partition(c, chunk, idx);
    if (rc) return rc;
    return SKM_OK;
}

int32_t check_ingest_args(skm_ctx *c, uint32_t chunk, const void *p, uint64_t n) {
    if (!c) return SKM_ERR_INVALID_ARG;
    if (c->finalized) return fail(c, SKM_ERR_STATE, "ingest after finalize");
    if (chunk >= c->n_chunks) return fail(c, SKM_ERR_INVALID_ARG, "chunk_index %u out of range (n_chunks %u)", chunk, c->n_chunks);
    if (n && !p) return fail(c, SKM_ERR_INVALID_ARG, "null buffer");
    return SKM_OK;
}

// Ascending-key order for exported (k-mer, count) pairs: a device radix sort over the 2k significant
// bits (CUB, a library call: presentation order of the read side, not the counting path).
int32_t sort_pairs_device(skm_ctx *c, unsigned long long *d_keys, uint32_t *d_counts, uint64_t n) {
    if (n < 2) return SKM_OK;
    if (n > (uint64_t)INT32_MAX * 2) return fail(c, SKM_ERR_INVALID_ARG, "sorted export limited to 2^32 entries per call");
    unsigned long long *k2 = nullptr;
    uint32_t *c2 = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    const int end_bit = std::min<int>(64, 2 * (int)c->p.k);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, k2, d_counts, c2, n, 0, end_bit, c->stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&k2, n * sizeof(uint64_t), c->stream);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&c2, n * sizeof(uint32_t), c->stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&tmp, std::max<size_t>(tmp_bytes, 1), c->stream);
    if (e == cudaSuccess)
        e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d_keys, k2, d_counts, c2, n, 0, end_bit, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_keys, k2, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_counts, c2, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream);
    if (k2) cudaFreeAsync(k2, c->stream);
    if (c2) cudaFreeAsync(c2, c->stream);
    if (tmp) cudaFreeAsync(tmp, c->stream);
    c->launches += 4;
    if (e != cudaSuccess)
        return fail(c, e == cudaErrorMemoryAllocation ? SKM_ERR_OOM : SKM_ERR_CUDA, "device sort failed: %s", cudaGetErrorString(e));
    return SKM_OK;
}

// A private stream for one read-side call (returned to the pool by the caller); call with c->mu held.
cudaStream_t take_read_stream(skm_ctx *c) {
    if (!c->read_streams.empty()) {
        cudaStream_t s = c->read_streams.back();
        c->read_streams.pop_back();
        return s;
    }
    cudaStream_t s = nullptr;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    return s;
}

// Stream-ordered temporary: freed on every exit path (the CU() macro returns early on errors).
template <class T>
struct DevTmp {
    T *p = nullptr;
    cudaStream_t st;
    explicit DevTmp(cudaStream_t s) : st(s) {}
    DevTmp(const DevTmp &) = delete;
    DevTmp &operator=(const DevTmp &) = delete;
    ~DevTmp() {
        if (p) cudaFreeAsync(p, st);
    }
    cudaError_t alloc(size_t n) { return cudaMallocAsync((void **)&p, std::max<size_t>(n, 1) * sizeof(T), st); }
};

// The read side (lookups, scans, export, totals of the table) describes the COUNTED table: batches
// that were ingested but not yet counted (inserts are deferred to skm_finalize) would silently be
// missing from the answer, so those calls fail instead.
int32_t check_counted(skm_ctx *c, const char *what) {
    if (c->finalized) return SKM_OK;
    for (auto &cs : c->chunks)
        if (!cs.segs.empty())
            return fail(c, SKM_ERR_STATE, "%s before skm_finalize: ingested batches are not counted yet", what);
    return SKM_OK;
}

}  // namespace

// ============================================================================
// C ABI
// ============================================================================

extern "C" {

uint32_t skm_abi_version(void) { return SKM_ABI_VERSION; }

int32_t skm_create(const skm_params *params, skm_ctx **out) {
    if (!params || !out) return SKM_ERR_INVALID_ARG;
    *out = nullptr;
    skm_ctx *c = new (std::nothrow) skm_ctx();
    if (!c) return SKM_ERR_OOM;
    *out = c;  // returned even on failure so the caller can read skm_last_error, then destroy
    if (params->struct_size != sizeof(skm_params))
        return fail(c, SKM_ERR_INVALID_ARG, "skm_params.struct_size %u != %zu", params->struct_size, sizeof(skm_params));
    c->p = *params;
    // src/cli.rs:662-667
    if (c->p.k < 1 || c->p.k >= 32) return fail(c, SKM_ERR_INVALID_ARG, "k must be less than 32 (and at least 1), got %u", c->p.k);
    if (c->p.k % 2 == 0) return fail(c, SKM_ERR_INVALID_ARG, "k must be odd, got %u", c->p.k);
    // src/cli.rs:668-673
    if (c->p.histo_max < 1 || c->p.histo_max > 1000000)
        return fail(c, SKM_ERR_INVALID_ARG, "histo_max must be between 1 and 1000000, got %llu", (unsigned long long)c->p.histo_max);
    c->n_chunks = c->p.chunks == 0 ? 1 : c->p.chunks;  // src/io.rs:378
    c->n_ranks = c->p.n_ranks == 0 ? 1 : c->p.n_ranks;
    if (c->p.rank >= c->n_ranks) return fail(c, SKM_ERR_INVALID_ARG, "rank %u >= n_ranks %u", c->p.rank, c->n_ranks);
    if (c->n_ranks > kMaxOwners) return fail(c, SKM_ERR_INVALID_ARG, "at most %u ranks, got %u", kMaxOwners, c->n_ranks);
    if (c->p.insert_mode > SKM_INSERT_PARTITIONED) return fail(c, SKM_ERR_INVALID_ARG, "bad insert_mode");

    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(c, SKM_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
    int dev = c->p.device;
    if (dev < 0) CU(cudaGetDevice(&dev));
    if (dev >= n_dev) return fail(c, SKM_ERR_INVALID_ARG, "device %d out of range (%d devices)", dev, n_dev);
    c->device = dev;
    CU(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));
    c->sm_count = prop.multiProcessorCount;
    if (c->p.stream) {
        c->stream = (cudaStream_t)(uintptr_t)c->p.stream;
        c->own_stream = false;
    } else {
        // the inserts outrank the bucketing of later batches (routing / pack streams, default = lowest
        // priority): CTAs of the bucketing kernels only fill what the insert kernel leaves free
        int prio_least = 0, prio_greatest = 0;
        CU(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
        CU(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_greatest));
    }
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->part_stream, cudaStreamNonBlocking));
    {
        int prio_least = 0, prio_greatest = 0;
        CU(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
        CU(cudaStreamCreateWithPriority(&c->pack_stream, cudaStreamNonBlocking, prio_greatest));
    }
    CU(cudaStreamCreateWithFlags(&c->dma_stream, cudaStreamNonBlocking));
    if (c->n_ranks > 1) {
        c->dma_peer.assign(c->n_ranks, nullptr);
        for (auto &ds : c->dma_peer) CU(cudaStreamCreateWithFlags(&ds, cudaStreamNonBlocking));
    }
    CU(cudaStreamCreateWithFlags(&c->sort_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_sort, cudaEventDisableTiming));
    if (const char *g = getenv("SKM_SORT_OVERLAP")) c->sort_overlap = atoi(g) != 0;
    CU(cudaEventCreateWithFlags(&c->ev_dma, cudaEventDisableTiming));
    CU(cudaEventRecord(c->ev_dma, c->dma_stream));
    CU(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    c->work = c->stream;
    if (const char *g = getenv("SKM_INSERT_CTAS")) c->insert_ctas_per_sm = std::max(1, atoi(g));
    // keep freed blocks in the stream-ordered pool (staging buffers are recycled every batch)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    CU(cudaFuncSetAttribute(histogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    // (Asking for the largest shared-memory carve-out on the insert kernels, so that the 108-112 KB
    // scatter CTAs of the next chunk can always move in beside them, DOUBLES the insert time: the
    // insert kernel needs its L1 — profiles/experiments_r01.md #30.  Left at the default.)
    CU(cudaFuncSetAttribute(bucket_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)scatter_smem_bytes(kMaxBuckets)));
    CU(cudaFuncSetAttribute(bucket_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)scatter_smem_bytes(kMaxBuckets, true)));

    CU(cudaMalloc((void **)&c->d_cc, c->n_chunks * sizeof(ChunkCounters)));
    CU(cudaMemset(c->d_cc, 0, c->n_chunks * sizeof(ChunkCounters)));
    CU(cudaMalloc((void **)&c->d_gc, sizeof(GlobalCounters)));
    GlobalCounters gc{};
    gc.first_bad = ~0ull;
    CU(cudaMemcpy(c->d_gc, &gc, sizeof gc, cudaMemcpyHostToDevice));
    CU(cudaMalloc((void **)&c->d_bins, (c->p.histo_max + 2) * sizeof(uint64_t)));
    CU(cudaMalloc((void **)&c->d_hist, (c->p.histo_max + 2) * sizeof(uint64_t)));
    CU(cudaMemset(c->d_hist, 0, (c->p.histo_max + 2) * sizeof(uint64_t)));
    // chunks == 0: the reference keeps no histogram (src/io.rs:1133-1158), so the cheaper RED path is used.
    // SKM_HISTO_SCAN=1 (diagnostic) falls back to one table scan per chunk.
    c->track_histo = c->p.chunks > 0 && !getenv("SKM_HISTO_SCAN");
    if (const char *g = getenv("SKM_PIPE_DEPTH")) c->pipe_depth = atoi(g);
    if (const char *g = getenv("SKM_EAGER")) c->eager = atoi(g) != 0;
    if (const char *g = getenv("SKM_CAPPED")) c->capped = atoi(g) != 0;
    if (const char *g = getenv("SKM_MAX_BUCKETS")) c->max_buckets = std::min<uint32_t>(kMaxBuckets, std::max(1, atoi(g)));
    {
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        c->mem_budget = free_b;
        // SKM_MEM_BUDGET=<MiB>: pretend the device is this small (tests of the memory-bounded paths)
        if (const char *g = getenv("SKM_MEM_BUDGET")) c->mem_budget = std::min<size_t>(free_b, (size_t)atoll(g) << 20);
    }
    CU(cudaEventCreateWithFlags(&c->ev_alloc, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
    if (const char *g = getenv("SKM_REGION_LOG2")) c->region_log2 = std::max(10, std::min(30, atoi(g)));
    if (const char *g = getenv("SKM_L2_FETCH")) {  // diagnostic: 32 / 64 / 128
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
    }
    CU(cudaMalloc((void **)&c->d_tot, sizeof(HistoTotals)));
    c->h_pinned_words = kMaxBuckets + 16;
    CU(cudaMallocHost((void **)&c->h_pinned, c->h_pinned_words * sizeof(uint64_t)));
    CU(cudaMallocHost((void **)&c->h_snap, skm_ctx::kSnapRing * sizeof(uint64_t)));
    // (multi-GPU senders keep one array of kFineRegions + 1 entries per owner in each of the three)
    const size_t scratch_words = std::max<size_t>(kMaxBuckets + 1, (size_t)std::max(c->n_ranks, 1u) * (kFineRegions + 1));
    CU(cudaMalloc((void **)&c->d_bucket_counts, scratch_words * sizeof(uint64_t)));
    CU(cudaMalloc((void **)&c->d_bucket_offsets, scratch_words * sizeof(uint64_t)));
    CU(cudaMalloc((void **)&c->d_bucket_cursors, scratch_words * sizeof(uint64_t)));
    CU(cudaMalloc((void **)&c->d_sort_counters, skm_ctx::kSortCounters * sizeof(unsigned int)));

    c->chunks.resize(c->n_chunks);
    c->histos.resize(c->n_chunks);
    c->col_pending.assign(c->n_chunks, false);
    if ((uint64_t)c->n_chunks * (c->p.histo_max + 2) * sizeof(uint64_t) <= (256ull << 20))
        CU(cudaMallocHost((void **)&c->h_cols, (size_t)c->n_chunks * (c->p.histo_max + 2) * sizeof(uint64_t)));
    c->have_histo.assign(c->n_chunks, false);
    c->col_pending.assign(c->n_chunks, false);
    c->h_cc.assign(c->n_chunks, ChunkCounters{});
    c->chunk_bases_read.assign(c->n_chunks, 0);

    uint32_t l2 = kMinLog2Cap;
    if (c->p.capacity_hint) l2 = std::max(l2, ceil_log2((uint64_t)((double)c->p.capacity_hint / kTargetLoad) + 1));
    if (l2 > 36) return fail(c, SKM_ERR_INVALID_ARG, "capacity_hint too large");
    int32_t rc = alloc_table(c, l2, &c->table);
    if (rc) return rc;
    c->log2cap = l2;
    c->capacity = 1ull << l2;
    // sub-bucket bits of the k-mer lists: with a capacity_hint, exactly as fine as the table's
    // partitions (one sub-bucket per partition); without one, 2^7 sub-buckets per region — a
    // partition then covers several adjacent sub-buckets, or filters a shared one
    {
        // Multi-GPU exchange layout.  Slices need the owner's partitions to be no finer than a sender sorts a coarse
        // bucket, which is known only when a capacity_hint sized the table (the same hint on every rank).  Up to 2^10
        // sub-buckets per coarse bucket the cluster sort pays (C2 at N=8: 54.7 vs 75.2 ms per step); at 2^11 it did not
        // (C3 at N=8: 93.5 ms with slices, 88.6 ms with owner lists), so finer tables (C3, C5) and ctxs without a hint
        // give every owner a re-bucketed list of its own.
        c->mg_slices = c->p.capacity_hint && l2 - kPartLog2 <= route_log2_regions(c) + 10;
        if (const char *g = getenv("SKM_MG_SLICES")) c->mg_slices = atoi(g) != 0;
        // tile size: one CTA sorts 2^13 k-mers; a cluster of 2 / 4 / 8 CTAs sorts 2^14 / 2^15 / 2^16 as one tile.
        // Multi-GPU slices need the large tile: an owner's coarse bucket is sorted by log2(n_ranks) more bits.
        // The tile grows with the number of owners, so that an owner's partition still finds runs of ~64 k-mers.
        c->tile_log2 = kTileLog2;
        if (c->n_ranks > 1 && c->mg_slices) c->tile_log2 = std::min<uint32_t>(kMaxTileLog2, kTileLog2 + ceil_log2(c->n_ranks));
        if (const char *g = getenv("SKM_TILE_LOG2")) c->tile_log2 = (uint32_t)std::max<int>(kTileLog2, std::min<int>(atoi(g), kMaxTileLog2));
        const int g1 = (int)list_log2_regions(c);
        // without a hint: as fine as the partitions of a 2^29-slot table (2^17 of them)
        int g2 = c->p.capacity_hint ? (int)l2 - (int)kPartLog2 - g1 : 17 - g1;
        if (const char *g = getenv("SKM_G2")) g2 = atoi(g);
        c->g2 = (uint32_t)std::max(0, std::min<int>(g2, (int)kMaxSubLog2));
    }
    CU(cudaFuncSetAttribute(tile_sort_cluster_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_sort_cluster_smem_bytes()));
    CU(cudaFuncSetAttribute(tile_sort_cluster_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_sort_cluster_smem_bytes()));
    CU(cudaFuncSetAttribute(tile_sort_cluster_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_sort_cluster_smem_bytes()));
    CU(cudaFuncSetAttribute(tile_sort_cluster_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_sort_cluster_smem_bytes()));
    CU(cudaFuncSetAttribute(tile_sort_cluster_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_sort_cluster_smem_bytes()));
    CU(cudaFuncSetAttribute(tile_sort_cluster_kernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_sort_cluster_smem_bytes()));
    CU(cudaFuncSetAttribute(tile_sort_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)tile_sort_smem_bytes(kMaxSubLog2)));
    CU(cudaFuncSetAttribute(tile_sort_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)tile_sort_smem_bytes(kMaxSubLog2)));
    CU(cudaFuncSetAttribute(tile_sort_owners_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)tile_sort_smem_bytes(kMaxSubLog2)));
    CU(cudaFuncSetAttribute(tile_rebucket_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_rebucket_smem_bytes()));
    CU(cudaFuncSetAttribute(tile_rebucket_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_rebucket_smem_bytes()));
    CU(cudaFuncSetAttribute(tile_rebucket_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_rebucket_smem_bytes()));
    CU(cudaFuncSetAttribute(tile_insert_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileInsertSmemBudget));
    CU(cudaFuncSetAttribute(tile_insert_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileInsertSmemBudget));
    CU(cudaStreamSynchronize(c->stream));
    return SKM_OK;
}

void skm_destroy(skm_ctx *c) {
    if (!c) return;
    if (c->stream) {
        DeviceGuard g(c->device);
        cudaStreamSynchronize(c->stream);
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamSynchronize(c->pack_stream);
        cudaStreamSynchronize(c->part_stream);
        cudaStreamSynchronize(c->sort_stream);
        cudaStreamSynchronize(c->dma_stream);
        for (cudaStream_t ds : c->dma_peer) cudaStreamSynchronize(ds);
        for (auto &cs : c->chunks)
            for (auto &sg : cs.segs) {
                cudaFree(sg.codes);
                cudaFree(sg.breaks);
                cudaFree(sg.list);
                cudaFree(sg.d_counts);
                cudaFree(sg.meta);
                cudaFree(sg.tile_off);
                for (auto &ol : sg.owners) {
                    cudaFree(ol.list);
                    cudaFree(ol.meta);
                    cudaFree(ol.tile_off);
                }
            }
        for (auto &cb : c->buf_cache) {
            cudaFree(cb.p);
            cudaEventDestroy(cb.done);
        }
        cudaFree(c->d_delta);
        cudaFree(c->d_recount);
        cudaFree(c->d_fail);
        cudaFree(c->d_part_ids);
        cudaFree(c->d_segs);
        cudaFreeHost(c->h_segs);
        cudaFree(c->table);
        cudaFree(c->d_cc);
        cudaFree(c->d_gc);
        cudaFree(c->d_bins);
        cudaFree(c->d_hist);
        cudaFree(c->d_tot);
        for (uint32_t r = 0; r < skm_ctx::kMaxPeers; r++)
            if (c->mg_peer_ipc[r]) cudaIpcCloseMemHandle(c->mg_peer[r]);
        cudaFree(c->mg_arena);
        for (uint32_t i = 0; i < skm_ctx::kRawRing; i++) {
            cudaFree(c->raw_buf[i]);
            if (c->raw_copied[i]) cudaEventDestroy(c->raw_copied[i]);
            if (c->raw_packed[i]) cudaEventDestroy(c->raw_packed[i]);
        }
        for (auto b : c->off_blocks) cudaFreeHost(b);
        if (c->ev_alloc) cudaEventDestroy(c->ev_alloc);
        if (c->ev_copy) cudaEventDestroy(c->ev_copy);
        cudaFreeHost(c->h_cols);
        cudaFreeHost(c->h_snap);
        for (auto e : c->snap_event) if (e) cudaEventDestroy(e);
        cudaFree(c->d_bucket_counts);
        cudaFree(c->d_bucket_offsets);
        cudaFree(c->d_bucket_cursors);
        cudaFree(c->d_sort_counters);
        cudaFree(c->d_list);
        cudaFreeHost(c->h_pinned);
        for (auto &t : c->spans) {
            cudaEventDestroy(t.a);
            cudaEventDestroy(t.b);
        }
        for (auto e : c->event_pool) cudaEventDestroy(e);
        for (auto rs : c->read_streams) cudaStreamDestroy(rs);
        if (c->own_stream) cudaStreamDestroy(c->stream);
        cudaStreamDestroy(c->copy_stream);
        cudaStreamDestroy(c->part_stream);
        cudaStreamDestroy(c->pack_stream);
        cudaStreamDestroy(c->dma_stream);
        for (cudaStream_t ds : c->dma_peer) cudaStreamDestroy(ds);
        cudaStreamDestroy(c->sort_stream);
        if (c->ev_sort) cudaEventDestroy(c->ev_sort);
        if (c->ev_dma) cudaEventDestroy(c->ev_dma);
        if (c->ev_main) cudaEventDestroy(c->ev_main);
    }
    delete c;
}

const char *skm_last_error(const skm_ctx *c) { return c ? c->err.c_str() : "null ctx"; }

int32_t skm_pinned_alloc(skm_ctx *c, size_t bytes, void **out) {
    if (!c || !out) return SKM_ERR_INVALID_ARG;
    DeviceGuard g(c->device);
    CU(cudaMallocHost(out, bytes ? bytes : 1));
    return SKM_OK;
}

int32_t skm_pinned_free(skm_ctx *c, void *ptr) {
    if (!c) return SKM_ERR_INVALID_ARG;
    DeviceGuard g(c->device);
    CU(cudaFreeHost(ptr));
    return SKM_OK;
}

int32_t skm_device_alloc(skm_ctx *c, size_t bytes, void **out) {
    if (!c || !out) return SKM_ERR_INVALID_ARG;
    DeviceGuard g(c->device);
    CU(cudaMalloc(out, bytes ? bytes : 1));
    return SKM_OK;
}

int32_t skm_device_free(skm_ctx *c, void *ptr) {
    if (!c) return SKM_ERR_INVALID_ARG;
    DeviceGuard g(c->device);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFree(ptr));
    return SKM_OK;
}

int32_t skm_memcpy_d2h(skm_ctx *c, void *dst, const void *d_src, size_t bytes) {
    if (!c) return SKM_ERR_INVALID_ARG;
    DeviceGuard g(c->device);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return SKM_OK;
}

int32_t skm_memcpy_h2d(skm_ctx *c, void *d_dst, const void *src, size_t bytes) {
    if (!c) return SKM_ERR_INVALID_ARG;
    DeviceGuard g(c->device);
    CU(cudaMemcpy(d_dst, src, bytes, cudaMemcpyHostToDevice));
    return SKM_OK;
}

int32_t skm_ingest_batch(skm_ctx *c, uint32_t chunk, const uint8_t *seqs, uint64_t n_bytes, uint32_t flags) {
    int32_t rc = check_ingest_args(c, chunk, seqs, n_bytes);
    if (rc) return rc;
    if (n_bytes == 0) return SKM_OK;
    if (seqs[n_bytes - 1] != '\n')
        return fail(c, SKM_ERR_INVALID_ARG, "batch must end with a newline-terminated sequence line");
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    // The copy runs on its own stream, into a ring of persistent raw buffers, so that it overlaps
    // the kernels of earlier batches (a buffer is reused once its pack kernel has finished):
    //   copy stream   : wait [packed(b-R)] -> H2D(b) -> [copied(b)] -> H2D(b+1) -> ...
    //   pack stream   : wait [copied(b)] -> pack(b) -> [packed(b)]
    //   routing stream: wait [packed(b)] -> pass A(b) -> wait [packed(b+1)] -> pass A(b+1) ...
    //   sort stream   : pass B(b) beside pass A(b+1)
    //   main stream   : inserts (skm_finalize)
    // The copies have a stream of their own: a pack kernel that is waiting for room on the SMs must
    // not hold back the next batch's copy.  So have the pack kernels (round 1 kept them on the routing
    // stream, beside a persistent insert kernel that left room for one CTA per SM): between pass A of
    // two batches, a pack kernel ran 12x slower beside the tile sort and kept pass A from overlapping it.
    const uint32_t b = c->raw_next++ % skm_ctx::kRawRing;
    if (!c->raw_copied[b]) {
        CU(cudaEventCreateWithFlags(&c->raw_copied[b], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->raw_packed[b], cudaEventDisableTiming));
    }
    if (c->raw_cap[b] < n_bytes) {
        if (c->raw_buf[b]) {
            CU(cudaStreamSynchronize(c->copy_stream));
                    CU(cudaFree(c->raw_buf[b]));
            c->raw_buf[b] = nullptr;
            c->raw_cap[b] = 0;
        }
        const size_t cap = std::max<size_t>(n_bytes, 1 << 20);
        CU(cudaMalloc((void **)&c->raw_buf[b], cap));
        c->raw_cap[b] = cap;
    }
    if (c->raw_in_use[b]) CU(cudaStreamWaitEvent(c->copy_stream, c->raw_packed[b], 0));  // buffer free again
    {
        Span sp(c, ST_H2D, c->copy_stream);
        CU(cudaMemcpyAsync(c->raw_buf[b], seqs, n_bytes, cudaMemcpyHostToDevice, c->copy_stream));
    }
    CU(cudaEventRecord(c->raw_copied[b], c->copy_stream));
    if (!(flags & SKM_INGEST_ASYNC)) CU(cudaEventSynchronize(c->raw_copied[b]));
    CU(cudaStreamWaitEvent(c->pack_stream, c->raw_copied[b], 0));
    WorkStream ws(c, c->part_stream);
    c->raw_in_use[b] = true;
    cudaEvent_t copied = get_event(c);
    CU(cudaEventRecord(copied, c->copy_stream));
    const size_t n_before = c->chunks[chunk].segs.size();
    rc = stage_device(c, chunk, c->raw_buf[b], n_bytes, c->pack_stream, c->raw_packed[b]);
    if (c->chunks[chunk].segs.size() > n_before) c->chunks[chunk].segs.back().copied = copied;
    else c->event_pool.push_back(copied);
    return rc;
}

int32_t skm_ingest_reads(skm_ctx *c, uint32_t chunk, const uint8_t *bases, const uint64_t *offsets,
                         uint64_t n_reads) {
    int32_t rc = check_ingest_args(c, chunk, offsets, n_reads);
    if (rc) return rc;
    if (n_reads == 0) return SKM_OK;
    const uint64_t n_bases = offsets[n_reads] - offsets[0];
    if (n_bases && !bases) return fail(c, SKM_ERR_INVALID_ARG, "null buffer");
    for (uint64_t i = 0; i < n_reads; i++)
        if (offsets[i + 1] < offsets[i]) return fail(c, SKM_ERR_INVALID_ARG, "offsets must be non-decreasing");
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    // Everything runs on the routing stream, like skm_ingest_batch / skm_ingest_device: the bucketing
    // scratch (d_bucket_counts / offsets / cursors) belongs to that stream, so mixing the three entry
    // points in one run cannot race on it.
    cudaStream_t st = c->part_stream;
    WorkStream ws(c, st);
    uint8_t *d_bases = nullptr, *d_lines = nullptr;
    uint64_t *d_off = nullptr;
    const uint64_t n_out = offsets[n_reads] + n_reads;  // dst offset of read r = offsets[r] + r
    CU(cudaMallocAsync((void **)&d_bases, offsets[n_reads] + 1, st));
    CU(cudaMallocAsync((void **)&d_off, (n_reads + 1) * sizeof(uint64_t), st));
    CU(cudaMallocAsync((void **)&d_lines, n_out, st));
    {
        Span sp(c, ST_H2D, st);
        if (offsets[n_reads])
            CU(cudaMemcpyAsync(d_bases, bases, offsets[n_reads], cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_off, offsets, (n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    }
    CU(cudaStreamSynchronize(st));  // the caller's (pageable) buffers may be reused on return
    add_separators_kernel<<<grid_for(n_reads * 32, 256), 256, 0, st>>>(d_bases, d_off, n_reads, d_lines);
    c->launches++;
    CU(cudaGetLastError());
    // bytes before offsets[0] are not part of any read: skip them
    rc = stage_device(c, chunk, d_lines + offsets[0], n_out - offsets[0]);
    CU(cudaFreeAsync(d_bases, st));
    CU(cudaFreeAsync(d_off, st));
    CU(cudaFreeAsync(d_lines, st));
    return rc;
}

int32_t skm_ingest_device(skm_ctx *c, uint32_t chunk, const uint8_t *d_seqs, uint64_t n_bytes) {
    int32_t rc = check_ingest_args(c, chunk, d_seqs, n_bytes);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    // d_seqs was produced on the ctx's main stream (or is already complete): order the pack after it
    CU(cudaEventRecord(c->ev_main, c->stream));
    CU(cudaStreamWaitEvent(c->pack_stream, c->ev_main, 0));
    WorkStream ws(c, c->part_stream);
    return stage_device(c, chunk, d_seqs, n_bytes, c->pack_stream);
}

int32_t skm_sync(skm_ctx *c) {
    if (!c) return SKM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    int32_t rc = sync_all(c);
    if (rc) return rc;
    collect_spans(c);
    return check_sticky(c);
}

static int32_t finalize_common(skm_ctx *c) {
    if (c->finalized) return fail(c, SKM_ERR_STATE, "finalize called twice");
    if (c->n_ranks > 1) return fail(c, SKM_ERR_STATE, "a ctx with n_ranks > 1 is finalized with skm_mg_finalize (collective)");
    bool any = false;
    for (auto &cs : c->chunks) any = any || !cs.segs.empty();
    int32_t rc;
    if (!any) {
        rc = sync_all(c);
        if (rc) return rc;
        rc = check_sticky(c);
        if (rc) return rc;
        // src/io.rs:578-580 (every non-empty batch holds at least one read)
        return fail(c, SKM_ERR_NO_READS, "No reads were ingested. Check that input files contain valid FASTQ records.");
    }
    c->finalized = true;
    // The chunk loop starts while later batches may still be packing / bucketing on part_stream:
    // the inserts of chunk c (main stream) overlap the bucketing of chunks > c.  An invalid base
    // found by a late pack kernel is reported at the end (the run is aborted either way).

    cudaEvent_t e0 = get_event(c), e1 = get_event(c);
    cudaEventRecord(e0, c->stream);
    c->recount_valid = false;
    // A fresh table is sized once, for the whole input: the tiled insert counts all chunks in one
    // launch and cannot grow in the middle of it without a retry.  With a capacity_hint the table
    // already has its size; without one, the number of ingested positions bounds the distinct k-mers.
    if (c->table_fresh && c->n_ranks == 1) {
        uint64_t positions = 0;
        for (auto &cs : c->chunks) positions += cs.n_bytes;
        uint64_t bound = positions;
        if (c->p.k < 31) bound = std::min<uint64_t>(bound, (1ull << (2 * c->p.k)) / 2 + (1ull << c->p.k));  // canonical k-mers
        if (want_partitioned(c, positions) || c->capacity * sizeof(Slot) < (96ull << 20)) {
            uint32_t want = c->log2cap;
            if (!c->p.capacity_hint) want = std::max(want, ceil_log2((uint64_t)((double)bound / kTargetLoad) + 1));
            size_t free_b = 0, total_b = 0;
            CU(cudaMemGetInfo(&free_b, &total_b));
            while (want > c->log2cap && ((size_t)sizeof(Slot) << want) > (free_b + c->capacity * sizeof(Slot)) / 2) want--;
            if (want > c->log2cap && want_partitioned_cap(c, positions, 1ull << want)) {
                rc = grow_table(c, want);
                if (rc) return rc;
            }
        }
    }
    std::vector<SegDesc> group;       // lists of consecutive chunks, to be counted by one tiled launch
    std::vector<Segment *> group_segs;
    uint32_t group_chunk0 = 0;
    auto flush_group = [&](uint32_t end_chunk) -> int32_t {
        if (group.empty()) {
            group_chunk0 = end_chunk;
            return SKM_OK;
        }
        const uint32_t n_l = end_chunk - group_chunk0;
        int32_t r = launch_tiled(c, group, group_chunk0, n_l);
        if (r) return r;
        for (Segment *psg : group_segs) {
            Segment &sg = *psg;
            uint64_t nk = 0;
            if (sg.cap) for (uint32_t b = 0; b < sg.n_buckets; b++) nk += sg.h_offsets[b];
            else nk = sg.h_offsets[sg.n_buckets];
            c->insert_kmers += nk;
            release_list(c, sg, c->stream);
        }
        group.clear();
        group_segs.clear();
        group_chunk0 = end_chunk;
        return SKM_OK;
    };
    const uint32_t pbits_now = c->log2cap - kPartLog2, g1_now = route_log2_regions(c);
    const uint64_t nbr_now = pbits_now >= g1_now ? 1ull : (1ull << (g1_now - pbits_now));
    uint64_t positions_all = 0, group_positions = 0, positions_seen = 0;
    for (auto &cs : c->chunks) positions_all += cs.n_bytes;
    const bool early_flush = !getenv("SKM_NO_EARLY_FLUSH");

    // SPECULATIVE LAUNCH.  When the whole input is already on the device and every batch was bucketed at ingest time
    // with the capped layout, the insert is queued behind the lists' events NOW — before the host has waited for
    // any of them — so that a host that is late (a descheduled thread: about one step in ten stalled for 40-70 ms
    // between the last tile sort and the launch, profiles/experiments_r02.md #14) does not leave the GPU idle.
    // The one thing the host would have checked first, an overflowed capped list, is checked on the device
    // (spec_guard_kernel): then the launch does nothing and the loop below takes over.
    bool spec_done = false;
    if (c->table_fresh && c->n_ranks == 1 && c->n_chunks <= kMaxChunksPerLaunch && !getenv("SKM_NO_SPEC")) {
        std::vector<SegDesc> all;
        GuardWords gw{};
        uint32_t n_words = 0;
        bool ok = true;
        for (uint32_t ch = 0; ch < c->n_chunks && ok; ch++)
            for (auto &sg : c->chunks[ch].segs) {
                const bool arriving = sg.copied && cudaEventQuery(sg.copied) == cudaErrorNotReady;
                if (!sg.ready || !sg.list || !sg.tiled || !sg.cap || arriving || n_words >= kMaxGuardWords) {
                    ok = false;
                    break;
                }
                all.push_back(local_desc(c, sg, ch));
                gw.w[n_words++] = (const unsigned long long *)(sg.h_offsets + sg.n_buckets);
            }
        cudaGetLastError();
        if (ok && !all.empty() && all.size() * nbr_now <= kMaxVseg) {
            for (auto &cs : c->chunks)
                for (auto &sg : cs.segs) CU(cudaStreamWaitEvent(c->stream, sg.ready, 0));
            unsigned long long *flag = &c->d_gc->pad;
            spec_guard_kernel<<<1, 32, 0, c->stream>>>(gw, n_words, flag);
            c->launches++;
            rc = launch_tiled(c, all, 0, c->n_chunks, flag);
            if (rc > 0) return rc;
            if (rc == SKM_OK) {
                for (auto &cs : c->chunks) {
                    for (auto &sg : cs.segs) {
                        for (uint32_t b = 0; b < sg.n_buckets; b++) c->insert_kmers += sg.h_offsets[b];
                        if (sg.codes) CU(cudaFreeAsync(sg.codes, c->stream));
                        if (sg.breaks) CU(cudaFreeAsync(sg.breaks, c->stream));
                        sg.codes = nullptr;
                        sg.breaks = nullptr;
                        release_list(c, sg, c->stream);
                    }
                    cs.counted = true;
                }
                spec_done = true;
            }
            // (rc == kSpecAborted: some list overflowed; nothing was counted)
        }
    }
    for (uint32_t ch = 0; ch < c->n_chunks && !spec_done; ch++) {
        ChunkState &cs = c->chunks[ch];
        // Host-fed input that is still on its way over PCIe: rather than idle until the last batch has arrived,
        // count what is listed so far (one more pass over the table, hidden behind the copies).  Only while a
        // good part of the input is still to come: a launch just before the last batch lands costs a table
        // pass and holds the SMs when that batch wants to be bucketed (measured: 3 + 6 + 1 chunks 50.7 ms,
        // 3 + 7 chunks 48.0 ms end to end on C2).
        if (early_flush && !group.empty() && group_positions >= std::max<uint64_t>(64ull << 20, positions_all / 4) &&
            positions_all - positions_seen >= positions_all / 4) {
            bool arriving = false;
            for (auto &sg : cs.segs)
                if (sg.copied && cudaEventQuery(sg.copied) == cudaErrorNotReady) arriving = true;
            cudaGetLastError();
            if (arriving) {
                rc = flush_group(ch);
                if (rc) return rc;
                group_positions = 0;
            }
        }
        group_positions += cs.n_bytes;
        positions_seen += cs.n_bytes;
        for (auto &sg : cs.segs) {
            if (!sg.ready) continue;
            if (sg.list) CU(cudaEventSynchronize(sg.ready));      // host needs the bucket totals
            else CU(cudaStreamWaitEvent(c->stream, sg.ready, 0));  // packed data is consumed on the main stream
        }
        // (1) lists built at ingest time; a capped list that overflowed (heavily repeated k-mers) is
        //     rebuilt with the exact two-pass layout from the packed form, which was kept for this
        for (size_t si = 0; si < cs.segs.size(); si++) {
            Segment &sg = cs.segs[si];
            if (!sg.list || !sg.tiled) continue;
            if (sg.cap && sg.h_offsets[sg.n_buckets] > 0) {
                WorkStream ws(c, c->part_stream);
                rc = drop_capped_list(c, ch, sg, c->part_stream);
                if (rc) return rc;
                c->n_capped_fallbacks++;
                uint64_t *h_off = alloc_offsets(c, sg.n_buckets + 1);
                if (!h_off) return fail(c, SKM_ERR_OOM, "pinned allocation failed");
                rc = build_list(c, ch, si, h_off, /*exact=*/true, /*must=*/true);
                if (rc) return rc;
                CU(cudaEventSynchronize(sg.ready));
            } else if (sg.codes) {  // the packed form is no longer needed
                CU(cudaFreeAsync(sg.codes, c->stream));
                CU(cudaFreeAsync(sg.breaks, c->stream));
                sg.codes = nullptr;
                sg.breaks = nullptr;
            }
        }
        // (2) this chunk's lists join the group; segments still packed get their lists built now when the
        //     chunk is worth it.  When memory is short (inputs larger than the GPU: BASELINE config 5) or a
        //     launch cannot read more lists, what is grouped so far is counted first — if need be in the
        //     middle of a chunk: the order inside a chunk does not matter, and every launch that touches a
        //     chunk rewrites its column, so the last one leaves it complete.
        if (ch - group_chunk0 >= kMaxChunksPerLaunch) {
            rc = flush_group(ch);
            if (rc) return rc;
        }
        uint64_t packed_bytes = 0;
        for (auto &sg : cs.segs)
            if (sg.codes && !sg.list) packed_bytes += sg.n_bytes;
        const bool build_now = packed_bytes && c->n_ranks == 1 && want_partitioned(c, packed_bytes);
        for (size_t si = 0; si < cs.segs.size(); si++) {
            Segment &sg = cs.segs[si];
            if (!sg.list && sg.codes && build_now) {
                WorkStream ws(c, c->part_stream);
                for (int attempt = 0; attempt < 2 && !sg.list; attempt++) {
                    uint64_t *h_off = alloc_offsets(c, (c->n_ranks << route_log2_regions(c)) + 1);
                    if (!h_off) return fail(c, SKM_ERR_OOM, "pinned allocation failed");
                    rc = build_list(c, ch, si, h_off, /*exact=*/true, /*must=*/false, /*reserve_tables=*/1);
                    if (rc) return rc;
                    if (!sg.list && attempt == 0) {  // memory is short: count what is listed so far, then retry
                        if (group.empty()) break;
                        rc = flush_group(ch + 1);
                        if (rc) return rc;
                        group_chunk0 = ch;
                    }
                }
                if (sg.list) CU(cudaEventSynchronize(sg.ready));
            }
            if (!(sg.list && sg.tiled)) continue;
            if ((group.size() + 1) * nbr_now > kMaxVseg) {  // more lists than one launch reads
                rc = flush_group(ch + 1);
                if (rc) return rc;
                group_chunk0 = ch;
            }
            group.push_back(local_desc(c, sg, ch - group_chunk0));
            group_segs.push_back(&sg);
        }
        // (4) what is still packed goes through the direct kernel (small inputs, or no memory for a list);
        //     a chunk's lists are counted first, so that the chunk's column is complete afterwards
        bool direct = false;
        for (auto &sg : cs.segs) direct = direct || (sg.codes && !sg.list);
        if (direct) {
            rc = flush_group(ch + 1);
            if (rc) return rc;
            for (auto &sg : cs.segs) {
                if (!sg.codes || sg.list) continue;
                rc = insert_segment_direct(c, sg, ch);
                if (rc) return rc;
            }
            if (c->p.chunks > 0) {
                rc = snapshot_histogram(c, ch);
                if (rc) return rc;
            }
        } else if (group.empty() && c->p.chunks > 0) {
            // an empty chunk between launches: its column is the running histogram
            rc = flush_group(ch);
            if (rc) return rc;
            rc = snapshot_histogram(c, ch);
            if (rc) return rc;
            group_chunk0 = ch + 1;
        }
        for (auto &sg : cs.segs) {  // drop(chunk), src/io.rs:1025 (lists are released after their launch)
            if (sg.list && sg.tiled) continue;
            if (sg.codes) CU(cudaFreeAsync(sg.codes, c->stream));
            if (sg.breaks) CU(cudaFreeAsync(sg.breaks, c->stream));
            sg.codes = nullptr;
            sg.breaks = nullptr;
            release_list(c, sg, c->stream);
        }
        cs.counted = true;
    }
    rc = flush_group(c->n_chunks);
    if (rc) return rc;
    for (auto &cs : c->chunks)
        for (auto &sg : cs.segs) {
            if (sg.ready) c->event_pool.push_back(sg.ready);
            if (sg.copied) c->event_pool.push_back(sg.copied);
            sg.ready = nullptr;
            sg.copied = nullptr;
        }
    // Totals for the conservation checks and an independent recount of the final histogram: the
    // tiled insert histograms every partition as it writes it back; otherwise one scan of the table.
    std::vector<uint64_t> rescan;
    if (c->recount_valid) {
        rescan.assign(c->p.histo_max + 2, 0);
        CU(cudaMemcpyAsync(rescan.data(), c->d_recount, rescan.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(&c->last_tot, c->d_tot, sizeof(HistoTotals), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->have_tot = true;
    } else {
        rc = scan_table(c, false, &rescan);
        if (rc) return rc;
    }
    cudaEventRecord(e1, c->stream);
    rc = sync_all(c);
    if (rc) return rc;
    materialize_cols(c);
    rc = check_sticky(c);
    if (rc) return rc;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    c->stage_ms[ST_FINALIZE] += ms;
    c->event_pool.push_back(e0);
    c->event_pool.push_back(e1);
    collect_spans(c);
    rc = refresh_chunk_counters(c);
    if (rc) return rc;

    // conservation identities (src/io.rs:1042-1047, 1120-1132)
    uint64_t n_windows = 0;
    for (auto &cc : c->h_cc) n_windows += cc.n_windows;
    uint64_t d = 0;
    rc = read_distinct(c, &d);
    if (rc) return rc;
    c->distinct_ub = d;
    if (c->last_tot.n_saturated == 0 && c->last_tot.n_kmers != n_windows)
        return fail(c, SKM_ERR_CONSERVATION,
                    "The total count of hashed kmers (%llu) does not equal the number of ingested kmers (%llu)",
                    (unsigned long long)c->last_tot.n_kmers, (unsigned long long)n_windows);
    if (c->last_tot.n_distinct != d)
        return fail(c, SKM_ERR_CONSERVATION,
                    "The total count of unique kmers in the histogram (%llu) does not equal the total count of hashed kmers (%llu)",
                    (unsigned long long)c->last_tot.n_distinct, (unsigned long long)d);
    if (c->p.chunks > 0) {
        const std::vector<uint64_t> &h = c->histos[c->n_chunks - 1];
        if (h != rescan)
            return fail(c, SKM_ERR_CONSERVATION, "The incremental histogram does not match a recount of the table");
        uint64_t uniq = 0;
        for (size_t i = 1; i < h.size(); i++) uniq += h[i];
        if (uniq != d)
            return fail(c, SKM_ERR_CONSERVATION,
                        "The total count of unique kmers in the histogram (%llu) does not equal the total count of hashed kmers (%llu)",
                        (unsigned long long)uniq, (unsigned long long)d);
    }
    return SKM_OK;
}

int32_t skm_finalize(skm_ctx *c) {
    if (!c) return SKM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    return finalize_common(c);
}

int32_t skm_reset(skm_ctx *c) {
    if (!c) return SKM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    {
        int32_t rc0 = sync_all(c);
        if (rc0) return rc0;
    }
    for (auto &cs : c->chunks) {
        for (auto &sg : cs.segs) {
            if (sg.ready) c->event_pool.push_back(sg.ready);
            if (sg.copied) c->event_pool.push_back(sg.copied);
            if (sg.codes) CU(cudaFreeAsync(sg.codes, c->stream));
            if (sg.breaks) CU(cudaFreeAsync(sg.breaks, c->stream));
            release_list(c, sg, c->stream);
        }
        cs = ChunkState{};
    }
    for (auto &cur : c->mg_cursor) cur = 0;
    c->mg_sent.clear();
    c->mg_bytes_sent = 0;
    // the clear is deferred (ensure_physical): a tiled insert never needs it
    c->table_fresh = true;
    c->table_zombie = true;
    c->recount_valid = false;
    c->n_tiled_launches = c->n_tiled_retries = 0;
    CU(cudaMemsetAsync(c->d_cc, 0, c->n_chunks * sizeof(ChunkCounters), c->stream));
    CU(cudaMemsetAsync(c->d_hist, 0, (c->p.histo_max + 2) * sizeof(uint64_t), c->stream));
    GlobalCounters gc{};
    gc.first_bad = ~0ull;
    CU(cudaMemcpyAsync(c->d_gc, &gc, sizeof gc, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    collect_spans(c);
    for (int i = 0; i < ST_N; i++) {
        c->stage_ms[i] = 0;
        c->stage_launches[i] = 0;
    }
    c->insert_kmers = c->insert_bases = 0;
    while (c->off_blocks.size() > 1) {
        cudaFreeHost(c->off_blocks.back());
        c->off_blocks.pop_back();
    }
    c->off_used = 0;
    c->list_bytes = 0;
    c->launches = 0;
    c->n_grows = 0;
    c->distinct_ub = 0;
    c->launched_total = 0;
    for (auto &p : c->snap_pending) p = false;
    c->pos_base = 0;
    c->have_tot = false;
    c->finalized = false;
    c->sticky_error = false;
    c->have_histo.assign(c->n_chunks, false);
    c->col_pending.assign(c->n_chunks, false);
    c->h_cc.assign(c->n_chunks, ChunkCounters{});
    c->err.clear();
    return SKM_OK;
}

int32_t skm_histogram(skm_ctx *c, uint32_t chunk_i, uint64_t *out, uint64_t out_len) {
    if (!c || !out) return SKM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if (chunk_i < c->n_chunks && c->col_pending[chunk_i]) {
        DeviceGuard g(c->device);
        CU(cudaStreamSynchronize(c->stream));
        materialize_cols(c);
    }
    if (chunk_i >= c->n_chunks || !c->have_histo[chunk_i])
        return fail(c, SKM_ERR_STATE, "no histogram snapshot for chunk %u (chunks=%u, finalized=%d)", chunk_i, c->p.chunks, (int)c->finalized);
    if (out_len < c->p.histo_max + 2) return fail(c, SKM_ERR_INVALID_ARG, "histogram buffer too small");
    memcpy(out, c->histos[chunk_i].data(), (c->p.histo_max + 2) * sizeof(uint64_t));
    return SKM_OK;
}

static int32_t refresh_totals(skm_ctx *c) {
    // table-wide totals come from a histogram pass; refresh if the table changed since
    int32_t rc = SKM_OK;
    if (!c->have_tot) rc = scan_table(c, false, nullptr);
    return rc;
}

int32_t skm_totals_get(skm_ctx *c, skm_totals *out) {
    if (!c || !out) return SKM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    int32_t rc = sync_all(c);
    if (rc) return rc;
    materialize_cols(c);
    rc = refresh_chunk_counters(c);
    if (rc) return rc;
    rc = refresh_totals(c);
    if (rc) return rc;
    memset(out, 0, sizeof *out);
    for (uint32_t i = 0; i < c->n_chunks; i++) {
        out->n_reads += c->h_cc[i].n_reads;
        out->n_bases += c->h_cc[i].n_bases;
        out->n_bases_read += c->chunks[i].n_bytes - c->h_cc[i].n_reads;
    }
    out->n_kmers = c->last_tot.n_kmers;
    out->n_unique = c->last_tot.n_distinct;
    out->n_saturated = c->last_tot.n_saturated;
    if (c->p.chunks > 0 && c->have_histo[c->n_chunks - 1]) out->n_singletons = c->histos[c->n_chunks - 1][1];
    return SKM_OK;
}

int32_t skm_chunk_totals(skm_ctx *c, uint32_t chunk, skm_totals *out) {
    if (!c || !out) return SKM_ERR_INVALID_ARG;
    if (chunk >= c->n_chunks) return fail(c, SKM_ERR_INVALID_ARG, "chunk_index out of range");
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    int32_t rc = sync_all(c);
    if (rc) return rc;
    rc = refresh_chunk_counters(c);
    if (rc) return rc;
    memset(out, 0, sizeof *out);
    out->n_reads = c->h_cc[chunk].n_reads;
    out->n_bases = c->h_cc[chunk].n_bases;
    out->n_bases_read = c->chunks[chunk].n_bytes - c->h_cc[chunk].n_reads;
    out->n_kmers = c->h_cc[chunk].n_windows;  // Chunk::get_n_kmers: occurrences counted for this chunk
    return SKM_OK;
}

int32_t skm_stage_times(skm_ctx *c, skm_stage_ms *out) {
    if (!c || !out) return SKM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    int32_t rc0 = sync_all(c);
    if (rc0) return rc0;
    collect_spans(c);
    out->h2d = c->stage_ms[ST_H2D];
    out->pack = c->stage_ms[ST_PACK];
    out->count = c->stage_ms[ST_COUNT];
    out->partition = c->stage_ms[ST_PART];
    out->insert = c->stage_ms[ST_INSERT];
    out->histogram = c->stage_ms[ST_HISTO];
    out->grow = c->stage_ms[ST_GROW];
    out->total_finalize = c->stage_ms[ST_FINALIZE];
    for (int i = 0; i < 8; i++) out->launches[i] = c->stage_launches[i];
    out->insert_kmers = c->insert_kmers;
    out->insert_bases = c->insert_bases;
    out->kernel_launches = c->launches;
    out->n_grows = c->n_grows;
    out->table_capacity = c->capacity;
    out->table_bytes = c->capacity * sizeof(Slot);
    out->sort = c->stage_ms[ST_SORT];
    out->scan = c->stage_ms[ST_SCAN];
    out->sort_launches = c->stage_launches[ST_SORT];
    out->scan_launches = c->stage_launches[ST_SCAN];
    out->tiled_launches = c->n_tiled_launches;
    out->tiled_retries = c->n_tiled_retries;
    return SKM_OK;
}

int32_t skm_table_len(skm_ctx *c, uint64_t *out) {
    if (!c || !out) return SKM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    return read_distinct(c, out);
}

int32_t skm_export(skm_ctx *c, uint64_t *keys, uint32_t *counts, uint64_t cap, int32_t sorted, uint64_t *n_out) {
    if (!c || !n_out) return SKM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    int32_t rc = check_counted(c, "skm_export");
    if (rc) return rc;
    rc = ensure_physical(c);
    if (rc) return rc;
    uint64_t n = 0;
    rc = read_distinct(c, &n);
    if (rc) return rc;
    *n_out = n;
    if (!keys && !counts) return SKM_OK;  // size query
    if (!keys || !counts || cap < n)
        return fail(c, SKM_ERR_INVALID_ARG, "export buffers too small: need %llu", (unsigned long long)n);
    if (n == 0) return SKM_OK;
    DevTmp<unsigned long long> d_keys(c->stream), d_cursor(c->stream);
    DevTmp<uint32_t> d_counts(c->stream);
    CU(d_keys.alloc(n));
    CU(d_counts.alloc(n));
    CU(d_cursor.alloc(1));
    CU(cudaMemsetAsync(d_cursor.p, 0, sizeof(uint64_t), c->stream));
    export_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->table, c->capacity, d_keys.p, d_counts.p, n, d_cursor.p);
    c->launches++;
    CU(cudaGetLastError());
    if (sorted) {
        // presentation order (the table itself has none): ascending k-mer, sorted on the device
        rc = sort_pairs_device(c, d_keys.p, d_counts.p, n);
        if (rc) return rc;
    }
    CU(cudaMemcpyAsync(keys, d_keys.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(counts, d_counts.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->h_pinned, d_cursor.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->h_pinned[0] != n)
        return fail(c, SKM_ERR_CONSERVATION, "export found %llu occupied slots, expected %llu",
                    (unsigned long long)c->h_pinned[0], (unsigned long long)n);
    return SKM_OK;
}

int32_t skm_table_digest(skm_ctx *c, uint64_t *out) {
    if (!c || !out) return SKM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    int32_t rc = check_counted(c, "skm_table_digest");
    if (rc) return rc;
    rc = scan_table(c, true, nullptr);
    if (rc) return rc;
    *out = c->last_tot.digest;
    return SKM_OK;
}

int32_t skm_lookup_batch(skm_ctx *c, const uint64_t *kmers, uint64_t n, uint32_t min_count, int32_t mode,
                         uint32_t *counts, uint8_t *found) {
    if (!c || (n && !kmers) || mode < 0 || mode > SKM_LOOKUP_EITHER) return SKM_ERR_INVALID_ARG;
    if (n == 0) return SKM_OK;
    // Readers share the table like the reference's rayon workers do (src/stats.rs:84-98): the ctx
    // lock is held only to take a private stream, never across the kernel or the copies.
    cudaStream_t st = nullptr;
    TableRef tr;
    uint32_t k;
    int sm_count;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        int32_t rc = check_counted(c, "skm_lookup_batch");
        if (rc) return rc;
        DeviceGuard g(c->device);
        rc = ensure_physical(c);
        if (rc) return rc;
        CU(cudaStreamSynchronize(c->stream));  // inserts queued before this call are visible
        st = take_read_stream(c);
        tr = tref(c);
        k = c->p.k;
        sm_count = c->sm_count;
        c->launches++;
    }
    DeviceGuard g(c->device);
    int32_t rc = SKM_OK;
    {
        DevTmp<unsigned long long> d_q(st);
        DevTmp<uint32_t> d_c(st);
        DevTmp<uint8_t> d_f(st);
        cudaError_t e = d_q.alloc(n);
        if (e == cudaSuccess) e = d_c.alloc(n);
        if (e == cudaSuccess) e = d_f.alloc(n);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_q.p, kmers, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            lookup_kernel<<<std::min<uint32_t>(grid_for(n, 256), sm_count * 8), 256, 0, st>>>(tr, k, d_q.p, n, min_count,
                                                                                             mode, d_c.p, d_f.p);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess && counts) e = cudaMemcpyAsync(counts, d_c.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && found) e = cudaMemcpyAsync(found, d_f.p, n, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            std::lock_guard<std::mutex> lk(c->mu);
            rc = fail(c, e == cudaErrorMemoryAllocation ? SKM_ERR_OOM : SKM_ERR_CUDA, "skm_lookup_batch: %s",
                      cudaGetErrorString(e));
        }
    }
    std::lock_guard<std::mutex> lk(c->mu);
    c->read_streams.push_back(st);
    return rc;
}

int32_t skm_scan_oligos(skm_ctx *c, const uint64_t *oligos, uint64_t n_oligos, uint32_t oligo_length,
                        uint32_t min_count, uint64_t *keys, uint32_t *counts, uint64_t cap, uint64_t *n_out) {
    if (!c || !n_out || (n_oligos && !oligos)) return SKM_ERR_INVALID_ARG;
    *n_out = 0;
    const uint32_t k = c->p.k;
    // src/pcr/primers.rs:168-187
    if (n_oligos == 0) return fail(c, SKM_ERR_INVALID_ARG, "find_oligos_in_kmers called with no oligos");
    if (oligo_length == 0 || oligo_length >= k)
        return fail(c, SKM_ERR_INVALID_ARG, "oligo length %u out of range for k=%u (must be 1..k-1); trim must be < k",
                    oligo_length, k);
    if (n_oligos > (1u << 24)) return fail(c, SKM_ERR_INVALID_ARG, "too many oligos");
    std::vector<uint64_t> sets(2 * n_oligos);  // [sorted forward prefixes | sorted reverse-complement suffixes]
    for (uint64_t i = 0; i < n_oligos; i++) {
        sets[i] = oligos[i] << (2 * (k - oligo_length));
        sets[n_oligos + i] = skm_revcomp_kmer(oligos[i], oligo_length);
    }
    std::sort(sets.begin(), sets.begin() + n_oligos);
    std::sort(sets.begin() + n_oligos, sets.end());
    const unsigned long long mask = ((1ull << (2 * oligo_length)) - 1) << (2 * k - 2 * oligo_length);
    const unsigned long long rc_mask = (1ull << (2 * oligo_length)) - 1;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    int32_t rc_code = check_counted(c, "skm_scan_oligos");
    if (rc_code) return rc_code;
    rc_code = ensure_physical(c);
    if (rc_code) return rc_code;
    DevTmp<unsigned long long> d_sets(c->stream), d_keys(c->stream), d_cursor(c->stream);
    DevTmp<uint32_t> d_counts(c->stream);
    const uint64_t out_cap = (keys && counts) ? cap : 0;
    CU(d_sets.alloc(2 * n_oligos));
    CU(d_keys.alloc(out_cap));
    CU(d_counts.alloc(out_cap));
    CU(d_cursor.alloc(1));
    CU(cudaMemcpyAsync(d_sets.p, sets.data(), 2 * n_oligos * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(d_cursor.p, 0, sizeof(uint64_t), c->stream));
    CU(cudaStreamSynchronize(c->stream));  // `sets` is a pageable host vector
    {
        Span sp(c, ST_SCAN, c->stream);
        scan_oligos_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->table, c->capacity, k, d_sets.p, d_sets.p + n_oligos,
                                                                   (uint32_t)n_oligos, mask, rc_mask, min_count, d_keys.p,
                                                                   d_counts.p, out_cap, d_cursor.p, oligo_length);
        c->launches++;
        c->stage_launches[ST_SCAN]++;
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_pinned, d_cursor.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const uint64_t n = c->h_pinned[0];
    *n_out = n;
    if (n == 0 || (!keys && !counts)) return SKM_OK;  // nothing found, or a size query
    if (!keys || !counts || cap < n)
        return fail(c, SKM_ERR_INVALID_ARG, "scan buffers too small: need %llu", (unsigned long long)n);
    // presentation order: ascending k-mer, like skm_export(sorted)
    rc_code = sort_pairs_device(c, d_keys.p, d_counts.p, n);
    if (rc_code) return rc_code;
    CU(cudaMemcpyAsync(keys, d_keys.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(counts, d_counts.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SKM_OK;
}

int32_t skm_insert_counts(skm_ctx *c, const uint64_t *keys, const uint32_t *counts, uint64_t n) {
    if (!c || (n && (!keys || !counts))) return SKM_ERR_INVALID_ARG;
    if (n == 0) return SKM_OK;
    for (uint64_t i = 0; i < n; i++) {
        if (keys[i] == SKM_EMPTY_KEY) return fail(c, SKM_ERR_INVALID_ARG, "key %llu is the EMPTY sentinel (k <= 31 keys never are)", (unsigned long long)i);
        // a zero count would claim a slot without histogram mass (KmerCounts never holds one:
        // src/kmer/counting.rs:82-85 inserts 1 or adds to an existing count)
        if (counts[i] == 0) return fail(c, SKM_ERR_INVALID_ARG, "count %llu is zero", (unsigned long long)i);
    }
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    unsigned long long *d_k = nullptr;
    uint32_t *d_c = nullptr;
    CU(cudaMallocAsync((void **)&d_k, n * sizeof(uint64_t), c->stream));
    CU(cudaMallocAsync((void **)&d_c, n * sizeof(uint32_t), c->stream));
    CU(cudaMemcpyAsync(d_k, keys, n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_c, counts, n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    int32_t rc = insert_list(c, d_k, d_c, n);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFreeAsync(d_k, c->stream));
    CU(cudaFreeAsync(d_c, c->stream));
    c->have_tot = false;
    return SKM_OK;
}

// ---- multi-GPU: see the skm_mg_* / skm_group_* entry points at the end of this file ----------

int32_t skm_insert_kmers_device(skm_ctx *c, const uint64_t *d_kmers, uint64_t n) {
    if (!c || (n && !d_kmers)) return SKM_ERR_INVALID_ARG;
    if (n == 0) return SKM_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    int32_t rc = insert_list(c, (const unsigned long long *)d_kmers, nullptr, n);
    if (rc) return rc;
    c->have_tot = false;
    return SKM_OK;  // asynchronous: ordered on the ctx's stream
}

int32_t skm_snapshot_histogram(skm_ctx *c, uint32_t chunk_i) {
    if (!c) return SKM_ERR_INVALID_ARG;
    if (chunk_i >= c->n_chunks) return fail(c, SKM_ERR_INVALID_ARG, "chunk_index out of range");
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    int32_t rc = snapshot_histogram(c, chunk_i);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return SKM_OK;
}

// ---- diagnostics -----------------------------------------------------------------

static int32_t pack_host_input(skm_ctx *c, const uint8_t *seqs, uint64_t n_bytes, Segment *sg, uint8_t **d_raw_out) {
    uint8_t *d_raw = nullptr;
    sg->n_bytes = n_bytes;
    sg->n_units = (n_bytes + 31) / 32;
    CU(cudaMallocAsync((void **)&d_raw, n_bytes + 1, c->stream));
    CU(cudaMemcpyAsync(d_raw, seqs, n_bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMallocAsync((void **)&sg->codes, (sg->n_units + 1) * sizeof(uint64_t), c->stream));
    CU(cudaMallocAsync((void **)&sg->breaks, (sg->n_units + 1) * sizeof(uint32_t), c->stream));
    ChunkCounters *d_cc = nullptr;
    GlobalCounters *d_gc = nullptr;
    CU(cudaMallocAsync((void **)&d_cc, sizeof(ChunkCounters), c->stream));
    CU(cudaMallocAsync((void **)&d_gc, sizeof(GlobalCounters), c->stream));
    CU(cudaMemsetAsync(d_cc, 0, sizeof(ChunkCounters), c->stream));
    CU(cudaMemsetAsync(d_gc, 0xFF, sizeof(GlobalCounters), c->stream));
    pack_kernel<<<grid_for(sg->n_units, 256 * kPackUnits), 256, 0, c->stream>>>(d_raw, n_bytes, 0, sg->codes, sg->breaks,
                                                                  sg->n_units, d_cc, d_gc);
    c->launches++;
    CU(cudaGetLastError());
    GlobalCounters gc;
    CU(cudaMemcpyAsync(&gc, d_gc, sizeof gc, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFreeAsync(d_cc, c->stream));
    CU(cudaFreeAsync(d_gc, c->stream));
    *d_raw_out = d_raw;
    if (gc.first_bad != ~0ull) {
        CU(cudaFreeAsync(d_raw, c->stream));
        CU(cudaFreeAsync(sg->codes, c->stream));
        CU(cudaFreeAsync(sg->breaks, c->stream));
        const unsigned ch = (unsigned)(gc.first_bad & 0xFF);
        return fail(c, SKM_ERR_INVALID_BASE,
                    "Invalid character '%c' in sequence. Only ACGTN allowed. (byte %llu of the ingested stream)",
                    (ch >= 0x20 && ch < 0x7F) ? (char)ch : '?', (unsigned long long)(gc.first_bad >> 8));
    }
    return SKM_OK;
}

int32_t skm_extract_kmers(skm_ctx *c, const uint8_t *seqs, uint64_t n_bytes, uint64_t *out) {
    if (!c || (n_bytes && (!seqs || !out))) return SKM_ERR_INVALID_ARG;
    if (n_bytes == 0) return SKM_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    Segment sg;
    uint8_t *d_raw = nullptr;
    int32_t rc = pack_host_input(c, seqs, n_bytes, &sg, &d_raw);
    if (rc) return rc;
    unsigned long long *d_out = nullptr;
    CU(cudaMallocAsync((void **)&d_out, n_bytes * sizeof(uint64_t), c->stream));
    extract_positions_kernel<<<grid_for(sg.n_units, 256), 256, 0, c->stream>>>(sg.codes, sg.breaks, sg.n_units,
                                                                               n_bytes, c->p.k, d_out);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d_out, n_bytes * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFreeAsync(d_out, c->stream));
    CU(cudaFreeAsync(d_raw, c->stream));
    CU(cudaFreeAsync(sg.codes, c->stream));
    CU(cudaFreeAsync(sg.breaks, c->stream));
    return SKM_OK;
}

int32_t skm_pack(skm_ctx *c, const uint8_t *seqs, uint64_t n_bytes, uint64_t *codes, uint32_t *breaks) {
    if (!c || (n_bytes && (!seqs || !codes || !breaks))) return SKM_ERR_INVALID_ARG;
    if (n_bytes == 0) return SKM_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    Segment sg;
    uint8_t *d_raw = nullptr;
    int32_t rc = pack_host_input(c, seqs, n_bytes, &sg, &d_raw);
    if (rc) return rc;
    CU(cudaMemcpyAsync(codes, sg.codes, sg.n_units * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(breaks, sg.breaks, sg.n_units * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFreeAsync(d_raw, c->stream));
    CU(cudaFreeAsync(sg.codes, c->stream));
    CU(cudaFreeAsync(sg.breaks, c->stream));
    return SKM_OK;
}

int32_t skm_synth_device(skm_ctx *c, uint64_t seed, uint64_t genome_len, uint32_t read_len, uint32_t sub_thresh,
                         uint32_t n_thresh, uint32_t chunk_index, uint32_t n_chunks, uint64_t first, uint64_t n,
                         uint8_t *d_out) {
    if (!c || (n && !d_out)) return SKM_ERR_INVALID_ARG;
    if (genome_len < read_len || read_len == 0 || n_chunks == 0) return fail(c, SKM_ERR_INVALID_ARG, "bad synth parameters");
    if (n == 0) return SKM_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    skm_synth_params sp{seed, genome_len, read_len, sub_thresh, n_thresh, 0};
    synth_kernel<<<c->sm_count * 16, 256, 0, c->stream>>>(sp, chunk_index, n_chunks, first, n, d_out);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return SKM_OK;
}

int32_t skm_bench_gups(skm_ctx *c, uint32_t log2_slots, uint64_t n_updates, uint32_t iters, int32_t variant,
                       float *ms_out) {
    if (!c || !ms_out || log2_slots < 10 || log2_slots > 36 || iters == 0) return SKM_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    Slot *t = nullptr;
    const uint64_t cap = 1ull << log2_slots;
    CU(cudaMalloc((void **)&t, cap * sizeof(Slot)));
    CU(cudaMemsetAsync(t, 0, cap * sizeof(Slot), c->stream));
    unsigned long long *d_sink = nullptr;
    CU(cudaMalloc((void **)&d_sink, 8));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const uint32_t region_log2 = std::min<uint32_t>(17, log2_slots);
    const uint32_t grid = variant >= 3 ? grid_for(n_updates, 256) : c->sm_count * 8;
    gups_kernel<<<grid, 256, 0, c->stream>>>(t, log2_slots, n_updates, 1, variant, region_log2, d_sink);  // warm-up
    cudaEventRecord(a, c->stream);
    for (uint32_t i = 0; i < iters; i++)
        gups_kernel<<<grid, 256, 0, c->stream>>>(t, log2_slots, n_updates, 1000003ull * (i + 2), variant,
                                                 region_log2, d_sink);
    cudaEventRecord(b, c->stream);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    *ms_out = ms / iters;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(t);
    cudaFree(d_sink);
    return SKM_OK;
}

}  // extern "C"

// ============================================================================
// multi-GPU: receive arenas, collective finalize, in-process group
// ============================================================================

namespace {

int32_t mg_allgather(skm_ctx *c, const skm_comm *comm, const void *send, void *recv, uint64_t bytes) {
    const int32_t rc = comm->allgather(comm->user, send, recv, bytes);
    if (rc) return fail(c, SKM_ERR_STATE, "all-gather callback failed (%d)", rc);
    return SKM_OK;
}

// Everything this rank has ingested is bucketed, tile-sorted and on its way to the owners.
int32_t mg_prepare_local(skm_ctx *c) {
    int32_t rc;
    CU(cudaStreamSynchronize(c->copy_stream));
    CU(cudaStreamSynchronize(c->pack_stream));
    CU(cudaStreamSynchronize(c->part_stream));
    CU(cudaStreamSynchronize(c->sort_stream));
    WorkStream ws(c, c->part_stream);
    const uint32_t nb_coarse = c->n_ranks << route_log2_regions(c);
    auto build = [&](uint32_t ch, size_t si, bool exact) -> int32_t {
        Segment &sg = c->chunks[ch].segs[si];
        uint64_t *h_off = alloc_offsets(c, nb_coarse + 1);
        if (!h_off) return fail(c, SKM_ERR_OOM, "pinned allocation failed");
        int32_t r = build_list(c, ch, si, h_off, exact || !c->capped || sg.n_bytes / nb_coarse < 4096, /*must=*/true, 1, /*sort=*/false);
        if (r) return r;
        r = refine_owner_lists(c, ch, si, exact);
        if (r) return r;
        CU(cudaEventSynchronize(sg.ready));
        return SKM_OK;
    };
    // k-mers of owner o in a list bucketed by (owner, coarse region), from its bucket totals / offsets
    const uint32_t R = 1u << route_log2_regions(c);
    auto owner_kmers = [&](const Segment &sg, uint32_t o) -> uint64_t {
        uint64_t nk = 0;
        for (uint32_t b = o * R; b < (o + 1) * R; b++)
            nk += sg.cap ? std::min<uint64_t>(sg.h_offsets[b], sg.cap) : sg.h_offsets[b + 1] - sg.h_offsets[b];
        return nk;
    };
    for (uint32_t ch = 0; ch < c->n_chunks && c->mg_slices; ch++) {
        ChunkState &cs = c->chunks[ch];
        for (size_t si = 0; si < cs.segs.size(); si++) {
            Segment &sg = cs.segs[si];
            auto build_sorted = [&](bool exact) -> int32_t {
                uint64_t *h_off = alloc_offsets(c, nb_coarse + 1);
                if (!h_off) return fail(c, SKM_ERR_OOM, "pinned allocation failed");
                return build_list(c, ch, si, h_off, exact || !c->capped || sg.n_bytes / nb_coarse < 4096, /*must=*/true, 1, /*sort=*/true);
            };
            if (!sg.list && sg.codes) {  // skipped at ingest time (memory): build it now
                rc = build_sorted(false);
                if (rc) return rc;
            }
            if (!sg.list) continue;
            CU(cudaEventSynchronize(sg.ready));   // the bucket totals are on the host
            if (sg.cap && sg.h_offsets[sg.n_buckets] > 0) {
                // the capped layout overflowed (heavily repeated k-mers): what was shipped is void; rebuild the
                // batch with the exact layout from the packed form and ship that
                if (sg.shipped)
                    for (uint32_t i = 0; i + 1 < c->n_ranks; i++) c->mg_sent[sg.rec_first + i].dead = 1;
                rc = sync_dma(c);  // the copies still read the old list
                if (rc) return rc;
                rc = drop_capped_list(c, ch, sg, c->part_stream);
                if (rc) return rc;
                sg.shipped = false;
                c->n_capped_fallbacks++;
                rc = build_sorted(true);
                if (rc) return rc;
                CU(cudaEventSynchronize(sg.ready));
            }
            rc = ship_slices(c, ch, si, /*sizes_on_host=*/true);
            if (rc) return rc;
            if (!sg.shipped) return fail(c, SKM_ERR_STATE, "receive arenas are not wired (skm_mg_open_peer / skm_mg_set_peer for every peer)");
            for (uint32_t i = 1; i < c->n_ranks; i++) c->mg_sent[sg.rec_first + i - 1].n_kmers = owner_kmers(sg, (c->p.rank + i) % c->n_ranks);
            if (sg.codes) {  // the packed form was kept for the overflow retry only
                CU(cudaFreeAsync(sg.codes, c->part_stream));
                CU(cudaFreeAsync(sg.breaks, c->part_stream));
                sg.codes = nullptr;
                sg.breaks = nullptr;
            }
        }
    }
    for (uint32_t ch = 0; ch < c->n_chunks && !c->mg_slices; ch++) {
        ChunkState &cs = c->chunks[ch];
        for (size_t si = 0; si < cs.segs.size(); si++) {
            Segment &sg = cs.segs[si];
            if (sg.owners.empty() && sg.codes) {  // skipped at ingest time (memory): build it now
                rc = build(ch, si, false);
                if (rc) return rc;
            }
            if (sg.owners.empty()) continue;
            bool overflow = sg.cap && sg.h_offsets[sg.n_buckets] > 0;
            for (auto &ol : sg.owners) overflow = overflow || (ol.cap && ol.h_off[kFineRegions] > 0);
            if (overflow) {
                // a capped layout overflowed (heavily repeated k-mers, or a lopsided owner): what was shipped is
                // void; rebuild the batch with exact layouts from the packed form and ship that
                for (auto &ol : sg.owners)
                    if (ol.shipped && ol.rec < c->mg_sent.size()) c->mg_sent[ol.rec].dead = 1;
                rc = sync_dma(c);  // the copies still read the old lists
                if (rc) return rc;
                rc = drop_capped_list(c, ch, sg, c->part_stream);
                if (rc) return rc;
                sg.shipped = false;
                c->n_capped_fallbacks++;
                rc = build(ch, si, true);
                if (rc) return rc;
            }
            rc = ship_segment(c, ch, si, /*sizes_on_host=*/true);
            if (rc) return rc;
            for (uint32_t o = 0; o < c->n_ranks; o++) {
                OwnerList &ol = sg.owners[o];
                if (o == c->p.rank) continue;
                if (!ol.shipped) return fail(c, SKM_ERR_STATE, "receive arenas are not wired (skm_mg_open_peer / skm_mg_set_peer for every peer)");
                uint64_t nk = 0;   // k-mers for this destination, now that the totals are on the host
                if (ol.cap) for (uint32_t b = 0; b < kFineRegions; b++) nk += std::min<uint64_t>(ol.h_off[b], ol.cap);
                else nk = ol.h_off[kFineRegions];
                c->mg_sent[ol.rec].n_kmers = nk;
            }
            if (sg.codes) {  // the packed form was kept for the overflow retry only
                CU(cudaFreeAsync(sg.codes, c->part_stream));
                CU(cudaFreeAsync(sg.breaks, c->part_stream));
                sg.codes = nullptr;
                sg.breaks = nullptr;
            }
        }
    }
    DBG_SYNC("mg_prepare_local (bucketing / shipping)");
    CU(cudaStreamSynchronize(c->part_stream));
    CU(cudaStreamSynchronize(c->sort_stream));
    rc = sync_dma(c);  // every slice of mine has landed
    if (rc) return rc;
    rc = check_sticky(c);
    if (rc) return rc;
    return refresh_chunk_counters(c);
}

// The collective part of a finalize is host all-gathers through a callback (gloo over loopback under
// torchrun: some 0.5 ms each at 8 ranks, and 10 ms for 8 x 800 KB of histogram columns).  So the two
// messages that are almost always small travel inside the fixed-size ones: up to kInlineRecs
// directory records inside the header, and the first kInlineBins bins of every column behind the
// totals (bins above a rank's highest occupied one are zero).  Two all-gathers per finalize then.
constexpr uint32_t kInlineRecs = 160;
constexpr size_t kInlineBins = 512;

struct MgHeader {
    int32_t status;
    uint32_t n_records;
    uint64_t n_windows, n_reads;
    MgRecord rec[kInlineRecs];
};

struct MgTotals {
    int32_t status;
    uint32_t top_bin;   // 1 + the highest occupied bin over this rank's columns
    uint64_t n_kmers, n_distinct_scan, n_distinct, n_saturated, n_windows;
};

// final = false: skm_mg_flush — count what has been ingested so far and free the lists / arenas
// (chunks == 0 only: without histogram columns the order of counting does not matter).
int32_t mg_finalize_impl(skm_ctx *c, const skm_comm *comm, bool final) {
    const uint32_t N = c->n_ranks, me = c->p.rank;
    const size_t nbins = c->p.histo_max + 2;
    // ---- phase 1: local work, then the directory of what everybody shipped (doubles as the
    //      "all slices have landed" barrier: a rank joins only after its own copies are done) ----
    std::vector<MgHeader> hdrs(N + 1);
    MgHeader &mine = hdrs[N];
    mine = MgHeader{};
    mine.status = mg_prepare_local(c);
    mine.n_records = (uint32_t)c->mg_sent.size();
    for (auto &cc : c->h_cc) {
        mine.n_windows += cc.n_windows;
        mine.n_reads += cc.n_reads;
    }
    std::copy(c->mg_sent.begin(), c->mg_sent.begin() + std::min<size_t>(c->mg_sent.size(), kInlineRecs), mine.rec);
    int32_t rc = mg_allgather(c, comm, &mine, hdrs.data(), sizeof(MgHeader));
    if (rc) return rc;
    uint32_t max_rec = 0;
    uint64_t windows_all = 0, reads_all = 0;
    for (uint32_t r = 0; r < N; r++) {
        if (hdrs[r].status != SKM_OK) {
            if (mine.status != SKM_OK) return mine.status;  // own message is already set
            return fail(c, hdrs[r].status, "rank %u failed before the exchange (see its skm_last_error)", r);
        }
        max_rec = std::max(max_rec, hdrs[r].n_records);
        windows_all += hdrs[r].n_windows;
        reads_all += hdrs[r].n_reads;
    }
    if (reads_all == 0 && final)  // src/io.rs:578-580
        return fail(c, SKM_ERR_NO_READS, "No reads were ingested. Check that input files contain valid FASTQ records.");
    std::vector<MgRecord> all, sendbuf(std::max(max_rec, 1u));
    if (max_rec > kInlineRecs) {   // a long directory (many batches): one more all-gather for all of it
        all.resize((size_t)N * sendbuf.size());
        std::copy(c->mg_sent.begin(), c->mg_sent.end(), sendbuf.begin());
        rc = mg_allgather(c, comm, sendbuf.data(), all.data(), (uint64_t)sendbuf.size() * sizeof(MgRecord));
        if (rc) return rc;
    }
    auto record = [&](uint32_t src, uint32_t i) -> const MgRecord & {
        return max_rec > kInlineRecs ? all[(size_t)src * sendbuf.size() + i] : hdrs[src].rec[i];
    };

    // ---- phase 2: the lists this rank counts, in chunk order: its own + the slices in its arena ----
    struct Item { SegDesc d; uint32_t chunk; uint64_t n_kmers; };
    std::vector<Item> items;
    const uint32_t g1 = list_log2_regions(c);
    uint64_t kmers_mine = 0;
    for (uint32_t ch = 0; ch < c->n_chunks && c->mg_slices; ch++)
        for (auto &sg : c->chunks[ch].segs) {
            if (!sg.list || !sg.tiled) continue;
            // this rank's own slice of its own list, read in place (bucket0 = rank << g1)
            uint64_t nk = 0;
            const uint32_t R = 1u << g1;
            for (uint32_t b = me * R; b < (me + 1) * R; b++)
                nk += sg.cap ? std::min<uint64_t>(sg.h_offsets[b], sg.cap) : sg.h_offsets[b + 1] - sg.h_offsets[b];
            items.push_back(Item{local_desc(c, sg, 0u), ch, nk});
            kmers_mine += nk;
        }
    for (uint32_t ch = 0; ch < c->n_chunks && !c->mg_slices; ch++)
        for (auto &sg : c->chunks[ch].segs) {
            if (sg.owners.empty()) continue;
            const OwnerList &ol = sg.owners[me];
            uint64_t nk = 0;
            if (ol.cap) for (uint32_t b = 0; b < kFineRegions; b++) nk += std::min<uint64_t>(ol.h_off[b], ol.cap);
            else nk = ol.h_off[kFineRegions];
            const ListMeta m = list_meta_at(ol.meta, kFineRegions);
            items.push_back(Item{SegDesc{ol.list, ol.tile_off, m.tile_begin, m.cell_begin, 0u, 0u, 0u, 0u}, ch, nk});
            kmers_mine += nk;
        }
    for (uint32_t src = 0; src < N; src++)
        for (uint32_t i = 0; i < hdrs[src].n_records; i++) {
            const MgRecord &rec = record(src, i);
            if (rec.dst != me || rec.dead) continue;
            if (rec.chunk >= c->n_chunks) return fail(c, SKM_ERR_STATE, "rank %u shipped chunk %u (this ctx has %u chunks)", src, rec.chunk, c->n_chunks);
            uint8_t *base = c->mg_arena + (size_t)src * c->mg_sub_bytes;
            SegDesc d{};
            d.list = (const unsigned long long *)(base + rec.off_cells);
            d.tile_off = (const tile_off_t *)(base + rec.off_tile_off);
            d.tile_begin = (const uint32_t *)(base + rec.off_tile_begin);
            d.cell_begin = (const unsigned long long *)(base + rec.off_cell_begin);
            d.bucket0 = 0;
            // slices: the arrays are the sender's, cut at this owner's first bucket (indices relative to entry 0);
            // a whole owner list has its metadata in the receiver's numbering already
            d.rel = c->mg_slices ? 1u : 0u;
            items.push_back(Item{d, rec.chunk, rec.n_kmers});
            kmers_mine += rec.n_kmers;
        }
    std::stable_sort(items.begin(), items.end(), [](const Item &a, const Item &b) { return a.chunk < b.chunk; });

    // ---- phase 3: size the (new) table for this rank's share, then count, chunk group by chunk group ----
    int32_t status = SKM_OK;
    c->recount_valid = false;
    if (c->table_fresh && !c->p.capacity_hint) {
        uint64_t bound = kmers_mine;
        if (c->p.k < 31) bound = std::min<uint64_t>(bound, ((1ull << (2 * c->p.k)) / 2 + (1ull << c->p.k)) / N + 1024);
        uint32_t want = std::max(c->log2cap, ceil_log2((uint64_t)((double)bound / kTargetLoad) + 1));
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
            while (want > c->log2cap && ((size_t)sizeof(Slot) << want) > (free_b + c->capacity * sizeof(Slot)) / 2) want--;
        if (want > c->log2cap) status = grow_table(c, want);
    }
    std::vector<SegDesc> group;
    uint32_t group_chunk0 = 0;
    auto flush = [&](uint32_t end_chunk) -> int32_t {
        if (group.empty()) {
            // chunks without k-mers for this rank: their columns are the running histogram
            for (uint32_t ch = group_chunk0; ch < end_chunk && c->p.chunks > 0; ch++) {
                int32_t r = snapshot_histogram(c, ch);
                if (r) return r;
            }
            group_chunk0 = end_chunk;
            return SKM_OK;
        }
        int32_t r = launch_tiled(c, group, group_chunk0, end_chunk - group_chunk0);
        group.clear();
        group_chunk0 = end_chunk;
        return r;
    };
    size_t at = 0;
    for (uint32_t ch = 0; ch < c->n_chunks && status == SKM_OK; ch++) {
        size_t end = at;
        while (end < items.size() && items[end].chunk == ch) end++;
        const uint32_t pbits = c->log2cap - kPartLog2;
        const uint64_t nbr = pbits >= g1 ? 1ull : (1ull << (g1 - pbits));
        if ((group.size() + (end - at)) * nbr > kMaxVseg || ch - group_chunk0 >= kMaxChunksPerLaunch) status = flush(ch);
        for (; at < end && status == SKM_OK; at++) {
            if ((group.size() + 1) * nbr > kMaxVseg) {  // one chunk, more lists than a launch reads: several launches
                status = flush(ch + 1);
                group_chunk0 = ch;
                if (status) break;
            }
            SegDesc d = items[at].d;
            d.chunk = ch - group_chunk0;
            group.push_back(d);
            c->insert_kmers += items[at].n_kmers;
        }
    }
    if (status == SKM_OK) status = flush(c->n_chunks);
    if (status == SKM_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) status = fail(c, SKM_ERR_CUDA, "stream synchronisation failed");
    if (status == SKM_OK) {
        materialize_cols(c);
        status = check_sticky(c);
    }

    // ---- phase 4: totals and histogram columns over all ranks ----
    const size_t wbins = c->p.chunks > 0 ? std::min(nbins, kInlineBins) : 0;
    const size_t msg_words = (sizeof(MgTotals) + 7) / 8 + (size_t)c->n_chunks * wbins;
    std::vector<uint64_t> msg(msg_words, 0), msgs((size_t)N * msg_words);
    MgTotals tm{};
    tm.status = status;
    if (status == SKM_OK) {
        if (c->recount_valid) {
            if (cudaMemcpy(&c->last_tot, c->d_tot, sizeof(HistoTotals), cudaMemcpyDeviceToHost) != cudaSuccess)
                tm.status = fail(c, SKM_ERR_CUDA, "read-back of the totals failed");
            c->have_tot = true;
        } else {
            tm.status = scan_table(c, false, nullptr);
        }
        uint64_t d = 0;
        if (tm.status == SKM_OK) tm.status = read_distinct(c, &d);
        c->distinct_ub = d;
        tm.n_kmers = c->last_tot.n_kmers;
        tm.n_distinct_scan = c->last_tot.n_distinct;
        tm.n_distinct = d;
        tm.n_saturated = c->last_tot.n_saturated;
        tm.n_windows = mine.n_windows;
        for (uint32_t ch = 0; ch < c->n_chunks && c->p.chunks > 0 && tm.status == SKM_OK; ch++) {
            if (!c->have_histo[ch] || c->histos[ch].size() != nbins) {
                tm.status = fail(c, SKM_ERR_STATE, "internal: column %u missing", ch);
                break;
            }
            const std::vector<uint64_t> &h = c->histos[ch];
            size_t top = nbins;
            while (top > 0 && h[top - 1] == 0) top--;
            tm.top_bin = std::max(tm.top_bin, (uint32_t)top);
            std::copy(h.begin(), h.begin() + wbins, msg.begin() + (sizeof(MgTotals) + 7) / 8 + (size_t)ch * wbins);
        }
    }
    std::memcpy(msg.data(), &tm, sizeof(MgTotals));
    rc = mg_allgather(c, comm, msg.data(), msgs.data(), (uint64_t)msg_words * 8);
    if (rc) return rc;
    uint64_t g_kmers = 0, g_scan = 0, g_distinct = 0, g_sat = 0;
    uint32_t g_top = 0;
    for (uint32_t r = 0; r < N; r++) {
        MgTotals t;
        std::memcpy(&t, msgs.data() + (size_t)r * msg_words, sizeof(MgTotals));
        if (t.status != SKM_OK) {
            if (tm.status != SKM_OK) return tm.status;
            return fail(c, t.status, "rank %u failed while counting (see its skm_last_error)", r);
        }
        g_kmers += t.n_kmers;
        g_scan += t.n_distinct_scan;
        g_distinct += t.n_distinct;
        g_sat += t.n_saturated;
        g_top = std::max(g_top, t.top_bin);
    }
    // conservation identities on the global totals (src/io.rs:1042-1047, 1120-1132)
    if (g_sat == 0 && g_kmers != windows_all)
        return fail(c, SKM_ERR_CONSERVATION,
                    "The total count of hashed kmers (%llu) does not equal the number of ingested kmers (%llu)",
                    (unsigned long long)g_kmers, (unsigned long long)windows_all);
    if (g_scan != g_distinct)
        return fail(c, SKM_ERR_CONSERVATION,
                    "The total count of unique kmers in the histogram (%llu) does not equal the total count of hashed kmers (%llu)",
                    (unsigned long long)g_scan, (unsigned long long)g_distinct);
    if (c->p.chunks > 0) {
        // column i = sum over the partitions' columns (a k-mer lives in exactly one partition)
        const uint64_t *src0 = msgs.data() + (sizeof(MgTotals) + 7) / 8;
        size_t stride_rank = msg_words, stride_chunk = wbins, width = wbins;
        std::vector<uint64_t> mycols, allcols;
        if (g_top > wbins) {   // somebody has counts beyond the inline bins: gather the whole columns
            mycols.resize((size_t)c->n_chunks * nbins);
            allcols.resize((size_t)N * c->n_chunks * nbins);
            for (uint32_t ch = 0; ch < c->n_chunks; ch++)
                std::copy(c->histos[ch].begin(), c->histos[ch].end(), mycols.begin() + (size_t)ch * nbins);
            rc = mg_allgather(c, comm, mycols.data(), allcols.data(), (uint64_t)mycols.size() * sizeof(uint64_t));
            if (rc) return rc;
            src0 = allcols.data();
            stride_rank = (size_t)c->n_chunks * nbins;
            stride_chunk = width = nbins;
        }
        for (uint32_t ch = 0; ch < c->n_chunks; ch++) {
            std::vector<uint64_t> &h = c->histos[ch];
            std::fill(h.begin(), h.end(), 0);
            for (uint32_t r = 0; r < N; r++) {
                const uint64_t *src = src0 + (size_t)r * stride_rank + (size_t)ch * stride_chunk;
                for (size_t b = 0; b < width; b++) h[b] += src[b];
            }
        }
        const std::vector<uint64_t> &last = c->histos[c->n_chunks - 1];
        uint64_t uniq = 0;
        for (size_t i = 1; i < last.size(); i++) uniq += last[i];
        if (uniq != g_distinct)
            return fail(c, SKM_ERR_CONSERVATION,
                        "The total count of unique kmers in the histogram (%llu) does not equal the total count of hashed kmers (%llu)",
                        (unsigned long long)uniq, (unsigned long long)g_distinct);
    }
    // lists and arena contents are no longer needed; the last all-gather above is also the barrier
    // after which a peer may write into this rank's arena again (next sample)
    for (auto &cs : c->chunks) {
        for (auto &sg : cs.segs) {
            release_list(c, sg, c->stream);
            if (sg.ready) c->event_pool.push_back(sg.ready);
            sg.ready = nullptr;
        }
        if (!final) cs.segs.clear();   // (the chunk's byte and read counters stay)
    }
    if (!final) {
        for (auto &cur : c->mg_cursor) cur = 0;
        c->mg_sent.clear();
    }
    collect_spans(c);
    return SKM_OK;
}

}  // namespace

extern "C" {

int32_t skm_mg_arena_create(skm_ctx *c, uint64_t bytes) {
    if (!c || bytes == 0) return SKM_ERR_INVALID_ARG;
    if (c->n_ranks > skm_ctx::kMaxPeers) return fail(c, SKM_ERR_INVALID_ARG, "at most %u ranks", skm_ctx::kMaxPeers);
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    CU(cudaStreamSynchronize(c->stream));
    if (c->mg_arena) CU(cudaFree(c->mg_arena));
    c->mg_arena = nullptr;
    const size_t sub = ((size_t)(bytes / c->n_ranks) + 255) & ~(size_t)255;
    // cudaMalloc, not the stream-ordered pool: the block must be exportable through CUDA IPC
    CU(cudaMalloc((void **)&c->mg_arena, sub * c->n_ranks));
    c->mem_budget -= std::min(c->mem_budget, sub * c->n_ranks - c->mg_arena_bytes);
    c->mg_arena_bytes = sub * c->n_ranks;
    c->mg_sub_bytes = sub;
    c->mg_peer[c->p.rank] = c->mg_arena;
    return SKM_OK;
}

int32_t skm_mg_arena_handle(skm_ctx *c, uint8_t *handle64) {
    if (!c || !handle64) return SKM_ERR_INVALID_ARG;
    if (!c->mg_arena) return fail(c, SKM_ERR_STATE, "no arena: call skm_mg_arena_create first");
    DeviceGuard g(c->device);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, c->mg_arena));
    memcpy(handle64, &h, 64);
    return SKM_OK;
}

int32_t skm_mg_arena_ptr(skm_ctx *c, void **out) {
    if (!c || !out) return SKM_ERR_INVALID_ARG;
    *out = c->mg_arena;
    return SKM_OK;
}

int32_t skm_mg_open_peer(skm_ctx *c, uint32_t peer_rank, const uint8_t *handle64) {
    if (!c || !handle64 || peer_rank >= c->n_ranks || peer_rank >= skm_ctx::kMaxPeers) return SKM_ERR_INVALID_ARG;
    if (peer_rank == c->p.rank) return SKM_OK;
    DeviceGuard g(c->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->mg_peer[peer_rank] = (uint8_t *)p;
    c->mg_peer_ipc[peer_rank] = true;
    return SKM_OK;
}

int32_t skm_mg_set_peer(skm_ctx *c, uint32_t peer_rank, void *d_ptr, int32_t peer_device) {
    if (!c || peer_rank >= c->n_ranks || peer_rank >= skm_ctx::kMaxPeers) return SKM_ERR_INVALID_ARG;
    if (peer_rank == c->p.rank) return SKM_OK;
    DeviceGuard g(c->device);
    if (peer_device >= 0 && peer_device != c->device) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return fail(c, SKM_ERR_CUDA, "no peer access from device %d to device %d: %s", c->device, peer_device, cudaGetErrorString(e));
        cudaGetLastError();
    }
    c->mg_peer[peer_rank] = (uint8_t *)d_ptr;
    c->mg_peer_ipc[peer_rank] = false;
    return SKM_OK;
}

int32_t skm_mg_bytes_sent(skm_ctx *c, uint64_t *out) {
    if (!c || !out) return SKM_ERR_INVALID_ARG;
    *out = c->mg_bytes_sent;
    return SKM_OK;
}

static int32_t mg_finalize_or_flush(skm_ctx *c, const skm_comm *comm, bool final) {
    if (!c || !comm || !comm->allgather) return SKM_ERR_INVALID_ARG;
    if (getenv("SKM_DEBUG")) {
        std::fprintf(stderr, "[skm] rank %u host ms since last report: build %.1f (mallocs %.1f, launches %.1f) refine %.1f (mallocs %.1f) ship %.1f\n",
                     c->p.rank, c->host_ms[0], c->host_ms[4], c->host_ms[5], c->host_ms[2], c->host_ms[1], c->host_ms[3]);
        for (auto &x : c->host_ms) x = 0;
    }
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard g(c->device);
    if (c->finalized) return fail(c, SKM_ERR_STATE, final ? "finalize called twice" : "flush after finalize");
    if (!final && c->p.chunks > 0)
        return fail(c, SKM_ERR_STATE, "skm_mg_flush needs chunks == 0 (histogram columns fix the order of counting)");
    if (c->n_ranks == 1) return final ? finalize_common(c) : fail(c, SKM_ERR_STATE, "skm_mg_flush needs n_ranks > 1");
    if (!c->mg_arena) return fail(c, SKM_ERR_STATE, "no receive arena: call skm_mg_arena_create and wire the peers first");
    c->finalized = final;
    cudaEvent_t e0 = get_event(c), e1 = get_event(c);
    cudaEventRecord(e0, c->stream);
    const int32_t rc = mg_finalize_impl(c, comm, final);
    cudaEventRecord(e1, c->stream);
    if (cudaStreamSynchronize(c->stream) == cudaSuccess) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) c->stage_ms[ST_FINALIZE] += ms;
    }
    c->event_pool.push_back(e0);
    c->event_pool.push_back(e1);
    return rc;
}

int32_t skm_mg_finalize(skm_ctx *c, const skm_comm *comm) { return mg_finalize_or_flush(c, comm, true); }
int32_t skm_mg_flush(skm_ctx *c, const skm_comm *comm) { return mg_finalize_or_flush(c, comm, false); }

}  // extern "C"

// ---- all ranks in one process ------------------------------------------------------------------

#include <condition_variable>
#include <thread>

struct skm_group {
    std::vector<skm_ctx *> ctx;
    std::mutex mu;
    std::condition_variable cv;
    uint32_t arrived = 0;
    uint64_t generation = 0;
    std::vector<const void *> send;
    std::string err;

    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        const uint64_t gen = generation;
        if (++arrived == ctx.size()) {
            arrived = 0;
            generation++;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return generation != gen; });
        }
    }
};

namespace {
struct GroupMember {
    skm_group *g;
    uint32_t rank;
};
int32_t group_allgather(void *user, const void *send, void *recv, uint64_t bytes) {
    GroupMember *m = (GroupMember *)user;
    skm_group *g = m->g;
    g->send[m->rank] = send;
    g->barrier();
    for (size_t r = 0; r < g->ctx.size(); r++) memcpy((uint8_t *)recv + r * bytes, g->send[r], bytes);
    g->barrier();  // nobody's send buffer is reused before everybody has read it
    return 0;
}
}  // namespace

extern "C" {

int32_t skm_group_create(const skm_params *params, uint32_t n_ranks, const int32_t *devices, uint64_t arena_bytes_per_rank,
                         skm_group **out) {
    if (!params || !out || n_ranks == 0 || n_ranks > skm_ctx::kMaxPeers) return SKM_ERR_INVALID_ARG;
    skm_group *g = new (std::nothrow) skm_group();
    if (!g) return SKM_ERR_OOM;
    *out = g;
    g->send.assign(n_ranks, nullptr);
    for (uint32_t r = 0; r < n_ranks; r++) {
        skm_params p = *params;
        p.n_ranks = n_ranks;
        p.rank = r;
        p.device = devices ? devices[r] : -1;
        p.stream = 0;
        skm_ctx *c = nullptr;
        int32_t rc = skm_create(&p, &c);
        if (c) g->ctx.push_back(c);
        if (rc == SKM_OK && n_ranks > 1) rc = skm_mg_arena_create(c, arena_bytes_per_rank);
        if (rc) {
            g->err = c ? c->err : "skm_create failed";
            return rc;
        }
    }
    for (uint32_t a = 0; a < n_ranks && n_ranks > 1; a++)
        for (uint32_t b = 0; b < n_ranks; b++) {
            if (a == b) continue;
            const int32_t rc = skm_mg_set_peer(g->ctx[a], b, g->ctx[b]->mg_arena, g->ctx[b]->device);
            if (rc) {
                g->err = g->ctx[a]->err;
                return rc;
            }
        }
    return SKM_OK;
}

skm_ctx *skm_group_ctx(skm_group *g, uint32_t rank) { return g && rank < g->ctx.size() ? g->ctx[rank] : nullptr; }

static int32_t group_collective(skm_group *g, bool final);
int32_t skm_group_finalize(skm_group *g) { return group_collective(g, true); }
int32_t skm_group_flush(skm_group *g) { return group_collective(g, false); }

static int32_t group_collective(skm_group *g, bool final) {
    if (!g || g->ctx.empty()) return SKM_ERR_INVALID_ARG;
    const size_t n = g->ctx.size();
    std::vector<int32_t> rcs(n, SKM_OK);
    std::vector<GroupMember> members(n);
    std::vector<std::thread> threads;
    for (size_t r = 0; r < n; r++) {
        members[r] = GroupMember{g, (uint32_t)r};
        threads.emplace_back([&, r] {
            skm_comm comm{&members[r], group_allgather};
            rcs[r] = final ? skm_mg_finalize(g->ctx[r], &comm) : skm_mg_flush(g->ctx[r], &comm);
        });
    }
    for (auto &t : threads) t.join();
    for (size_t r = 0; r < n; r++)
        if (rcs[r]) {
            g->err = "rank " + std::to_string(r) + ": " + g->ctx[r]->err;
            return rcs[r];
        }
    return SKM_OK;
}

int32_t skm_group_reset(skm_group *g) {
    if (!g) return SKM_ERR_INVALID_ARG;
    for (auto c : g->ctx) {
        const int32_t rc = skm_reset(c);
        if (rc) {
            g->err = c->err;
            return rc;
        }
    }
    return SKM_OK;
}

const char *skm_group_last_error(skm_group *g) { return g ? g->err.c_str() : "null group"; }

void skm_group_destroy(skm_group *g) {
    if (!g) return;
    for (auto c : g->ctx) skm_destroy(c);
    delete g;
}

}  // extern "C"
