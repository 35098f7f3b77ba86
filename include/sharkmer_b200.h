/*
 * sharkmer_b200.h — C ABI of the B200-native k-mer counting engine.
 *
 * Drop-in boundary for the k-mer counting path of caseywdunn/sharkmer v3.1.0.
 * The reference has no FFI: its boundary is the Rust module surface re-exported
 * at src/kmer/mod.rs:10-17 (Chunk, KmerCounts, FilteredKmerCounts, Histogram,
 * kmers_from_ascii, ...), as called from src/io.rs (ingest + consolidate),
 * src/stats.rs:81 and the src/pcr modules.  Every entry point below names the reference
 * item it replaces.  INTEGRATION.md shows the Rust `extern "C"` block and the
 * build.rs (nvcc, sm_100a) a maintainer would add.
 *
 * Conventions
 *  - every call returns int32 status, SKM_OK == 0; nothing throws across the
 *    boundary; skm_last_error(ctx) is a ctx-owned NUL-terminated string.
 *  - plain pointers and sizes only.  The caller owns all host buffers; the ctx
 *    owns all device memory and streams.
 *  - a ctx is single-producer while ingesting; after skm_finalize the read
 *    entry points (histogram/totals/export/lookup) may be called from any
 *    thread, one call at a time per ctx (they serialise on an internal mutex).
 *  - there is NO CPU fallback: without a CUDA device skm_create fails with
 *    SKM_ERR_CUDA.
 */
#ifndef SHARKMER_B200_H
#define SHARKMER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKM_ABI_VERSION 1

enum {
    SKM_OK = 0,
    SKM_ERR_INVALID_ARG = 1,   /* k even / out of 1..31, histo_max out of range (src/cli.rs:662-673), bad pointers */
    SKM_ERR_INVALID_BASE = 2,  /* byte other than A,C,G,T,N in a sequence (src/kmer/encoding.rs:353-356) */
    SKM_ERR_CUDA = 3,          /* CUDA runtime failure or no device */
    SKM_ERR_OOM = 4,           /* device or pinned allocation failed */
    SKM_ERR_STATE = 5,         /* call not valid in this state (e.g. ingest after finalize) */
    SKM_ERR_CONSERVATION = 6,  /* a conservation identity failed (src/io.rs:1042-1047,1120-1132) */
    SKM_ERR_K_MISMATCH = 7,    /* merge of tables with different k (src/kmer/counting.rs:158-160) */
    SKM_ERR_NO_READS = 8,      /* nothing was ingested (src/io.rs:578-580) */
    SKM_ERR_CAPACITY = 9       /* the table could not be grown far enough (device memory), or a table partition filled up */
};

typedef struct skm_ctx skm_ctx;

/* Insert strategies.  AUTO picks per chunk. */
enum { SKM_INSERT_AUTO = 0, SKM_INSERT_DIRECT = 1, SKM_INSERT_PARTITIONED = 2 };

typedef struct skm_params {
    uint32_t struct_size;    /* sizeof(skm_params), for ABI evolution */
    uint32_t k;              /* odd, 1..31            (src/cli.rs:662-667) */
    uint32_t chunks;         /* --chunks; 0 => one internal chunk, no histograms (src/io.rs:378) */
    uint32_t insert_mode;    /* SKM_INSERT_* */
    uint64_t histo_max;      /* 1..1000000           (src/cli.rs:668-673) */
    uint64_t capacity_hint;  /* expected distinct k-mers owned by this ctx; 0 => grow on demand.
                              * n_ranks > 1: the same value on every rank (it fixes the geometry of the
                              * k-mer lists the ranks exchange) */
    int32_t device;          /* CUDA ordinal; -1 => current device */
    uint32_t n_ranks;        /* table partitions (GPUs); 0/1 => single GPU */
    uint32_t rank;           /* this ctx owns k-mers with skm_owner_rank(hash, n_ranks) == rank (a contiguous hash range) */
    uint32_t reserved;
    uint64_t stream;         /* cudaStream_t to run on (e.g. a torch stream's handle); 0 => ctx creates one */
} skm_params;

/* Chunk::get_n_reads / get_n_bases / get_n_kmers (src/kmer/chunk.rs:36-46) and
 * KmerCounts::get_n_kmers / get_n_unique_kmers (src/kmer/counting.rs:254-260). */
typedef struct skm_totals {
    uint64_t n_reads;       /* sequences ingested (reads, not N-split subreads; chunk.rs:27) */
    uint64_t n_bases;       /* non-N bases (count_valid_bases, encoding.rs:374-376) */
    uint64_t n_bases_read;  /* all sequence bytes incl. N (state.n_bases_read, io.rs:335) */
    uint64_t n_kmers;       /* k-mer occurrences = sum of table counts */
    uint64_t n_unique;      /* distinct canonical k-mers */
    uint64_t n_singletons;  /* histogram bin 1 of the last snapshot (0 when chunks == 0) */
    uint64_t n_saturated;   /* k-mers whose count reached u32::MAX (counting.rs:190-200 warns) */
} skm_totals;

/* Device time per stage, accumulated over the ctx's life (CUDA events). */
typedef struct skm_stage_ms {
    float h2d, pack, count, partition, insert, histogram, grow, total_finalize;
    /* kernel launches per stage, same order as the floats above (h2d = copies, not kernels) */
    uint32_t launches[8];
    uint32_t kernel_launches; /* all kernels launched by this ctx */
    uint32_t n_grows;
    uint64_t table_capacity; /* slots */
    uint64_t table_bytes;
    uint64_t insert_kmers;   /* k-mers handed to the insert kernels */
    uint64_t insert_bases;   /* packed bases read by the fused extract+insert kernel */
    float sort;              /* tile_sort_kernel (pass B of the tiled insert) */
    float scan;              /* skm_scan_oligos table passes */
    uint32_t sort_launches, scan_launches;
    uint32_t tiled_launches; /* tile_insert_kernel launches ... */
    uint32_t tiled_retries;  /* ... of which retries after growing the table */
} skm_stage_ms;

uint32_t skm_abi_version(void);

/* Chunk::new x n + KmerCounts::new_with_capacity (src/io.rs:378-379,1005-1006). */
int32_t skm_create(const skm_params *params, skm_ctx **out);
void skm_destroy(skm_ctx *ctx);
const char *skm_last_error(const skm_ctx *ctx);

/* Pinned host buffers for the FASTQ reader's read batches (replaces the
 * Vec<String> `state.seqs`, src/io.rs:200-206). */
int32_t skm_pinned_alloc(skm_ctx *ctx, size_t bytes, void **out);
int32_t skm_pinned_free(skm_ctx *ctx, void *ptr);

/* drain_batch -> Chunk::ingest_seq x n (src/io.rs:355-361, src/kmer/chunk.rs:25-30).
 * `seqs` = the sequence lines of the batch, each terminated by '\n' (exactly
 * FASTQ line 2 + newline), host memory.  Any number of 1000-read batches that
 * belong to the same chunk may be concatenated in one call.  An empty line is
 * an empty read and is counted.  flags: SKM_INGEST_ASYNC => the buffer stays
 * untouched until skm_sync/skm_finalize (pinned memory makes the copy truly
 * asynchronous); default: the call returns once the buffer may be reused. */
#define SKM_INGEST_ASYNC 1u
int32_t skm_ingest_batch(skm_ctx *ctx, uint32_t chunk_index, const uint8_t *seqs, uint64_t n_bytes,
                         uint32_t flags);
/* Same, for concatenated bases + n_reads+1 offsets (no separators), host memory. */
int32_t skm_ingest_reads(skm_ctx *ctx, uint32_t chunk_index, const uint8_t *bases,
                         const uint64_t *offsets, uint64_t n_reads);
/* Same as skm_ingest_batch but `d_seqs` is DEVICE memory on the ctx's device
 * (kernel-path benchmarking; reads produced on the device). */
int32_t skm_ingest_device(skm_ctx *ctx, uint32_t chunk_index, const uint8_t *d_seqs,
                          uint64_t n_bytes);
/* Wait for all queued work; reports sticky errors (invalid base). */
int32_t skm_sync(skm_ctx *ctx);

/* consolidate_and_histogram's chunk loop (src/io.rs:1016-1047,1096-1157):
 * counts chunk 0..n-1 in order into the table, snapshots the cumulative
 * histogram after each (when chunks > 0) and runs the conservation checks. */
int32_t skm_finalize(skm_ctx *ctx);
/* Empty the table, drop staged reads, histograms, counters and stage times; the
 * ctx (streams, table allocation) is reused for another sample. */
int32_t skm_reset(skm_ctx *ctx);

/* Histogram::get_vector after merging chunks 0..chunk_i (src/kmer/histogram.rs:125-134):
 * histo_max+2 entries, bin 0 = 0, last bin = all counts > histo_max. */
int32_t skm_histogram(skm_ctx *ctx, uint32_t chunk_i, uint64_t *out, uint64_t out_len);
int32_t skm_totals_get(skm_ctx *ctx, skm_totals *out);
int32_t skm_chunk_totals(skm_ctx *ctx, uint32_t chunk_index, skm_totals *out);
int32_t skm_stage_times(skm_ctx *ctx, skm_stage_ms *out);

/* KmerCounts::len / iter (src/kmer/counting.rs:239-246).  Counts are the u32
 * saturated values.  sorted != 0 => ascending k-mer order (the parity artefact). */
int32_t skm_table_len(skm_ctx *ctx, uint64_t *out);
int32_t skm_export(skm_ctx *ctx, uint64_t *keys, uint32_t *counts, uint64_t cap, int32_t sorted,
                   uint64_t *n_out);
/* Wrapping u64 sum of skm_pair_digest(kmer, count) over the table (skm_common.h). */
int32_t skm_table_digest(skm_ctx *ctx, uint64_t *out);

/* Batched lookups.  mode SKM_LOOKUP_CANONICAL: FilteredKmerCounts::get_canonical_count
 * (src/kmer/counting.rs:205-209,339-342): canonicalises (min(kmer, revcomp)), 0 if
 * absent or < min_count.  mode SKM_LOOKUP_EXACT: KmerCounts::get_count (no
 * canonicalisation).  found[i] (optional) = 1 if present and >= min_count
 * (get_canonical, counting.rs:218-222,328-336). */
/* mode SKM_LOOKUP_EITHER: KmerCounts::get_canonical (counting.rs:218-222): probe the
 * k-mer as given, else its reverse complement. */
enum { SKM_LOOKUP_CANONICAL = 0, SKM_LOOKUP_EXACT = 1, SKM_LOOKUP_EITHER = 2 };
int32_t skm_lookup_batch(skm_ctx *ctx, const uint64_t *kmers, uint64_t n, uint32_t min_count,
                         int32_t mode, uint32_t *counts, uint8_t *found);

/* find_oligos_in_kmers (src/pcr/primers.rs:163-226), sPCR's full-table scan, as one streaming
 * pass on the device.  `oligos`: unshifted 2-bit oligos of `oligo_length` bases (0 < len < k).
 * A table k-mer with count >= min_count matches if it STARTS with an oligo (reported as is),
 * else if it ENDS with the reverse complement of one (its reverse complement is reported).
 * Output in ascending k-mer order; *n_out = number of matches (call with keys = NULL to size). */
int32_t skm_scan_oligos(skm_ctx *ctx, const uint64_t *oligos, uint64_t n_oligos, uint32_t oligo_length,
                        uint32_t min_count, uint64_t *keys, uint32_t *counts, uint64_t cap,
                        uint64_t *n_out);

/* KmerCounts::insert / extend (src/kmer/counting.rs:152-166): saturating add of
 * pre-counted (k-mer, count) pairs.  Allowed before or after finalize. */
int32_t skm_insert_counts(skm_ctx *ctx, const uint64_t *keys, const uint32_t *counts, uint64_t n);

/* ---- multi-GPU: the table sharded by k-mer hash over n_ranks GPUs ----------------------------
 *
 * The reference has no multi-GPU mode; this is the path BASELINE.json's north star adds (k-mers
 * routed to the owning GPU over NVLink).  One ctx per GPU (skm_params.n_ranks / rank); every ctx
 * ingests ITS share of the read batches with the ordinary skm_ingest_* calls and owns the k-mers
 * whose hash falls in its contiguous hash range (skm_owner_rank, skm_common.h).
 *
 * Exchange.  At ingest time a batch is bucketed by (owner, table region of that owner) and
 * tile-sorted like on one GPU; the slices of the other owners are then pushed into THEIR receive
 * arenas by the copy engines (peer copies over NVLink/NVSwitch, no SM time, no collective): a
 * sender owns a fixed sub-arena inside every peer's arena and allocates in it on its own, so the
 * exchange needs no coordination and overlaps the ingest of the following batches.
 *   skm_mg_arena_create   allocates this rank's receive arena (bytes: this rank's share of all k-mers
 *                         x 8 B x ~1.2, plus ~3 %% for the tile offsets)
 *   skm_mg_arena_handle   64-byte CUDA IPC handle of the arena, for peers in other processes
 *   skm_mg_arena_ptr      the arena's device pointer, for peers in the same process
 *   skm_mg_open_peer      maps a peer's arena from its IPC handle
 *   skm_mg_set_peer       same from a raw device pointer (peer_device: its CUDA ordinal)
 * Wire the arenas before the first skm_ingest_* call (batches ingested earlier are shipped at
 * finalize instead).
 *
 * skm_mg_finalize is COLLECTIVE: every rank calls it once all of ITS batches are ingested.  The
 * caller lends it one primitive, an all-gather of host bytes among the ranks (skm_comm; a barrier
 * is an all-gather of one byte): torch.distributed / MPI / sockets across processes, or the
 * in-process implementation of skm_group below.  It waits for the exchange, counts every chunk in
 * order from the local lists and the arena (one tile_insert_kernel launch), runs the conservation
 * checks of src/io.rs:1042-1047,1120-1132 on the GLOBAL totals and leaves in every ctx the
 * histogram columns summed over all ranks (skm_histogram), while table, lookups, scans, export
 * and skm_totals_get describe the ctx's own partition.  Results do not depend on n_ranks.
 */
typedef struct skm_comm {
    void *user;
    /* every rank contributes `bytes` bytes from `send`; `recv` receives n_ranks * bytes, in rank
     * order.  Returns 0 on success.  Called from the thread that called skm_mg_finalize. */
    int32_t (*allgather)(void *user, const void *send, void *recv, uint64_t bytes);
} skm_comm;
int32_t skm_mg_arena_create(skm_ctx *ctx, uint64_t bytes);
int32_t skm_mg_arena_handle(skm_ctx *ctx, uint8_t *handle64);
int32_t skm_mg_arena_ptr(skm_ctx *ctx, void **out);
int32_t skm_mg_open_peer(skm_ctx *ctx, uint32_t peer_rank, const uint8_t *handle64);
int32_t skm_mg_set_peer(skm_ctx *ctx, uint32_t peer_rank, void *d_ptr, int32_t peer_device);
int32_t skm_mg_finalize(skm_ctx *ctx, const skm_comm *comm);
/* Inputs larger than the GPUs' memory (BASELINE config 5: 1e9 reads): COLLECTIVE, chunks == 0 only —
 * counts everything ingested so far into the tables and frees the k-mer lists and the arenas; ingest
 * goes on afterwards and skm_mg_finalize ends the run as usual. */
int32_t skm_mg_flush(skm_ctx *ctx, const skm_comm *comm);
/* bytes this rank pushed to its peers since create / reset */
int32_t skm_mg_bytes_sent(skm_ctx *ctx, uint64_t *out);

/* All ranks in ONE process (what sharkmer's single-process main, src/main.rs:112-131, would drive):
 * n ctx's on the given devices (repeats allowed: several ranks may share a GPU), arenas wired by
 * device pointers, skm_group_finalize runs skm_mg_finalize on n threads with an in-process
 * all-gather.  Ingest through skm_group_ctx(g, i) with the ordinary calls (rank i counts what it is
 * given; any split of the batches over the ranks gives the same result). */
typedef struct skm_group skm_group;
int32_t skm_group_create(const skm_params *params /* rank, n_ranks, device are filled in per member */,
                         uint32_t n_ranks, const int32_t *devices, uint64_t arena_bytes_per_rank,
                         skm_group **out);
skm_ctx *skm_group_ctx(skm_group *g, uint32_t rank);
int32_t skm_group_finalize(skm_group *g);
int32_t skm_group_flush(skm_group *g);
int32_t skm_group_reset(skm_group *g);
const char *skm_group_last_error(skm_group *g);
void skm_group_destroy(skm_group *g);

/* Insert `n` k-mers (device memory) that this rank owns; asynchronous on the ctx's stream. */
int32_t skm_insert_kmers_device(skm_ctx *ctx, const uint64_t *d_kmers, uint64_t n);
/* Snapshot this rank's partial histogram of its table partition as column
 * `chunk_i` (the caller sums columns over ranks). */
int32_t skm_snapshot_histogram(skm_ctx *ctx, uint32_t chunk_i);

/* ---- diagnostics / test entry points (same kernels, small inputs) -------- */

/* kmers_from_ascii on the device (src/kmer/encoding.rs:332-371): for newline-
 * terminated sequences `seqs` (host), out[p] = canonical k-mer of the window
 * ENDING at byte p, or SKM_EMPTY (all ones) where no window ends.  out has
 * n_bytes entries, so order of emission is checkable exactly. */
int32_t skm_extract_kmers(skm_ctx *ctx, const uint8_t *seqs, uint64_t n_bytes, uint64_t *out);
/* The pack kernel's output for `seqs` (host): 2-bit codes MSB-first, 32 bases
 * per u64 (same bit order as Read::from_str, src/kmer/encoding.rs:60-95), and
 * the break mask (1 = N / separator / padding), MSB-first, 32 bases per u32.
 * Both arrays have ceil(n_bytes/32) entries. */
int32_t skm_pack(skm_ctx *ctx, const uint8_t *seqs, uint64_t n_bytes, uint64_t *codes,
                 uint32_t *breaks);
/* Synthetic reads on the device (skm_common.h): writes the
 * newline-terminated sequence lines of chunk-local reads [first, first+n) of
 * chunk `chunk_index` (global read = skm_chunk_read_to_global) into d_out
 * (n * (read_len+1) bytes, device memory). */
int32_t skm_synth_device(skm_ctx *ctx, uint64_t seed, uint64_t genome_len, uint32_t read_len,
                         uint32_t sub_thresh, uint32_t n_thresh, uint32_t chunk_index,
                         uint32_t n_chunks, uint64_t first, uint64_t n, uint8_t *d_out);
int32_t skm_device_alloc(skm_ctx *ctx, size_t bytes, void **out);
int32_t skm_device_free(skm_ctx *ctx, void *ptr);
int32_t skm_memcpy_d2h(skm_ctx *ctx, void *dst, const void *d_src, size_t bytes);
int32_t skm_memcpy_h2d(skm_ctx *ctx, void *d_dst, const void *src, size_t bytes);

/* Random-access roofline probe ("GUPS"): n_updates updates of uniformly random
 * 16-byte slots (key load + 64-bit RED add, i.e. the insert kernel's memory
 * behaviour without extraction) on a table of 2^log2_slots slots.  Returns the
 * kernel's device time in ms (average over `iters`). */
int32_t skm_bench_gups(skm_ctx *ctx, uint32_t log2_slots, uint64_t n_updates, uint32_t iters,
                       int32_t variant, float *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* SHARKMER_B200_H */
