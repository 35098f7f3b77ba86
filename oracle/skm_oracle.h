/*
 * skm_oracle.h — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the k-mer counting path of caseywdunn/sharkmer
 * v3.1.0 (Rust; cannot be compiled in this environment — no cargo/rustc):
 *
 *   src/kmer/encoding.rs   kmers_from_ascii, count_valid_bases, revcomp_kmer,
 *                          Read::from_str / Read::get_kmers / seq_to_reads
 *   src/kmer/counting.rs   KmerCounts (saturating u32 upsert, extend,
 *                          extend_with_histogram, canonical lookups, totals)
 *   src/kmer/chunk.rs      Chunk::ingest_seq
 *   src/kmer/histogram.rs  Histogram (dense bins + sparse tail, move_count,
 *                          get_vector)
 *   src/io.rs              read_fastq / read_fastq_paired / drain_batch
 *                          batching, consolidate_and_histogram, .histo writers
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library.  The product
 * (sharkmer_b200/, include/sharkmer_b200.h) never links or calls it.
 *
 * Parity pinning: every known-answer vector the reference's own unit tests
 * hold for this path (src/kmer/mod.rs:61-305, src/kmer/counting.rs:365-510,
 * src/pcr/mod.rs:1236-1342) is replayed against this file by
 * tests/test_oracle_kats.py.  The reference's hash (ahash, randomly keyed)
 * is unobservable; the oracle's own hash map is an implementation detail and
 * all comparisons are made on sorted (k-mer, count) pairs.
 */
#ifndef SKM_ORACLE_H
#define SKM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_OK 0
#define ORC_ERR_INVALID_BASE (-1)  /* encoding.rs:353-356 */
#define ORC_ERR_BAD_K (-2)         /* encoding.rs:333 */
#define ORC_ERR_K_MISMATCH (-3)    /* counting.rs:158-160 */
#define ORC_ERR_CONSERVATION (-4)  /* io.rs:1042-1047,1120-1132 */
#define ORC_ERR_IO (-5)
#define ORC_ERR_FASTQ (-6)         /* io.rs:161-198, 291-318 */
#define ORC_ERR_NO_READS (-7)      /* io.rs:578-580 */

/* ---- encoding.rs ------------------------------------------------------- */

/* kmers_from_ascii (encoding.rs:332-371).  `out` must hold at least
 * max(len,1) entries.  Returns the number of k-mers or a negative error. */
int64_t orc_kmers_from_ascii(const char *seq, size_t len, uint32_t k, uint64_t *out);
/* count_valid_bases (encoding.rs:374-376). */
uint64_t orc_count_valid_bases(const char *seq, size_t len);
/* revcomp_kmer (encoding.rs:235-262), restated with the byte LUT. */
uint64_t orc_revcomp_kmer(uint64_t kmer, uint32_t k);
/* seq_to_kmer (encoding.rs:379-392); returns ORC_ERR_INVALID_BASE via *err. */
uint64_t orc_seq_to_kmer(const char *seq, size_t len, int *err);
/* kmer_to_seq (encoding.rs:311-325); writes k chars + NUL. */
void orc_kmer_to_seq(uint64_t kmer, uint32_t k, char *out);

/* Second, independent encoder (the reference's test-only `Read` pipeline):
 * Read::from_str (encoding.rs:60-95): MSB-first 2-bit packing, 4 bases/byte.
 * `out` must hold len/4+1 bytes.  Returns number of bytes written or <0. */
int64_t orc_read_pack(const char *seq, size_t len, uint8_t *out);
/* Read::get_kmers (encoding.rs:132-189) on a packed subread of `length` bases.
 * Like the reference it emits the k-mers of the padding bases first and then
 * truncates them, so `out` needs 4*n_bytes entries. */
int64_t orc_read_get_kmers(const uint8_t *packed, size_t n_bytes, size_t length, uint32_t k,
                           uint64_t *out);
/* seq_to_reads + get_kmers on every subread (mod.rs:240-247 `kmers_via_reads`).
 * `out` needs len + 4 entries. */
int64_t orc_kmers_via_reads(const char *seq, size_t len, uint32_t k, uint64_t *out);

/* ---- counting.rs ------------------------------------------------------- */

typedef struct orc_counts orc_counts;
orc_counts *orc_counts_new(uint32_t k);
orc_counts *orc_counts_with_capacity(uint32_t k, uint64_t capacity);
void orc_counts_free(orc_counts *);
uint32_t orc_counts_k(const orc_counts *);
/* ingest_seq (counting.rs:144-149). */
int orc_counts_ingest_seq(orc_counts *, const char *seq, size_t len);
/* insert (counting.rs:152-154) = insert_or_add, saturating (counting.rs:82-85). */
void orc_counts_insert(orc_counts *, uint64_t kmer, uint32_t count);
/* insert_or_add_get_counts (counting.rs:86-92). */
void orc_counts_insert_get(orc_counts *, uint64_t kmer, uint32_t count, uint32_t *old_count,
                           uint32_t *new_count);
/* extend (counting.rs:157-166). */
int orc_counts_extend(orc_counts *, const orc_counts *other);
/* get (counting.rs:212-214): returns 1 and *count if present. */
int orc_counts_get(const orc_counts *, uint64_t kmer, uint32_t *count);
/* get_canonical_count (counting.rs:205-209). */
uint32_t orc_counts_get_canonical_count(const orc_counts *, uint64_t kmer);
/* get_canonical (counting.rs:218-222): probe kmer, else revcomp. */
int orc_counts_get_canonical(const orc_counts *, uint64_t kmer, uint32_t *count);
/* FilteredKmerCounts (counting.rs:316-350). */
int orc_filtered_get_canonical(const orc_counts *, uint32_t min_count, uint64_t kmer,
                               uint32_t *count);
uint32_t orc_filtered_get_canonical_count(const orc_counts *, uint32_t min_count, uint64_t kmer);
uint64_t orc_counts_len(const orc_counts *);            /* get_n_unique_kmers :258 */
uint64_t orc_counts_n_kmers(const orc_counts *);        /* get_n_kmers :254 */
uint32_t orc_counts_max_count(const orc_counts *);      /* :275 */
uint32_t orc_counts_median_count(const orc_counts *);   /* :279-300 */
void orc_counts_remove_low(orc_counts *, uint32_t min_count); /* :234-236 */
/* iter() exported in ascending k-mer order (the parity artefact). */
uint64_t orc_counts_export_sorted(const orc_counts *, uint64_t *keys, uint32_t *counts,
                                  uint64_t cap);
/* wrapping sum of skm_pair_digest over all entries. */
uint64_t orc_counts_digest(const orc_counts *);

/* find_oligos_in_kmers (src/pcr/primers.rs:163-226): the full-table scan sPCR runs per primer
 * direction and mismatch level.  `oligos` are unshifted 2-bit oligos of `oligo_length` bases
 * (0 < oligo_length < k).  A table k-mer with count >= min_count matches if its first
 * oligo_length bases equal an oligo (kept as is), else if its last oligo_length bases equal the
 * reverse complement of an oligo (then its reverse complement is reported).  Output sorted by
 * k-mer; returns the number of matches (keys/counts may be NULL to size). */
uint64_t orc_find_oligos(const orc_counts *, const uint64_t *oligos, uint64_t n_oligos,
                         uint32_t oligo_length, uint32_t min_count, uint64_t *keys, uint32_t *counts,
                         uint64_t cap);

/* ---- histogram.rs ------------------------------------------------------ */

typedef struct orc_histo orc_histo;
orc_histo *orc_histo_new(uint64_t histo_max);
void orc_histo_free(orc_histo *);
void orc_histo_move_count(orc_histo *, uint64_t old_count, uint64_t new_count); /* :51-85 */
void orc_histo_ingest(orc_histo *, const orc_counts *);                         /* :31-41 */
/* get_vector (:125-134): out has histo_max+2 entries. */
void orc_histo_get_vector(const orc_histo *, uint64_t *out);
uint64_t orc_histo_n_kmers(const orc_histo *);        /* :103-117 */
uint64_t orc_histo_n_unique(const orc_histo *);       /* :119-123 */
/* extend_with_histogram (counting.rs:171-202); *saturated set if any count hit MAX. */
int orc_counts_extend_with_histogram(orc_counts *, const orc_counts *other, orc_histo *,
                                     int *saturated);

/* ---- chunk.rs + io.rs -------------------------------------------------- */

typedef struct orc_run orc_run;
/* ingest_reads setup (io.rs:378-386): n_chunks = max(1, chunks). */
orc_run *orc_run_new(uint32_t k, uint32_t chunks, uint64_t histo_max);
void orc_run_free(orc_run *);
/* One sequence line as read_fastq sees it (io.rs:334-343): push, count,
 * drain every 1000th read.  Returns 0 or a negative error (invalid base). */
int orc_run_push_seq(orc_run *, const char *seq, size_t len);
/* Newline-terminated sequences, pushed one by one. */
int orc_run_push_lines(orc_run *, const char *buf, size_t n_bytes);
/* read_fastq over a whole file (plain or gzip), io.rs:271-352 + :598-625.
 * max_reads = 0 means unlimited.  Returns 1 if max_reads was reached, 0 at
 * EOF, <0 on error (message via orc_run_error). */
int orc_run_read_fastq(orc_run *, const char *path, uint64_t max_reads, uint64_t validate_every);
/* read_fastq_paired, io.rs:630-697 (max_reads is rounded up to even by the
 * caller as in io.rs:483-485). */
int orc_run_read_fastq_paired(orc_run *, const char *path1, const char *path2, uint64_t max_reads,
                              uint64_t validate_every);
/* Final drain + totals (io.rs:541-552, 578-580). */
int orc_run_finish_ingest(orc_run *);
/* consolidate_and_histogram (io.rs:977-1161) without the file writes. */
int orc_run_consolidate(orc_run *);
/* Writes {dir}{sample}.histo and .final.histo exactly as io.rs:1049-1094
 * (only when chunks > 0, as in the reference). */
int orc_run_write_histo(const orc_run *, const char *directory, const char *sample);
/* {dir}{sample}.stats.yaml with the RunStats scalar fields (stats.rs:26-45). */
int orc_run_write_stats(const orc_run *, const char *directory, const char *sample,
                        const char *command);

const char *orc_run_error(const orc_run *);
uint32_t orc_run_n_chunks(const orc_run *);
uint64_t orc_run_n_reads_read(const orc_run *);
uint64_t orc_run_n_bases_read(const orc_run *);
uint64_t orc_run_n_reads_ingested(const orc_run *);
uint64_t orc_run_n_bases_ingested(const orc_run *);
uint64_t orc_run_n_kmers_ingested(const orc_run *);
uint64_t orc_run_chunk_n_reads(const orc_run *, uint32_t chunk);
uint64_t orc_run_chunk_n_bases(const orc_run *, uint32_t chunk);
uint64_t orc_run_chunk_n_kmers(const orc_run *, uint32_t chunk);
/* Merged table (valid after consolidate). */
const orc_counts *orc_run_table(const orc_run *);
/* Histogram vector after merging chunks 0..chunk_i (valid when chunks > 0). */
int orc_run_histogram(const orc_run *, uint32_t chunk_i, uint64_t *out);
int orc_run_n_singletons(const orc_run *, uint64_t *out);

/* ---- synthetic reads (include/skm_common.h) ------------- */

/* Writes reads [first, first+n) as newline-terminated sequence lines.
 * `out` must hold n*(L+1) bytes. */
void orc_synth_reads(uint64_t seed, uint64_t genome_len, uint32_t read_len, uint32_t sub_thresh,
                     uint32_t n_thresh, uint64_t first, uint64_t n, char *out);
/* Same reads as a 4-line FASTQ file (headers @r{index}, qualities 'I'). */
int orc_synth_fastq(uint64_t seed, uint64_t genome_len, uint32_t read_len, uint32_t sub_thresh,
                    uint32_t n_thresh, uint64_t first, uint64_t n, const char *path, int gzip);

#ifdef __cplusplus
}
#endif
#endif
