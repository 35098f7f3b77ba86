/*
 * skm_common.h — arithmetic shared bit-for-bit by the device kernels, the host
 * side of the C ABI, the multi-GPU router and the CPU oracle.
 *
 * Plain C99 / C++ / CUDA (every function is `static inline` and, under nvcc,
 * `__host__ __device__`).  Nothing here comes from the reference: sharkmer's
 * hash is `ahash::RandomState::new()` (src/kmer/counting.rs:72-81), randomly
 * keyed per process, so hash values are unobservable there.  "Bit-identical
 * hash" therefore means ONE function used by every component of this repo.
 */
#ifndef SKM_COMMON_H
#define SKM_COMMON_H

#include <stdint.h>

#ifdef __CUDACC__
#define SKM_HD __host__ __device__ __forceinline__
#else
#define SKM_HD static inline
#endif

/* Empty-slot sentinel.  k <= 31 => every key has its top two bits clear
 * (src/cli.rs:662-667 bounds k; src/kmer/encoding.rs:334 builds the mask). */
#define SKM_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

/* murmur3 fmix64: a bijection on u64 with full avalanche. */
SKM_HD uint64_t skm_mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

/* The table hash. */
SKM_HD uint64_t skm_hash_kmer(uint64_t kmer) { return skm_mix64(kmer); }

SKM_HD uint64_t skm_mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

/* Multi-GPU: the hash space [0, 2^64) is cut into n_ranks contiguous ranges.
 *   owner rank  = floor(h * n_ranks / 2^64)
 *   local hash  = position inside the owner's range rescaled to 64 bits
 *               = low 64 bits of h * n_ranks
 * so a list ordered by hash is ordered by (owner, local hash): one bucketing
 * pass on the sender yields, for every owner, runs already sorted by that
 * owner's table regions.  With n_ranks == 1: owner 0, local hash == h. */
SKM_HD uint32_t skm_owner_rank(uint64_t h, uint32_t n_ranks) {
    return (uint32_t)skm_mulhi64(h, (uint64_t)n_ranks);
}
SKM_HD uint64_t skm_local_hash(uint64_t h, uint32_t n_ranks) { return h * (uint64_t)n_ranks; }

/* Home slot in a table of 2^log2_capacity slots = top bits of the local hash. */
SKM_HD uint64_t skm_home_slot(uint64_t local_hash, uint32_t log2_capacity) {
    return log2_capacity ? (local_hash >> (64u - log2_capacity)) : 0ull;
}

/* Order-independent digest of one (kmer, count) pair; a table digest is the
 * wrapping u64 sum of these over all entries. */
SKM_HD uint64_t skm_pair_digest(uint64_t kmer, uint32_t count) {
    return skm_mix64(kmer ^ skm_mix64(0x9e3779b97f4a7c15ull + (uint64_t)count));
}

/* Reverse complement of a k-mer (2 bits/base, A=0 C=1 G=2 T=3, first base in
 * the most significant position).  Same result as revcomp_kmer,
 * src/kmer/encoding.rs:235-262 (which walks a byte LUT); done here with
 * bit-parallel swaps so the device can use it too. */
SKM_HD uint64_t skm_revcomp_kmer(uint64_t kmer, uint32_t k) {
    uint64_t x = ~kmer;
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
    x = (x >> 32) | (x << 32);
    return k >= 32 ? x : (x >> (64u - 2u * k));
}

/* ------------------------------------------------------------------------
 * Synthetic reads.  Counter-based: every base of every read
 * is a pure function of (seed, read index, position), so the device generator
 * and the host generator are bit-identical and any read can be regenerated
 * on its own.
 *
 *   genome base at p      = top 2 bits of rng(seed, 0, p)       (never stored)
 *   read r                : start = rng(seed,1,r) % (G-L+1), strand = bit 63 of rng(seed,2,r)
 *   base j of read r      : genome base (or complement of the mirrored one),
 *                           substituted with prob. sub_rate (uniform over the
 *                           3 other bases), then replaced by 'N' with prob. n_rate.
 * Rates are given as 32-bit thresholds: threshold = round(rate * 2^32).
 * ------------------------------------------------------------------------ */
typedef struct {
    uint64_t seed;
    uint64_t genome_len;   /* G */
    uint32_t read_len;     /* L */
    uint32_t sub_thresh;   /* substitution rate * 2^32 */
    uint32_t n_thresh;     /* N rate * 2^32 */
    uint32_t reserved;
} skm_synth_params;

SKM_HD uint64_t skm_rng(uint64_t seed, uint64_t stream, uint64_t ctr) {
    return skm_mix64(skm_mix64(seed + 0x9e3779b97f4a7c15ull * (stream + 1)) ^
                     (ctr * 0xd1342543de82ef95ull + 0x2545f4914f6cdd1dull));
}

SKM_HD uint32_t skm_synth_genome_base(const skm_synth_params *p, uint64_t pos) {
    return (uint32_t)(skm_rng(p->seed, 0, pos) >> 62);
}

/* ASCII base j (0 <= j < L) of read r. */
SKM_HD uint8_t skm_synth_read_base(const skm_synth_params *p, uint64_t r, uint32_t j) {
    const uint64_t span = p->genome_len - p->read_len + 1;
    const uint64_t start = skm_rng(p->seed, 1, r) % span;
    const uint32_t rev = (uint32_t)(skm_rng(p->seed, 2, r) >> 63);
    uint32_t b = rev ? 3u - skm_synth_genome_base(p, start + (p->read_len - 1 - j))
                     : skm_synth_genome_base(p, start + j);
    const uint64_t u = skm_rng(p->seed, 3, r * (uint64_t)p->read_len + j);
    if ((uint32_t)(u >> 32) < p->sub_thresh) b = (b + 1u + (uint32_t)((u >> 8) % 3u)) & 3u;
    const uint64_t v = skm_rng(p->seed, 4, r * (uint64_t)p->read_len + j);
    if ((uint32_t)(v >> 32) < p->n_thresh) return (uint8_t)'N';
    return (uint8_t)("ACGT"[b]);
}

/* Reads are striped over chunks in 1000-read batches (drain_batch,
 * src/io.rs:355-361; N_READS_PER_BATCH src/io.rs:15): the i-th read of chunk
 * c is global read skm_chunk_read_to_global(i, c, n_chunks). */
#define SKM_READS_PER_BATCH 1000ull
SKM_HD uint64_t skm_chunk_read_to_global(uint64_t i, uint32_t c, uint32_t n_chunks) {
    return ((i / SKM_READS_PER_BATCH) * n_chunks + c) * SKM_READS_PER_BATCH +
           (i % SKM_READS_PER_BATCH);
}

#endif /* SKM_COMMON_H */
