/*
 * skm_oracle.c — CPU ORACLE (test infrastructure only; see skm_oracle.h).
 *
 * Each function cites the reference lines (caseywdunn/sharkmer v3.1.0) whose
 * behaviour it restates.  Written from the behaviour, not transliterated: the
 * reference is Rust over std::HashMap; this is C over a small open-addressing
 * map whose layout and hash are private and never observable in any output.
 */
#define _GNU_SOURCE
#include "skm_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <zlib.h>

#include "../include/skm_common.h"

#define ORC_VERSION "3.1.0" /* CARGO_PKG_VERSION, Cargo.toml:3; used in io.rs:1009-1014 */
#define ORC_READS_PER_BATCH 1000u /* io.rs:15 */

/* ======================================================================= */
/* encoding.rs                                                             */
/* ======================================================================= */

static inline int base_code(unsigned char b) {
    switch (b) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return -1;
    }
}

/* encoding.rs:332-371 */
int64_t orc_kmers_from_ascii(const char *seq, size_t len, uint32_t k, uint64_t *out) {
    if (k == 0 || k >= 32) return ORC_ERR_BAD_K;
    const uint64_t mask = (1ull << (2 * k)) - 1;
    uint64_t frame = 0, revframe = 0;
    size_t n_valid = 0;
    int64_t n = 0;
    for (size_t i = 0; i < len; i++) {
        unsigned char b = (unsigned char)seq[i];
        int code = base_code(b);
        if (code < 0) {
            if (b == 'N') { /* reset — a new subread starts after the N */
                n_valid = 0;
                frame = 0;
                revframe = 0;
                continue;
            }
            return ORC_ERR_INVALID_BASE;
        }
        uint64_t base = (uint64_t)code;
        frame = (frame << 2) | base;
        revframe = (revframe >> 2) | ((3 - base) << (2 * (k - 1)));
        n_valid++;
        if (n_valid >= k) {
            uint64_t f = frame & mask, r = revframe & mask;
            out[n++] = f < r ? f : r;
        }
    }
    return n;
}

/* encoding.rs:374-376 */
uint64_t orc_count_valid_bases(const char *seq, size_t len) {
    uint64_t n = 0;
    for (size_t i = 0; i < len; i++) n += (seq[i] != 'N');
    return n;
}

/* encoding.rs:219-262: byte-LUT walk, 4 bases at a time, then 1-3 leftovers. */
static uint8_t g_rc_lut[256];
static int g_rc_lut_ready = 0;
static void rc_lut_init(void) {
    for (unsigned i = 0; i < 256; i++) {
        unsigned b0 = i & 3, b1 = (i >> 2) & 3, b2 = (i >> 4) & 3, b3 = (i >> 6) & 3;
        g_rc_lut[i] = (uint8_t)(((3 - b0) << 6) | ((3 - b1) << 4) | ((3 - b2) << 2) | (3 - b3));
    }
    g_rc_lut_ready = 1;
}
uint64_t orc_revcomp_kmer(uint64_t kmer, uint32_t k) {
    if (!g_rc_lut_ready) rc_lut_init();
    uint64_t rc = 0;
    uint32_t remaining = k, shift = 0;
    while (remaining >= 4) {
        rc = (rc << 8) | g_rc_lut[(kmer >> shift) & 0xFF];
        shift += 8;
        remaining -= 4;
    }
    for (uint32_t i = 0; i < remaining; i++) {
        uint64_t base = (kmer >> (shift + 2 * i)) & 3;
        rc = (rc << 2) | (3 - base);
    }
    return 2 * k < 64 ? rc & ((1ull << (2 * k)) - 1) : rc;
}

/* encoding.rs:379-392 */
uint64_t orc_seq_to_kmer(const char *seq, size_t len, int *err) {
    uint64_t kmer = 0;
    if (err) *err = ORC_OK;
    for (size_t i = 0; i < len; i++) {
        int c = base_code((unsigned char)seq[i]);
        if (c < 0) {
            if (err) *err = ORC_ERR_INVALID_BASE;
            return 0;
        }
        kmer = (kmer << 2) | (uint64_t)c;
    }
    return kmer;
}

/* encoding.rs:311-325 */
void orc_kmer_to_seq(uint64_t kmer, uint32_t k, char *out) {
    for (uint32_t i = 0; i < k; i++) out[i] = "ACGT"[(kmer >> (2 * (k - i - 1))) & 3];
    out[k] = 0;
}

/* encoding.rs:60-95 (Read::from_str): MSB-first, tail left-aligned. */
int64_t orc_read_pack(const char *seq, size_t len, uint8_t *out) {
    int64_t nb = 0;
    uint8_t frame = 0;
    size_t length = 0;
    for (size_t i = 0; i < len; i++) {
        int c = base_code((unsigned char)seq[i]);
        if (c < 0) return ORC_ERR_INVALID_BASE;
        length++;
        frame = (uint8_t)((frame << 2) | c);
        if (length % 4 == 0) {
            out[nb++] = frame;
            frame = 0;
        }
    }
    if (length % 4) {
        frame = (uint8_t)(frame << (2 * (4 - length % 4)));
        out[nb++] = frame;
    }
    return nb;
}

/* encoding.rs:132-189 (Read::get_kmers): walks the packed bytes, emits a
 * k-mer per base once k bases are in the frame, then drops the k-mers that
 * came from the padding bases of the last byte. */
int64_t orc_read_get_kmers(const uint8_t *packed, size_t n_bytes, size_t length, uint32_t k,
                           uint64_t *out) {
    if (k == 0 || k >= 32) return ORC_ERR_BAD_K;
    if (length < k) return 0;
    const uint64_t mask = (1ull << (2 * k)) - 1;
    uint64_t frame = 0, revframe = 0;
    size_t n_valid = 0;
    int64_t n = 0;
    for (size_t i = 0; i < n_bytes; i++) {
        for (int j = 0; j < 4; j++) {
            uint64_t base = (packed[i] >> ((3 - j) * 2)) & 3;
            frame = (frame << 2) | base;
            revframe = (revframe >> 2) | ((3 - base) << (2 * (k - 1)));
            n_valid++;
            if (n_valid >= k) {
                uint64_t f = frame & mask, r = revframe & mask;
                out[n++] = f < r ? f : r;
            }
        }
    }
    if (length % 4) {
        int64_t extra = (int64_t)(4 - length % 4);
        n = n > extra ? n - extra : 0;
    }
    if ((size_t)n != length - k + 1) return ORC_ERR_CONSERVATION;
    return n;
}

/* encoding.rs:284-298 (seq_to_reads) + mod.rs:240-247 (kmers_via_reads). */
int64_t orc_kmers_via_reads(const char *seq, size_t len, uint32_t k, uint64_t *out) {
    int64_t total = 0;
    uint8_t *packed = (uint8_t *)malloc(len / 4 + 2);
    size_t i = 0;
    while (i <= len) {
        size_t j = i;
        while (j < len && seq[j] != 'N') j++;
        if (j > i) {
            int64_t nb = orc_read_pack(seq + i, j - i, packed);
            if (nb < 0) {
                free(packed);
                return nb;
            }
            int64_t n = orc_read_get_kmers(packed, (size_t)nb, j - i, k, out + total);
            if (n < 0) {
                free(packed);
                return n;
            }
            total += n;
        }
        i = j + 1;
    }
    free(packed);
    return total;
}

/* ======================================================================= */
/* a private u64 -> u64 open-addressing map                                */
/* ======================================================================= */

typedef struct {
    uint64_t key;
    uint64_t val;
} u64cell;

typedef struct {
    u64cell *cells; /* one cell = one cache-line access per probe */
    uint64_t cap;   /* power of two, or 0 */
    uint64_t len;
} u64map;

/* Large tables ask for transparent huge pages (the host's THP mode is usually
 * "madvise"): this only makes the CPU baseline faster than a stock allocator
 * would be, i.e. it errs in the reference's favour. */
static u64cell *cells_alloc(uint64_t cap) {
    u64cell *c;
    const size_t bytes = cap * sizeof(u64cell);
    if (bytes >= (4u << 20)) {
        c = (u64cell *)aligned_alloc(2u << 20, (bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1));
        if (c) madvise(c, bytes, MADV_HUGEPAGE);
    } else {
        c = (u64cell *)malloc(bytes);
    }
    for (uint64_t i = 0; i < cap; i++) {
        c[i].key = SKM_EMPTY_KEY;
        c[i].val = 0;
    }
    return c;
}
static void map_init(u64map *m, uint64_t want) {
    m->cells = NULL;
    m->cap = 0;
    m->len = 0;
    if (want) {
        uint64_t cap = 16;
        while (cap * 7 / 10 < want) cap <<= 1;
        m->cap = cap;
        m->cells = cells_alloc(cap);
    }
}
static void map_free(u64map *m) {
    free(m->cells);
    m->cells = NULL;
    m->cap = m->len = 0;
}
static inline uint64_t map_home(const u64map *m, uint64_t key) {
    /* deliberately not the device's slot function: the oracle's layout is private */
    return (skm_mix64(key ^ 0x5bd1e9955bd1e995ull)) & (m->cap - 1);
}
static void map_grow(u64map *m) {
    u64map n;
    uint64_t cap = m->cap ? m->cap * 2 : 16;
    n.cap = cap;
    n.len = m->len;
    n.cells = cells_alloc(cap);
    for (uint64_t i = 0; i < m->cap; i++) {
        if (m->cells[i].key == SKM_EMPTY_KEY) continue;
        uint64_t s = map_home(&n, m->cells[i].key);
        while (n.cells[s].key != SKM_EMPTY_KEY) s = (s + 1) & (cap - 1);
        n.cells[s] = m->cells[i];
    }
    free(m->cells);
    *m = n;
}
/* entry(key).or_insert(0): returns pointer to the value */
static inline uint64_t *map_entry(u64map *m, uint64_t key) {
    if (m->cap == 0 || (m->len + 1) * 10 > m->cap * 7) map_grow(m);
    uint64_t s = map_home(m, key);
    for (;;) {
        if (m->cells[s].key == key) return &m->cells[s].val;
        if (m->cells[s].key == SKM_EMPTY_KEY) {
            m->cells[s].key = key;
            m->cells[s].val = 0;
            m->len++;
            return &m->cells[s].val;
        }
        s = (s + 1) & (m->cap - 1);
    }
}
static inline const uint64_t *map_get(const u64map *m, uint64_t key) {
    if (m->cap == 0) return NULL;
    uint64_t s = map_home(m, key);
    for (;;) {
        if (m->cells[s].key == key) return &m->cells[s].val;
        if (m->cells[s].key == SKM_EMPTY_KEY) return NULL;
        s = (s + 1) & (m->cap - 1);
    }
}
static void map_remove(u64map *m, uint64_t key) {
    /* linear-probing backward-shift deletion */
    if (m->cap == 0) return;
    uint64_t mask = m->cap - 1, s = map_home(m, key);
    while (m->cells[s].key != key) {
        if (m->cells[s].key == SKM_EMPTY_KEY) return;
        s = (s + 1) & mask;
    }
    uint64_t hole = s;
    for (;;) {
        s = (s + 1) & mask;
        if (m->cells[s].key == SKM_EMPTY_KEY) break;
        uint64_t h = map_home(m, m->cells[s].key);
        /* can the entry at s move into the hole? yes unless h lies cyclically in (hole, s] */
        int in_range = hole <= s ? (h > hole && h <= s) : (h > hole || h <= s);
        if (!in_range) {
            m->cells[hole] = m->cells[s];
            hole = s;
        }
    }
    m->cells[hole].key = SKM_EMPTY_KEY;
    m->len--;
}

/* ======================================================================= */
/* counting.rs                                                             */
/* ======================================================================= */

struct orc_counts {
    u64map map; /* values are u32 counts held in u64 cells */
    uint32_t k;
    uint64_t *scratch; /* per-read k-mer vector (kmers_from_ascii's Vec) */
    size_t scratch_cap;
};

orc_counts *orc_counts_with_capacity(uint32_t k, uint64_t capacity) {
    orc_counts *c = (orc_counts *)calloc(1, sizeof(orc_counts));
    c->k = k;
    map_init(&c->map, capacity);
    return c;
}
orc_counts *orc_counts_new(uint32_t k) { return orc_counts_with_capacity(k, 0); }
void orc_counts_free(orc_counts *c) {
    if (!c) return;
    map_free(&c->map);
    free(c->scratch);
    free(c);
}
uint32_t orc_counts_k(const orc_counts *c) { return c->k; }

typedef struct {
    uint64_t key;
    uint32_t count;
} kc_pair;
static int cmp_pair(const void *a, const void *b) {
    uint64_t x = ((const kc_pair *)a)->key, y = ((const kc_pair *)b)->key;
    return x < y ? -1 : x > y;
}
static inline uint32_t sat_add_u32(uint32_t a, uint32_t b) {
    uint64_t s = (uint64_t)a + b;
    return s > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)s;
}

/* counting.rs:82-85 */
void orc_counts_insert(orc_counts *c, uint64_t kmer, uint32_t count) {
    uint64_t *v = map_entry(&c->map, kmer);
    *v = sat_add_u32((uint32_t)*v, count);
}
/* counting.rs:86-92 */
void orc_counts_insert_get(orc_counts *c, uint64_t kmer, uint32_t count, uint32_t *old_count,
                           uint32_t *new_count) {
    uint64_t *v = map_entry(&c->map, kmer);
    uint32_t o = (uint32_t)*v, n = sat_add_u32(o, count);
    *v = n;
    *old_count = o;
    *new_count = n;
}
/* counting.rs:144-149 — k-mers of the whole read are extracted first; an
 * invalid base therefore fails the read before any of its k-mers is counted. */
int orc_counts_ingest_seq(orc_counts *c, const char *seq, size_t len) {
    if (c->scratch_cap < len + 1) {
        c->scratch_cap = len + 64;
        c->scratch = (uint64_t *)realloc(c->scratch, c->scratch_cap * sizeof(uint64_t));
    }
    int64_t n = orc_kmers_from_ascii(seq, len, c->k, c->scratch);
    if (n < 0) return (int)n;
    for (int64_t i = 0; i < n; i++) orc_counts_insert(c, c->scratch[i], 1);
    return ORC_OK;
}
/* counting.rs:157-166 */
int orc_counts_extend(orc_counts *c, const orc_counts *o) {
    if (c->k != o->k) return ORC_ERR_K_MISMATCH;
    for (uint64_t i = 0; i < o->map.cap; i++)
        if (o->map.cells[i].key != SKM_EMPTY_KEY)
            orc_counts_insert(c, o->map.cells[i].key, (uint32_t)o->map.cells[i].val);
    return ORC_OK;
}
int orc_counts_get(const orc_counts *c, uint64_t kmer, uint32_t *count) {
    const uint64_t *v = map_get(&c->map, kmer);
    if (!v) return 0;
    if (count) *count = (uint32_t)*v;
    return 1;
}
/* counting.rs:205-209 */
uint32_t orc_counts_get_canonical_count(const orc_counts *c, uint64_t kmer) {
    uint64_t rc = orc_revcomp_kmer(kmer, c->k);
    uint64_t canon = kmer < rc ? kmer : rc;
    const uint64_t *v = map_get(&c->map, canon);
    return v ? (uint32_t)*v : 0;
}
/* counting.rs:218-222 */
int orc_counts_get_canonical(const orc_counts *c, uint64_t kmer, uint32_t *count) {
    if (orc_counts_get(c, kmer, count)) return 1;
    return orc_counts_get(c, orc_revcomp_kmer(kmer, c->k), count);
}
/* counting.rs:328-336 */
int orc_filtered_get_canonical(const orc_counts *c, uint32_t min_count, uint64_t kmer,
                               uint32_t *count) {
    uint32_t v;
    if (!orc_counts_get_canonical(c, kmer, &v) || v < min_count) return 0;
    if (count) *count = v;
    return 1;
}
/* counting.rs:339-342 */
uint32_t orc_filtered_get_canonical_count(const orc_counts *c, uint32_t min_count, uint64_t kmer) {
    uint32_t v = orc_counts_get_canonical_count(c, kmer);
    return v >= min_count ? v : 0;
}
uint64_t orc_counts_len(const orc_counts *c) { return c->map.len; }
uint64_t orc_counts_n_kmers(const orc_counts *c) {
    uint64_t s = 0;
    for (uint64_t i = 0; i < c->map.cap; i++)
        if (c->map.cells[i].key != SKM_EMPTY_KEY) s += c->map.cells[i].val;
    return s;
}
uint32_t orc_counts_max_count(const orc_counts *c) {
    uint32_t m = 0;
    for (uint64_t i = 0; i < c->map.cap; i++)
        if (c->map.cells[i].key != SKM_EMPTY_KEY && c->map.cells[i].val > m) m = (uint32_t)c->map.cells[i].val;
    return m;
}
static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : x > y;
}
/* counting.rs:279-300: odd -> middle; even -> lower/2 + upper/2 */
uint32_t orc_counts_median_count(const orc_counts *c) {
    uint64_t n = c->map.len;
    if (!n) return 0;
    uint32_t *v = (uint32_t *)malloc(n * sizeof(uint32_t));
    uint64_t j = 0;
    for (uint64_t i = 0; i < c->map.cap; i++)
        if (c->map.cells[i].key != SKM_EMPTY_KEY) v[j++] = (uint32_t)c->map.cells[i].val;
    qsort(v, n, sizeof(uint32_t), cmp_u32);
    uint32_t r = (n % 2) ? v[n / 2] : v[n / 2 - 1] / 2 + v[n / 2] / 2;
    free(v);
    return r;
}
/* counting.rs:234-236 (retain count >= min) */
void orc_counts_remove_low(orc_counts *c, uint32_t min_count) {
    uint64_t n = 0, cap = c->map.cap;
    uint64_t *dead = (uint64_t *)malloc((c->map.len + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < cap; i++)
        if (c->map.cells[i].key != SKM_EMPTY_KEY && c->map.cells[i].val < min_count) dead[n++] = c->map.cells[i].key;
    for (uint64_t i = 0; i < n; i++) map_remove(&c->map, dead[i]);
    free(dead);
}

uint64_t orc_counts_export_sorted(const orc_counts *c, uint64_t *keys, uint32_t *counts,
                                  uint64_t cap) {
    uint64_t n = c->map.len;
    if (!keys || cap < n) return n;
    kc_pair *p = (kc_pair *)malloc((n + 1) * sizeof(kc_pair));
    uint64_t j = 0;
    for (uint64_t i = 0; i < c->map.cap; i++)
        if (c->map.cells[i].key != SKM_EMPTY_KEY) {
            p[j].key = c->map.cells[i].key;
            p[j].count = (uint32_t)c->map.cells[i].val;
            j++;
        }
    qsort(p, n, sizeof(kc_pair), cmp_pair);
    for (uint64_t i = 0; i < n; i++) {
        keys[i] = p[i].key;
        if (counts) counts[i] = p[i].count;
    }
    free(p);
    return n;
}
uint64_t orc_counts_digest(const orc_counts *c) {
    uint64_t d = 0;
    for (uint64_t i = 0; i < c->map.cap; i++)
        if (c->map.cells[i].key != SKM_EMPTY_KEY) d += skm_pair_digest(c->map.cells[i].key, (uint32_t)c->map.cells[i].val);
    return d;
}

/* src/pcr/primers.rs:163-226 */
uint64_t orc_find_oligos(const orc_counts *c, const uint64_t *oligos, uint64_t n_oligos,
                         uint32_t oligo_length, uint32_t min_count, uint64_t *keys, uint32_t *counts,
                         uint64_t cap) {
    const uint32_t k = c->k;
    if (!n_oligos || oligo_length == 0 || oligo_length >= k) return 0;
    const uint64_t mask = ((1ull << (2 * oligo_length)) - 1) << (2 * k - 2 * oligo_length);
    const uint64_t rc_mask = (1ull << (2 * oligo_length)) - 1;
    u64map fwd, rc;
    map_init(&fwd, n_oligos);
    map_init(&rc, n_oligos);
    for (uint64_t i = 0; i < n_oligos; i++) {
        *map_entry(&fwd, oligos[i] << (2 * (k - oligo_length))) = 1;
        *map_entry(&rc, orc_revcomp_kmer(oligos[i], oligo_length)) = 1;
    }
    uint64_t n = 0, mcap = 1024;
    kc_pair *m = (kc_pair *)malloc(mcap * sizeof(kc_pair));
    for (uint64_t i = 0; i < c->map.cap; i++) {
        const uint64_t kmer = c->map.cells[i].key;
        if (kmer == SKM_EMPTY_KEY) continue;
        const uint32_t count = (uint32_t)c->map.cells[i].val;
        if (count < min_count) continue;
        uint64_t out;
        if (map_get(&fwd, kmer & mask))
            out = kmer;
        else if (map_get(&rc, kmer & rc_mask))
            out = orc_revcomp_kmer(kmer, k);
        else
            continue;
        if (n == mcap) {
            mcap *= 2;
            m = (kc_pair *)realloc(m, mcap * sizeof(kc_pair));
        }
        m[n].key = out;
        m[n].count = count;
        n++;
    }
    qsort(m, n, sizeof(kc_pair), cmp_pair);
    if (keys && counts && cap >= n)
        for (uint64_t i = 0; i < n; i++) {
            keys[i] = m[i].key;
            counts[i] = m[i].count;
        }
    free(m);
    map_free(&fwd);
    map_free(&rc);
    return n;
}

/* ======================================================================= */
/* histogram.rs                                                            */
/* ======================================================================= */

struct orc_histo {
    uint64_t *histo;  /* bins 0..=histo_max, plus one spare slot (len = histo_max + 2) */
    u64map large;     /* count -> number of k-mers, for counts > histo_max */
    uint64_t histo_max;
};

/* histogram.rs:19-28 */
orc_histo *orc_histo_new(uint64_t histo_max) {
    orc_histo *h = (orc_histo *)calloc(1, sizeof(orc_histo));
    h->histo_max = histo_max;
    h->histo = (uint64_t *)calloc(histo_max + 2, sizeof(uint64_t));
    map_init(&h->large, 0);
    return h;
}
void orc_histo_free(orc_histo *h) {
    if (!h) return;
    free(h->histo);
    map_free(&h->large);
    free(h);
}
/* histogram.rs:51-85 */
void orc_histo_move_count(orc_histo *h, uint64_t old_count, uint64_t new_count) {
    if (old_count == new_count) return;
    if (old_count > 0) {
        if (old_count <= h->histo_max) {
            if (h->histo[old_count] > 0) h->histo[old_count]--; /* saturating_sub */
        } else {
            const uint64_t *v = map_get(&h->large, old_count);
            if (v) {
                uint64_t *w = map_entry(&h->large, old_count);
                if (*w > 0) (*w)--;
                if (*w == 0) map_remove(&h->large, old_count);
            }
        }
    }
    if (new_count <= h->histo_max)
        h->histo[new_count]++;
    else
        (*map_entry(&h->large, new_count))++;
}
/* histogram.rs:31-41 */
void orc_histo_ingest(orc_histo *h, const orc_counts *c) {
    for (uint64_t i = 0; i < c->map.cap; i++) {
        if (c->map.cells[i].key == SKM_EMPTY_KEY) continue;
        uint64_t count = c->map.cells[i].val;
        if (count <= h->histo_max)
            h->histo[count]++;
        else
            (*map_entry(&h->large, count))++;
    }
}
/* histogram.rs:125-134: the spare last slot receives every count > histo_max */
void orc_histo_get_vector(const orc_histo *h, uint64_t *out) {
    memcpy(out, h->histo, (h->histo_max + 2) * sizeof(uint64_t));
    for (uint64_t i = 0; i < h->large.cap; i++)
        if (h->large.cells[i].key != SKM_EMPTY_KEY) out[h->histo_max + 1] += h->large.cells[i].val;
}
/* histogram.rs:103-117 */
uint64_t orc_histo_n_kmers(const orc_histo *h) {
    uint64_t s = 0;
    for (uint64_t i = 1; i < h->histo_max + 2; i++) s += h->histo[i] * i;
    for (uint64_t i = 0; i < h->large.cap; i++)
        if (h->large.cells[i].key != SKM_EMPTY_KEY) s += h->large.cells[i].key * h->large.cells[i].val;
    return s;
}
/* histogram.rs:119-123 */
uint64_t orc_histo_n_unique(const orc_histo *h) {
    uint64_t s = 0;
    for (uint64_t i = 1; i < h->histo_max + 2; i++) s += h->histo[i];
    for (uint64_t i = 0; i < h->large.cap; i++)
        if (h->large.cells[i].key != SKM_EMPTY_KEY) s += h->large.cells[i].val;
    return s;
}
/* counting.rs:171-202 */
int orc_counts_extend_with_histogram(orc_counts *c, const orc_counts *o, orc_histo *h,
                                     int *saturated) {
    if (c->k != o->k) return ORC_ERR_K_MISMATCH;
    int any = 0;
    for (uint64_t i = 0; i < o->map.cap; i++) {
        if (o->map.cells[i].key == SKM_EMPTY_KEY) continue;
        uint32_t oldc, newc;
        orc_counts_insert_get(c, o->map.cells[i].key, (uint32_t)o->map.cells[i].val, &oldc, &newc);
        orc_histo_move_count(h, oldc, newc);
        if (newc == 0xFFFFFFFFu && oldc < 0xFFFFFFFFu) any = 1;
    }
    if (saturated) *saturated = any;
    return ORC_OK;
}

/* ======================================================================= */
/* chunk.rs + io.rs                                                        */
/* ======================================================================= */

typedef struct {
    orc_counts *counts;
    uint64_t n_reads, n_bases;
} orc_chunk;

struct orc_run {
    uint32_t k, chunks_arg, n_chunks;
    uint64_t histo_max;
    orc_chunk *chunks;
    uint32_t chunk_index;
    /* state.seqs: the pending batch, as concatenated bytes + offsets */
    char *seq_buf;
    size_t seq_len, seq_cap;
    size_t *seq_off;
    size_t n_seqs, off_cap;
    uint64_t n_reads_read, n_bases_read;
    uint64_t n_reads_ingested, n_bases_ingested, n_kmers_ingested;
    uint64_t *chunk_n_kmers;
    orc_counts *table;
    uint64_t *histo_vecs; /* n_chunks x (histo_max+2) */
    int have_histo;
    uint64_t n_singletons;
    int saturated;
    char err[512];
};

/* io.rs:378-386 */
orc_run *orc_run_new(uint32_t k, uint32_t chunks, uint64_t histo_max) {
    orc_run *r = (orc_run *)calloc(1, sizeof(orc_run));
    r->k = k;
    r->chunks_arg = chunks;
    r->n_chunks = chunks == 0 ? 1 : chunks;
    r->histo_max = histo_max;
    r->chunks = (orc_chunk *)calloc(r->n_chunks, sizeof(orc_chunk));
    r->chunk_n_kmers = (uint64_t *)calloc(r->n_chunks, sizeof(uint64_t));
    for (uint32_t i = 0; i < r->n_chunks; i++) r->chunks[i].counts = orc_counts_new(k);
    return r;
}
void orc_run_free(orc_run *r) {
    if (!r) return;
    for (uint32_t i = 0; i < r->n_chunks; i++) orc_counts_free(r->chunks[i].counts);
    free(r->chunks);
    free(r->chunk_n_kmers);
    free(r->seq_buf);
    free(r->seq_off);
    orc_counts_free(r->table);
    free(r->histo_vecs);
    free(r);
}
const char *orc_run_error(const orc_run *r) { return r->err; }

/* chunk.rs:25-30 */
static int chunk_ingest_seq(orc_chunk *c, const char *seq, size_t len) {
    int rc = orc_counts_ingest_seq(c->counts, seq, len);
    if (rc) return rc;
    c->n_reads += 1;
    c->n_bases += orc_count_valid_bases(seq, len);
    return ORC_OK;
}

/* io.rs:355-361 */
static int drain_batch(orc_run *r) {
    for (size_t i = 0; i < r->n_seqs; i++) {
        const char *s = r->seq_buf + r->seq_off[i];
        size_t len = r->seq_off[i + 1] - r->seq_off[i];
        int rc = chunk_ingest_seq(&r->chunks[r->chunk_index], s, len);
        if (rc) {
            for (size_t j = 0; j < len; j++)
                if (base_code((unsigned char)s[j]) < 0 && s[j] != 'N') {
                    snprintf(r->err, sizeof r->err,
                             "Invalid character '%c' in sequence. Only ACGTN allowed.", s[j]);
                    break;
                }
            r->n_seqs = 0;
            r->seq_len = 0;
            return rc;
        }
    }
    r->n_seqs = 0;
    r->seq_len = 0;
    r->chunk_index = (r->chunk_index + 1) % r->n_chunks;
    return ORC_OK;
}

static void push_pending(orc_run *r, const char *seq, size_t len) {
    if (r->seq_len + len > r->seq_cap) {
        r->seq_cap = (r->seq_len + len) * 2 + 4096;
        r->seq_buf = (char *)realloc(r->seq_buf, r->seq_cap);
    }
    if (r->n_seqs + 2 > r->off_cap) {
        r->off_cap = r->off_cap * 2 + 1024;
        r->seq_off = (size_t *)realloc(r->seq_off, r->off_cap * sizeof(size_t));
    }
    memcpy(r->seq_buf + r->seq_len, seq, len);
    r->seq_off[r->n_seqs] = r->seq_len;
    r->seq_len += len;
    r->n_seqs++;
    r->seq_off[r->n_seqs] = r->seq_len;
}

/* io.rs:334-343 */
int orc_run_push_seq(orc_run *r, const char *seq, size_t len) {
    r->n_bases_read += len;
    push_pending(r, seq, len);
    r->n_reads_read += 1;
    if (r->n_reads_read % ORC_READS_PER_BATCH == 0) return drain_batch(r);
    return ORC_OK;
}

int orc_run_push_lines(orc_run *r, const char *buf, size_t n_bytes) {
    size_t i = 0;
    while (i < n_bytes) {
        const char *nl = (const char *)memchr(buf + i, '\n', n_bytes - i);
        size_t len = nl ? (size_t)(nl - (buf + i)) : n_bytes - i;
        int rc = orc_run_push_seq(r, buf + i, len);
        if (rc) return rc;
        i += len + 1;
    }
    return ORC_OK;
}

/* ---- line reader: BufRead::lines() over a plain or single-member gzip file
 * (io.rs:598-625: gzip if the name ends in .gz/.gzip or the magic is 1f 8b;
 * flate2's GzDecoder reads ONE member). */
typedef struct {
    FILE *f;
    int gz;
    z_stream zs;
    int z_end;
    unsigned char *in;  /* compressed input buffer */
    unsigned char *buf; /* decoded bytes */
    size_t pos, fill, cap;
    int eof;
    char *line;
    size_t line_cap;
} line_reader;

static int ends_with(const char *s, const char *suf) {
    size_t a = strlen(s), b = strlen(suf);
    return a >= b && strcmp(s + a - b, suf) == 0;
}

static int lr_open(line_reader *lr, const char *path) {
    memset(lr, 0, sizeof *lr);
    lr->f = fopen(path, "rb");
    if (!lr->f) return ORC_ERR_IO;
    lr->cap = 1 << 20;
    lr->buf = (unsigned char *)malloc(lr->cap);
    lr->in = (unsigned char *)malloc(lr->cap);
    int gz = ends_with(path, ".gz") || ends_with(path, ".gzip");
    if (!gz) {
        int c0 = fgetc(lr->f), c1 = fgetc(lr->f);
        gz = (c0 == 0x1f && c1 == 0x8b);
        rewind(lr->f);
    }
    lr->gz = gz;
    if (gz) {
        if (inflateInit2(&lr->zs, 15 + 16) != Z_OK) return ORC_ERR_IO;
    }
    return ORC_OK;
}
static void lr_close(line_reader *lr) {
    if (lr->gz) inflateEnd(&lr->zs);
    if (lr->f) fclose(lr->f);
    free(lr->buf);
    free(lr->in);
    free(lr->line);
}
/* refill decoded buffer; returns bytes available (0 at EOF), <0 on error */
static long lr_fill(line_reader *lr) {
    lr->pos = 0;
    lr->fill = 0;
    if (lr->eof) return 0;
    if (!lr->gz) {
        lr->fill = fread(lr->buf, 1, lr->cap, lr->f);
        if (lr->fill == 0) lr->eof = 1;
        return (long)lr->fill;
    }
    while (lr->fill == 0 && !lr->z_end) {
        if (lr->zs.avail_in == 0) {
            lr->zs.next_in = lr->in;
            lr->zs.avail_in = (uInt)fread(lr->in, 1, lr->cap, lr->f);
            if (lr->zs.avail_in == 0) { /* truncated stream */
                lr->eof = 1;
                return ORC_ERR_IO;
            }
        }
        lr->zs.next_out = lr->buf;
        lr->zs.avail_out = (uInt)lr->cap;
        int rc = inflate(&lr->zs, Z_NO_FLUSH);
        lr->fill = lr->cap - lr->zs.avail_out;
        if (rc == Z_STREAM_END)
            lr->z_end = 1;
        else if (rc != Z_OK && rc != Z_BUF_ERROR) {
            lr->eof = 1;
            return ORC_ERR_IO;
        }
    }
    if (lr->fill == 0) lr->eof = 1;
    return (long)lr->fill;
}
/* Returns 1 and (*line,*len) for the next line (without \n or \r\n), 0 at
 * EOF, <0 on I/O error. */
static int lr_next(line_reader *lr, const char **line, size_t *len) {
    size_t n = 0;
    int got_any = 0;
    for (;;) {
        if (lr->pos == lr->fill) {
            long rc = lr_fill(lr);
            if (rc < 0) return (int)rc;
            if (rc == 0) break;
        }
        unsigned char *start = lr->buf + lr->pos;
        size_t avail = lr->fill - lr->pos;
        unsigned char *nl = (unsigned char *)memchr(start, '\n', avail);
        size_t take = nl ? (size_t)(nl - start) : avail;
        if (n + take + 1 > lr->line_cap) {
            lr->line_cap = (n + take + 1) * 2 + 256;
            lr->line = (char *)realloc(lr->line, lr->line_cap);
        }
        memcpy(lr->line + n, start, take);
        n += take;
        got_any = 1;
        if (nl) {
            lr->pos += take + 1;
            if (n > 0 && lr->line[n - 1] == '\r') n--;
            *line = lr->line;
            *len = n;
            return 1;
        }
        lr->pos += take;
    }
    if (!got_any) return 0;
    /* final line without a trailing newline (Rust's lines() keeps a \r here) */
    *line = lr->line;
    *len = n;
    return 1;
}

/* io.rs:161-198 */
static int validate_record(orc_run *r, const char *hdr, size_t hl, const char *sep, size_t sl,
                           size_t qual_len, size_t seq_len, uint64_t record_num) {
    if (hl > 0 && hdr[0] == '>') {
        snprintf(r->err, sizeof r->err,
                 "Input appears to be FASTA format, not FASTQ (record %llu starts with '>'). "
                 "sharkmer requires FASTQ input with quality scores.",
                 (unsigned long long)(record_num + 1));
        return ORC_ERR_FASTQ;
    }
    if (!(hl > 0 && hdr[0] == '@')) {
        snprintf(r->err, sizeof r->err,
                 "FASTQ record %llu has invalid header (expected '@', got '%c'): %.*s",
                 (unsigned long long)(record_num + 1), hl ? hdr[0] : ' ', (int)hl, hdr);
        return ORC_ERR_FASTQ;
    }
    if (!(sl > 0 && sep[0] == '+')) {
        snprintf(r->err, sizeof r->err,
                 "FASTQ record %llu has invalid separator line (expected '+', got '%c'): %.*s",
                 (unsigned long long)(record_num + 1), sl ? sep[0] : ' ', (int)sl, sep);
        return ORC_ERR_FASTQ;
    }
    if (qual_len != seq_len) {
        snprintf(r->err, sizeof r->err,
                 "FASTQ record %llu has mismatched sequence (%zu) and quality (%zu) lengths",
                 (unsigned long long)(record_num + 1), seq_len, qual_len);
        return ORC_ERR_FASTQ;
    }
    return ORC_OK;
}

/* io.rs:701-765 (read_one_fastq_record); also the body of read_fastq's loop
 * (io.rs:282-337).  Returns 1 at EOF, 0 after one record, <0 on error. */
static int read_one_record(orc_run *r, line_reader *lr, uint64_t validate_every, const char *name) {
    const char *l;
    size_t n;
    int rc = lr_next(lr, &l, &n);
    if (rc < 0) {
        snprintf(r->err, sizeof r->err, "Failed to read header line of record %llu in %s",
                 (unsigned long long)(r->n_reads_read + 1), name);
        return ORC_ERR_IO;
    }
    if (rc == 0) return 1;
    char *hdr = (char *)malloc(n + 1);
    memcpy(hdr, l, n);
    size_t hl = n;
    static const char *roles[3] = {"sequence", "separator", "quality"};
    char *parts[3] = {NULL, NULL, NULL};
    size_t lens[3] = {0, 0, 0};
    int result = 0;
    for (int i = 0; i < 3; i++) {
        rc = lr_next(lr, &l, &n);
        if (rc <= 0) {
            if (rc == 0)
                snprintf(r->err, sizeof r->err,
                         "Truncated FASTQ record at record %llu in %s: missing %s line",
                         (unsigned long long)(r->n_reads_read + 1), name, roles[i]);
            else
                snprintf(r->err, sizeof r->err, "Failed to read %s line of record %llu in %s",
                         roles[i], (unsigned long long)(r->n_reads_read + 1), name);
            result = rc == 0 ? ORC_ERR_FASTQ : ORC_ERR_IO;
            goto done;
        }
        parts[i] = (char *)malloc(n + 1);
        memcpy(parts[i], l, n);
        lens[i] = n;
    }
    {
        int should_validate =
            r->n_reads_read == 0 || (validate_every > 0 && r->n_reads_read % validate_every == 0);
        if (should_validate) {
            result = validate_record(r, hdr, hl, parts[1], lens[1], lens[2], lens[0], r->n_reads_read);
            if (result) goto done;
        }
        r->n_bases_read += lens[0];
        push_pending(r, parts[0], lens[0]);
        r->n_reads_read += 1;
    }
done:
    free(hdr);
    for (int i = 0; i < 3; i++) free(parts[i]);
    return result;
}

/* io.rs:271-352 */
int orc_run_read_fastq(orc_run *r, const char *path, uint64_t max_reads, uint64_t validate_every) {
    line_reader lr;
    if (lr_open(&lr, path)) {
        snprintf(r->err, sizeof r->err, "Failed to open file: %s", path);
        lr_close(&lr);
        return ORC_ERR_IO;
    }
    int result = 0;
    for (;;) {
        int rc = read_one_record(r, &lr, validate_every, path);
        if (rc < 0) {
            result = rc;
            break;
        }
        if (rc == 1) break;
        if (r->n_reads_read % ORC_READS_PER_BATCH == 0) {
            rc = drain_batch(r);
            if (rc) {
                result = rc;
                break;
            }
        }
        if (max_reads > 0 && r->n_reads_read >= max_reads) {
            result = 1;
            break;
        }
    }
    lr_close(&lr);
    return result;
}

/* io.rs:630-697.  Quirk kept: when R1 ends first, ONE further R2 record is
 * read — and ingested — by the EOF probe (io.rs:653-657). */
int orc_run_read_fastq_paired(orc_run *r, const char *path1, const char *path2, uint64_t max_reads,
                              uint64_t validate_every) {
    line_reader l1, l2;
    if (lr_open(&l1, path1)) {
        snprintf(r->err, sizeof r->err, "Failed to open file: %s", path1);
        lr_close(&l1);
        return ORC_ERR_IO;
    }
    if (lr_open(&l2, path2)) {
        snprintf(r->err, sizeof r->err, "Failed to open file: %s", path2);
        lr_close(&l1);
        lr_close(&l2);
        return ORC_ERR_IO;
    }
    int result = 0;
    for (;;) {
        int rc = read_one_record(r, &l1, validate_every, path1);
        if (rc < 0) {
            result = rc;
            break;
        }
        if (rc == 1) {
            rc = read_one_record(r, &l2, validate_every, path2);
            if (rc < 0) result = rc;
            break;
        }
        if (r->n_reads_read % ORC_READS_PER_BATCH == 0 && (rc = drain_batch(r))) {
            result = rc;
            break;
        }
        if (max_reads > 0 && r->n_reads_read >= max_reads) {
            result = 1;
            break;
        }
        rc = read_one_record(r, &l2, validate_every, path2);
        if (rc < 0) {
            result = rc;
            break;
        }
        if (rc == 1) break;
        if (r->n_reads_read % ORC_READS_PER_BATCH == 0 && (rc = drain_batch(r))) {
            result = rc;
            break;
        }
        if (max_reads > 0 && r->n_reads_read >= max_reads) {
            result = 1;
            break;
        }
    }
    lr_close(&l1);
    lr_close(&l2);
    return result;
}

/* io.rs:541-552, 578-580 */
int orc_run_finish_ingest(orc_run *r) {
    int rc = drain_batch(r); /* the partial batch goes to the NEXT chunk index */
    if (rc) return rc;
    r->n_reads_ingested = r->n_bases_ingested = r->n_kmers_ingested = 0;
    for (uint32_t i = 0; i < r->n_chunks; i++) {
        r->chunk_n_kmers[i] = orc_counts_n_kmers(r->chunks[i].counts);
        r->n_reads_ingested += r->chunks[i].n_reads;
        r->n_bases_ingested += r->chunks[i].n_bases;
        r->n_kmers_ingested += r->chunk_n_kmers[i];
    }
    if (r->n_reads_ingested == 0) {
        snprintf(r->err, sizeof r->err,
                 "No reads were ingested. Check that input files contain valid FASTQ records.");
        return ORC_ERR_NO_READS;
    }
    return ORC_OK;
}

/* io.rs:1005-1047, 1096-1157 */
int orc_run_consolidate(orc_run *r) {
    uint64_t est = 0;
    for (uint32_t i = 0; i < r->n_chunks; i++) est += orc_counts_len(r->chunks[i].counts);
    r->table = orc_counts_with_capacity(r->k, est);
    if (r->chunks_arg > 0) {
        const uint64_t w = r->histo_max + 2;
        r->histo_vecs = (uint64_t *)calloc((size_t)r->n_chunks * w, sizeof(uint64_t));
        orc_histo *running = orc_histo_new(r->histo_max);
        for (uint32_t i = 0; i < r->n_chunks; i++) {
            int sat = 0;
            orc_counts_extend_with_histogram(r->table, r->chunks[i].counts, running, &sat);
            r->saturated |= sat;
            orc_counts_free(r->chunks[i].counts); /* drop(chunk) */
            r->chunks[i].counts = orc_counts_new(r->k);
            orc_histo_get_vector(running, r->histo_vecs + (size_t)i * w);
        }
        r->have_histo = 1;
        uint64_t n_hashed = orc_counts_n_kmers(r->table);
        if (n_hashed != r->n_kmers_ingested) {
            snprintf(r->err, sizeof r->err,
                     "The total count of hashed kmers (%llu) does not equal the number of ingested kmers (%llu)",
                     (unsigned long long)n_hashed, (unsigned long long)r->n_kmers_ingested);
            orc_histo_free(running);
            return ORC_ERR_CONSERVATION;
        }
        r->n_singletons = r->histo_vecs[(size_t)(r->n_chunks - 1) * w + 1];
        uint64_t hu = orc_histo_n_unique(running), hk = orc_histo_n_kmers(running);
        orc_histo_free(running);
        if (hk != r->n_kmers_ingested) {
            snprintf(r->err, sizeof r->err,
                     "The total count of kmers in the histogram (%llu) does not equal the total expected count of kmers (%llu)",
                     (unsigned long long)hk, (unsigned long long)r->n_kmers_ingested);
            return ORC_ERR_CONSERVATION;
        }
        if (hu != orc_counts_len(r->table)) {
            snprintf(r->err, sizeof r->err,
                     "The total count of unique kmers in the histogram (%llu) does not equal the total count of hashed kmers (%llu)",
                     (unsigned long long)hu, (unsigned long long)orc_counts_len(r->table));
            return ORC_ERR_CONSERVATION;
        }
    } else {
        for (uint32_t i = 0; i < r->n_chunks; i++) {
            orc_counts_extend(r->table, r->chunks[i].counts);
            orc_counts_free(r->chunks[i].counts);
            r->chunks[i].counts = orc_counts_new(r->k);
        }
        uint64_t n_hashed = orc_counts_n_kmers(r->table);
        if (n_hashed != r->n_kmers_ingested) {
            snprintf(r->err, sizeof r->err,
                     "The total count of hashed kmers (%llu) does not equal the number of ingested kmers (%llu)",
                     (unsigned long long)n_hashed, (unsigned long long)r->n_kmers_ingested);
            return ORC_ERR_CONSERVATION;
        }
    }
    return ORC_OK;
}

/* io.rs:1049-1094 */
int orc_run_write_histo(const orc_run *r, const char *directory, const char *sample) {
    if (!r->have_histo) return ORC_OK;
    const uint64_t w = r->histo_max + 2;
    char path[4096];
    snprintf(path, sizeof path, "%s%s.histo", directory, sample);
    FILE *f = fopen(path, "w");
    if (!f) return ORC_ERR_IO;
    fprintf(f, "# sharkmer %s k=%u chunks=%u\n", ORC_VERSION, r->k, r->chunks_arg);
    fprintf(f, "count");
    for (uint32_t c = 1; c <= r->n_chunks; c++) fprintf(f, "\tchunk_%u", c);
    fprintf(f, "\n");
    for (uint64_t i = 1; i < r->histo_max + 2; i++) {
        fprintf(f, "%llu", (unsigned long long)i);
        for (uint32_t c = 0; c < r->n_chunks; c++)
            fprintf(f, "\t%llu", (unsigned long long)r->histo_vecs[(size_t)c * w + i]);
        fprintf(f, "\n");
    }
    fclose(f);
    snprintf(path, sizeof path, "%s%s.final.histo", directory, sample);
    f = fopen(path, "w");
    if (!f) return ORC_ERR_IO;
    fprintf(f, "# sharkmer %s k=%u chunks=%u\n", ORC_VERSION, r->k, r->chunks_arg);
    fprintf(f, "count\tfrequency\n");
    const uint64_t *last = r->histo_vecs + (size_t)(r->n_chunks - 1) * w;
    for (uint64_t i = 1; i < r->histo_max + 2; i++)
        fprintf(f, "%llu\t%llu\n", (unsigned long long)i, (unsigned long long)last[i]);
    fclose(f);
    return ORC_OK;
}

/* main.rs:182-197 + stats.rs:26-45 (scalar fields; serde_yaml plain scalars) */
int orc_run_write_stats(const orc_run *r, const char *directory, const char *sample,
                        const char *command) {
    char path[4096];
    snprintf(path, sizeof path, "%s%s.stats.yaml", directory, sample);
    FILE *f = fopen(path, "w");
    if (!f) return ORC_ERR_IO;
    fprintf(f, "sharkmer_version: %s\n", ORC_VERSION);
    fprintf(f, "command: %s\n", command);
    fprintf(f, "sample: %s\n", sample);
    fprintf(f, "kmer_length: %u\n", r->k);
    fprintf(f, "chunks: %u\n", r->chunks_arg);
    fprintf(f, "n_reads_read: %llu\n", (unsigned long long)r->n_reads_read);
    fprintf(f, "n_bases_read: %llu\n", (unsigned long long)r->n_bases_read);
    fprintf(f, "n_subreads_ingested: %llu\n", (unsigned long long)r->n_reads_ingested);
    fprintf(f, "n_bases_ingested: %llu\n", (unsigned long long)r->n_bases_ingested);
    fprintf(f, "n_kmers: %llu\n", (unsigned long long)r->n_kmers_ingested);
    if (r->have_histo) {
        /* main.rs:193: n_kmers (occurrences) minus singleton DISTINCT count, saturating */
        uint64_t multi = r->n_kmers_ingested > r->n_singletons ? r->n_kmers_ingested - r->n_singletons : 0;
        fprintf(f, "n_multi_kmers: %llu\n", (unsigned long long)multi);
        fprintf(f, "n_singleton_kmers: %llu\n", (unsigned long long)r->n_singletons);
    }
    fprintf(f, "peak_memory_bytes: 0\n");
    fclose(f);
    return ORC_OK;
}

uint32_t orc_run_n_chunks(const orc_run *r) { return r->n_chunks; }
uint64_t orc_run_n_reads_read(const orc_run *r) { return r->n_reads_read; }
uint64_t orc_run_n_bases_read(const orc_run *r) { return r->n_bases_read; }
uint64_t orc_run_n_reads_ingested(const orc_run *r) { return r->n_reads_ingested; }
uint64_t orc_run_n_bases_ingested(const orc_run *r) { return r->n_bases_ingested; }
uint64_t orc_run_n_kmers_ingested(const orc_run *r) { return r->n_kmers_ingested; }
uint64_t orc_run_chunk_n_reads(const orc_run *r, uint32_t c) { return r->chunks[c].n_reads; }
uint64_t orc_run_chunk_n_bases(const orc_run *r, uint32_t c) { return r->chunks[c].n_bases; }
uint64_t orc_run_chunk_n_kmers(const orc_run *r, uint32_t c) { return r->chunk_n_kmers[c]; }
const orc_counts *orc_run_table(const orc_run *r) { return r->table; }
int orc_run_histogram(const orc_run *r, uint32_t chunk_i, uint64_t *out) {
    if (!r->have_histo || chunk_i >= r->n_chunks) return ORC_ERR_IO;
    memcpy(out, r->histo_vecs + (size_t)chunk_i * (r->histo_max + 2),
           (r->histo_max + 2) * sizeof(uint64_t));
    return ORC_OK;
}
int orc_run_n_singletons(const orc_run *r, uint64_t *out) {
    if (!r->have_histo) return ORC_ERR_IO;
    *out = r->n_singletons;
    return ORC_OK;
}

/* ======================================================================= */
/* synthetic reads                                                         */
/* ======================================================================= */

void orc_synth_reads(uint64_t seed, uint64_t genome_len, uint32_t read_len, uint32_t sub_thresh,
                     uint32_t n_thresh, uint64_t first, uint64_t n, char *out) {
    skm_synth_params p = {seed, genome_len, read_len, sub_thresh, n_thresh, 0};
    for (uint64_t i = 0; i < n; i++) {
        char *dst = out + i * (read_len + 1ull);
        for (uint32_t j = 0; j < read_len; j++) dst[j] = (char)skm_synth_read_base(&p, first + i, j);
        dst[read_len] = '\n';
    }
}

int orc_synth_fastq(uint64_t seed, uint64_t genome_len, uint32_t read_len, uint32_t sub_thresh,
                    uint32_t n_thresh, uint64_t first, uint64_t n, const char *path, int gzip) {
    skm_synth_params p = {seed, genome_len, read_len, sub_thresh, n_thresh, 0};
    char *rec = (char *)malloc(2ull * read_len + 64);
    gzFile gz = NULL;
    FILE *f = NULL;
    if (gzip)
        gz = gzopen(path, "wb1");
    else
        f = fopen(path, "wb");
    if (!gz && !f) {
        free(rec);
        return ORC_ERR_IO;
    }
    for (uint64_t i = 0; i < n; i++) {
        int o = sprintf(rec, "@r%llu\n", (unsigned long long)(first + i));
        for (uint32_t j = 0; j < read_len; j++) rec[o++] = (char)skm_synth_read_base(&p, first + i, j);
        rec[o++] = '\n';
        rec[o++] = '+';
        rec[o++] = '\n';
        memset(rec + o, 'I', read_len);
        o += (int)read_len;
        rec[o++] = '\n';
        if (gz)
            gzwrite(gz, rec, (unsigned)o);
        else
            fwrite(rec, 1, (size_t)o, f);
    }
    if (gz) gzclose(gz);
    if (f) fclose(f);
    free(rec);
    return ORC_OK;
}
