// skm_kernels.cuh — device code of the B200 k-mer counting engine (sm_100a).
//
// Data layout in HBM
//   staged reads : codes[u]  u64, 32 bases per word, 2 bits/base, first base in the
//                  most significant bits (the bit order of Read::from_str,
//                  src/kmer/encoding.rs:60-95), A=0 C=1 G=2 T=3;
//                  breaks[u] u32, 1 bit/base, MSB-first, 1 = no base here
//                  ('N', the '\n' that ends a read, or padding past the end).
//   count table  : open addressing, linear probing, 16-byte slots
//                  { u64 key ; u64 count }, EMPTY key = all ones.  Two slots per
//                  32-byte DRAM sector, so one probe = one sector.  Counts are
//                  held in 64 bits on the device and reported as
//                  min(count, u32::MAX): identical to the reference's chain of
//                  u32 saturating adds (src/kmer/counting.rs:82-92,171-202) because
//                  saturating addition of non-negative terms equals the clamped sum.
//
// All kernels are integer / byte work bounded by HBM (streaming or random
// sector access); no tensor cores are involved.
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/skm_common.h"

namespace skm {

struct __align__(16) Slot {
    unsigned long long key;
    unsigned long long count;
};

// The table is cut into partitions of 2^kPartLog2 consecutive slots (64 KiB).  A probe sequence
// never leaves the partition of its home slot (it wraps inside it), so a partition is a
// self-contained open-addressing table: the bulk insert kernel (tile_insert_kernel) loads one into
// shared memory, counts every k-mer that hashes there with shared-memory atomics, and writes it
// back — the table streams through the SMs once per launch instead of taking one random DRAM
// sector per k-mer.  Every other kernel (direct insert, lookup, rehash) follows the same rule.
// Capacity is always >= 2^16 slots, i.e. a whole number of partitions.
static constexpr uint32_t kPartLog2 = 12;
static constexpr uint32_t kPartSlots = 1u << kPartLog2;

// A rank's table: 2^log2cap slots addressed by the rank-local hash.
struct TableRef {
    Slot *slots;
    uint32_t log2cap;
    uint32_t n_ranks;
    __host__ __device__ __forceinline__ uint64_t mask() const { return (1ull << log2cap) - 1; }
    __host__ __device__ __forceinline__ uint64_t local_hash(uint64_t kmer) const {
        const uint64_t h = skm_hash_kmer(kmer);
        return n_ranks == 1 ? h : skm_local_hash(h, n_ranks);
    }
    __host__ __device__ __forceinline__ uint64_t home(uint64_t kmer) const {
        return skm_home_slot(local_hash(kmer), log2cap);
    }
    // next slot of a probe sequence: wraps inside the partition
    __host__ __device__ __forceinline__ static uint64_t next(uint64_t s) {
        return (s & ~(uint64_t)(kPartSlots - 1)) | ((s + 1) & (uint64_t)(kPartSlots - 1));
    }
};

struct ChunkCounters {          // one per chunk, device memory
    unsigned long long n_reads;   // '\n' seen by the pack kernel
    unsigned long long n_bases;   // A/C/G/T seen by the pack kernel
    unsigned long long n_windows; // valid k-mer windows seen by the extract kernels
    unsigned long long pad;
};

struct GlobalCounters {         // one per ctx, device memory
    unsigned long long first_bad;   // min over (byte position << 8 | byte) of invalid bytes; ~0 = none
    unsigned long long n_distinct;  // keys claimed so far
    unsigned long long scratch[2];  // tile counters of the persistent kernels
    unsigned long long part_full;   // != 0: a probe sequence found its whole partition occupied by other keys
    unsigned long long n_failed;    // tile_insert_kernel: partitions that ran out of room in this launch
    unsigned long long fatal;       // tile_insert_kernel: a failed partition had already published histogram moves
    unsigned long long pad;
};

struct HistoTotals {
    unsigned long long n_distinct;
    unsigned long long n_kmers;      // sum of min(count, u32::MAX)
    unsigned long long n_saturated;  // slots with count >= u32::MAX
    unsigned long long digest;       // wrapping sum of skm_pair_digest
};

static constexpr uint32_t kU32Max = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------

__device__ __forceinline__ unsigned long long ld_cg_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void red_add_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint4 ld_nc_v4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// 8-byte asynchronous global -> shared copy (LDGSTS): the data never passes through registers
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Shared-memory accesses by 32-bit shared-window address (the generic-pointer forms re-derive the
// window base, S2R + LEA, at every use inside the probe loop of tile_insert_kernel).
__device__ __forceinline__ unsigned long long lds_u64(uint32_t saddr) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long atoms_cas_u64(uint32_t saddr, unsigned long long cmp, unsigned long long val) {
    unsigned long long old;
    asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(saddr), "l"(cmp), "l"(val) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t atoms_add_u32(uint32_t saddr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(saddr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void reds_add_u32(uint32_t saddr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Add a per-thread value into a global counter with ONE atomic per CTA.  Same-address global
// atomics retire at roughly one per nanosecond chip-wide, so a per-warp atomicAdd on a counter
// (hundreds of thousands per launch) costs more than the kernel's real work.
__device__ __forceinline__ void block_add(unsigned long long *counter, unsigned long long v) {
    __shared__ unsigned long long s_part[32];
    v = warp_sum(v);
    const uint32_t w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();  // s_part may still be in use by a previous block_add
    if ((threadIdx.x & 31) == 0) s_part[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (uint32_t i = 0; i < nw; i++) t += s_part[i];
        if (t) atomicAdd(counter, t);
    }
}

// 0x80 in every byte of v that is zero (exact, no borrow artefacts)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t v) {
    return ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;
}

// Reverse-complement of 32 packed bases (MSB-first): complement, then reverse
// the order of the 2-bit groups.
__device__ __forceinline__ uint64_t rc64(uint64_t x) {
    x = ~x;
    x = __brevll(x);  // reverses bits; fix the order inside each pair
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

// ---------------------------------------------------------------------------
// (1) pack: ASCII -> 2-bit codes + break mask.   1 B read + 0.375 B written / base
//     One thread = 32 input bytes = two 128-bit loads -> one u64 + one u32.
//     Also: n_reads ('\n'), n_bases (ACGT) and the first invalid byte
//     (anything but A,C,G,T,N,'\n'): the reference aborts on it
//     (src/kmer/encoding.rs:353-356), so it is reported, never masked.
// ---------------------------------------------------------------------------

__device__ __forceinline__ void pack_word(uint32_t w, uint32_t &code8, uint32_t &valid4,
                                          uint32_t &n_nl, uint32_t &bad) {
    const uint32_t acgt = zero_bytes(w ^ 0x41414141u) | zero_bytes(w ^ 0x43434343u) |
                          zero_bytes(w ^ 0x47474747u) | zero_bytes(w ^ 0x54545454u);
    uint32_t x = (w >> 1) & 0x03030303u;  // A0 C1 G3 T2
    x ^= (x >> 1) & 0x01010101u;          // A0 C1 G2 T3 (encoding.rs:341-345)
    x &= (acgt >> 7) * 3u;                // breaks pack as 00, like the padding of Read::from_str
    code8 = (x * 0x40100401u) >> 24;      // c0<<6 | c1<<4 | c2<<2 | c3
    const uint32_t is_n = zero_bytes(w ^ 0x4E4E4E4Eu);
    const uint32_t is_nl = zero_bytes(w ^ 0x0A0A0A0Au);
    valid4 = (((acgt >> 7) & 0x01010101u) * 0x08040201u) >> 24 & 0xFu;  // byte0 -> bit 3
    n_nl = __popc(is_nl);
    bad = ~(acgt | is_n | is_nl) & 0x80808080u;
}

static constexpr uint32_t kPackUnits = 4;  // units per thread: one CTA packs 32 KiB of input

__global__ void __launch_bounds__(256)
pack_kernel(const uint8_t *__restrict__ in, uint64_t n_bytes, uint64_t pos_base,
            uint64_t *__restrict__ codes, uint32_t *__restrict__ breaks, uint64_t n_units,
            ChunkCounters *__restrict__ cc, GlobalCounters *__restrict__ gc) {
    unsigned long long my_reads = 0, my_bases = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
#pragma unroll 1
    for (uint32_t r = 0; r < kPackUnits; r++) {
        const uint64_t u = ((uint64_t)blockIdx.x * kPackUnits + r) * blockDim.x + threadIdx.x;
        if (u >= n_units) break;
        const uint64_t off = u * 32;
        uint32_t w[8];
        const bool full = off + 32 <= n_bytes;
        if (full && aligned) {
            const uint4 a = ld_nc_v4(reinterpret_cast<const uint4 *>(in + off));
            const uint4 b = ld_nc_v4(reinterpret_cast<const uint4 *>(in + off) + 1);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
            w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint32_t v = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint64_t p = off + 4 * i + j;
                    // past the end: 0xFF is neither a base nor a newline; masked below
                    const uint32_t byte = p < n_bytes ? in[p] : 0xFFu;
                    v |= byte << (8 * j);
                }
                w[i] = v;
            }
        }
        uint64_t code = 0;
        uint32_t valid = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t c8, v4, nl, bad;
            pack_word(w[i], c8, v4, nl, bad);
            if (!full) {  // ignore bytes past the end
                const uint64_t p = off + 4 * i;
                uint32_t keep = 0;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (p + j < n_bytes) keep |= 0x80u << (8 * j);
                bad &= keep;
                // v4 / nl cannot be set by 0xFF filler bytes
            }
            code |= (uint64_t)c8 << (56 - 8 * i);
            valid |= v4 << (28 - 4 * i);
            my_reads += nl;
            if (bad) {
                const int j = (__ffs(bad) - 1) >> 3;
                const unsigned long long p = pos_base + off + 4 * i + j;
                atomicMin(&gc->first_bad, (p << 8) | ((w[i] >> (8 * j)) & 0xFFu));
            }
        }
        my_bases += __popc(valid);
        codes[u] = code;
        breaks[u] = ~valid;
    }
    block_add(&cc->n_reads, my_reads);
    block_add(&cc->n_bases, my_bases);
}

// Concatenated bases + offsets -> newline-terminated lines (skm_ingest_reads).
// One warp per read; dst offset of read r = offsets[r] + r.
__global__ void __launch_bounds__(256)
add_separators_kernel(const uint8_t *__restrict__ bases, const uint64_t *__restrict__ offsets,
                      uint64_t n_reads, uint8_t *__restrict__ out) {
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (r >= n_reads) return;
    const uint64_t a = offsets[r], b = offsets[r + 1];
    for (uint64_t i = a + lane; i < b; i += 32) out[i + r] = bases[i];
    if (lane == 0) out[b + r] = '\n';
}

// ---------------------------------------------------------------------------
// (2) extract: rolling canonical k-mers from the packed stream.
//     One thread = one 32-base unit.  It needs the k-1 <= 30 bases before its
//     unit: the previous unit's words arrive by warp shuffle from the lane
//     below (lane 0 re-loads them: an L1/L2 hit).  State update per base is the
//     reference's (src/kmer/encoding.rs:359-367):
//        fwd = (fwd << 2 | b) & mask ; rev = rev >> 2 | (3-b) << 2(k-1)
//        n_valid = 0 at a break, else +1 ; emit min(fwd, rev) iff n_valid >= k
//     A break ('N', end of read, padding) resets n_valid exactly as the
//     reference's N branch does (encoding.rs:346-351); a new read starts after
//     every '\n', so no window spans two reads.
// ---------------------------------------------------------------------------

struct UnitInput {
    uint64_t cur, prev;
    uint32_t inv, prev_inv;
};

__device__ __forceinline__ UnitInput load_unit(const uint64_t *__restrict__ codes,
                                               const uint32_t *__restrict__ breaks, uint64_t u,
                                               uint64_t u_end) {
    UnitInput in;
    const bool active = u < u_end;
    in.cur = active ? codes[u] : 0ull;
    in.inv = active ? breaks[u] : 0xFFFFFFFFu;
    in.prev = __shfl_up_sync(0xffffffffu, in.cur, 1);
    in.prev_inv = __shfl_up_sync(0xffffffffu, in.inv, 1);
    if ((threadIdx.x & 31) == 0) {
        if (active && u > 0) {
            in.prev = codes[u - 1];
            in.prev_inv = breaks[u - 1];
        } else {
            in.prev = 0;
            in.prev_inv = 0xFFFFFFFFu;  // nothing before the first unit
        }
    }
    return in;
}

template <class Emit>
__device__ __forceinline__ uint32_t extract_unit(const UnitInput &in, uint32_t k, Emit &&emit) {
    if (in.inv == 0xFFFFFFFFu) return 0;  // no base in this unit
    const uint64_t kmask = (1ull << (2 * k)) - 1;  // k <= 31
    const uint32_t top = 2 * (k - 1);
    uint64_t fwd = in.prev;
    uint64_t rev = rc64(in.prev) >> (64 - 2 * k);  // revcomp of the last k bases before the unit
    uint32_t n_valid = in.prev_inv ? (uint32_t)(__ffs(in.prev_inv) - 1) : 32u;
    uint32_t n_emitted = 0;
#pragma unroll 8
    for (int j = 0; j < 32; j++) {
        const uint64_t b = (in.cur >> (62 - 2 * j)) & 3ull;
        fwd = (fwd << 2) | b;
        rev = (rev >> 2) | ((3ull - b) << top);
        const bool brk = (in.inv >> (31 - j)) & 1u;
        n_valid = brk ? 0u : n_valid + 1u;
        if (n_valid >= k) {
            const uint64_t f = fwd & kmask;
            emit(f < rev ? f : rev, j);
            n_emitted++;
        }
    }
    return n_emitted;
}

// Diagnostic / parity kernel: out[p] = canonical k-mer of the window ending at
// byte p of the segment, or EMPTY.
__global__ void __launch_bounds__(256)
extract_positions_kernel(const uint64_t *__restrict__ codes, const uint32_t *__restrict__ breaks,
                         uint64_t n_units, uint64_t n_bytes, uint32_t k,
                         unsigned long long *__restrict__ out) {
    const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const UnitInput in = load_unit(codes, breaks, u, n_units);
    if (u >= n_units) return;
    for (int j = 0; j < 32; j++)
        if (u * 32 + j < n_bytes) out[u * 32 + j] = SKM_EMPTY_KEY;
    extract_unit(in, k, [&](uint64_t kmer, int j) {
        if (u * 32 + j < n_bytes) out[u * 32 + j] = kmer;
    });
}

// ---------------------------------------------------------------------------
// (3) insert: open-addressing upsert, software-pipelined.
//     One probe = one key load (ld.cg: L2 is the point of coherence); a hit
//     costs one atomic add on the same sector; a miss on EMPTY one atomicCAS.
//     Keys only ever change EMPTY -> key, so a key loaded early can never be a
//     stale "other key": a stale EMPTY is resolved by the CAS.
//
//     The kernels are bound by DRAM latency on random sectors (ncu: >90 % of
//     stalls are long_scoreboard), so every thread keeps D probes in flight:
//     the key load for k-mer j+D is issued before k-mer j is resolved
//     (InsertPipe).  When the running histogram is on, the count update is an
//     atomicAdd that returns the old count, and that result is consumed one
//     step later, so its round trip is hidden as well.
//
//     Running histogram (replaces a full table scan per chunk): moving a k-mer
//     from count `old` to `old+add` moves one unit of histogram mass — exactly
//     Histogram::move_count (src/kmer/histogram.rs:51-85).  Concurrent adds to
//     one k-mer get distinct `old` values from the atomic, and the moves
//     telescope, so the result is exact in any interleaving.  Deltas for counts
//     < kLowBins are privatised in shared memory (one signed copy per lane: no
//     bank conflicts, no same-address serialisation on the singleton bin) and
//     flushed once per CTA; higher bins go to global memory directly.
// ---------------------------------------------------------------------------

static constexpr int kLowBins = 64;

struct HistoSink {
    int *low;                      // shared: kLowBins * 32 signed deltas
    unsigned long long *g_hist;    // global: histo_max + 2 bins, wrapping u64 adds
    unsigned long long histo_max;
    uint32_t lane;
    __device__ __forceinline__ void bump(unsigned long long bin, int d) const {
        if (bin < (unsigned long long)kLowBins)
            atomicAdd(&low[(uint32_t)bin * 32 + lane], d);
        else
            atomicAdd(&g_hist[bin], (unsigned long long)(long long)d);
    }
    __device__ __forceinline__ void move(unsigned long long old, unsigned long long add) const {
        const unsigned long long top = histo_max + 1;
        const unsigned long long nw = old + add;
        const unsigned long long ob = old > histo_max ? top : old;
        const unsigned long long nb = nw > histo_max ? top : nw;
        if (ob == nb) return;
        if (old) bump(ob, -1);
        bump(nb, 1);
    }
};

__device__ __forceinline__ void histo_smem_init(int *low) {
    for (uint32_t i = threadIdx.x; i < kLowBins * 32; i += blockDim.x) low[i] = 0;
    __syncthreads();
}

__device__ __forceinline__ void histo_smem_flush(const int *low, unsigned long long *g_hist) {
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < (uint32_t)kLowBins; b += blockDim.x) {
        long long sum = 0;
        for (int l = 0; l < 32; l++) sum += low[b * 32 + ((l + b) & 31)];
        if (sum) atomicAdd(&g_hist[b], (unsigned long long)sum);
    }
}

template <int D, bool kHisto>
struct InsertPipe {
    TableRef tr;
    Slot *table;
    uint64_t capmask;
    HistoSink hs;
    unsigned long long rk[D], rs[D], rkey[D];
    uint32_t radd[D];
    uint32_t valid = 0;
    unsigned long long o_old = 0;
    uint32_t o_add = 0;
    bool o_valid = false;
    bool part_full = false;
    unsigned long long n_new = 0;

    __device__ __forceinline__ InsertPipe(const TableRef &t, const HistoSink &h)
        : tr(t), table(t.slots), capmask(t.mask()), hs(h) {}

    // resolve the oldest probe: find/claim the slot, add the count
    __device__ __forceinline__ void finish(unsigned long long kmer, unsigned long long s,
                                           unsigned long long key, uint32_t add) {
        uint32_t probes = 0;
        for (;;) {
            if (key == SKM_EMPTY_KEY) {
                key = atomicCAS(&table[s].key, (unsigned long long)SKM_EMPTY_KEY, kmer);
                if (key == SKM_EMPTY_KEY) {
                    n_new++;
                    break;
                }
            }
            if (key == kmer) break;
            if (++probes >= kPartSlots) {  // the whole partition holds other keys: report, never spin
                part_full = true;
                return;
            }
            s = TableRef::next(s);
            key = ld_cg_u64(&table[s].key);
        }
        if (kHisto) {
            if (o_valid) hs.move(o_old, o_add);  // previous atomic's result: long since back
            o_old = atomicAdd(&table[s].count, (unsigned long long)add);
            o_add = add;
            o_valid = true;
        } else {
            red_add_u64(&table[s].count, (unsigned long long)add);
        }
    }

    __device__ __forceinline__ void step(bool have, unsigned long long kmer, uint32_t add) {
        if (valid & 1u) finish(rk[0], rs[0], rkey[0], radd[0]);
#pragma unroll
        for (int i = 0; i + 1 < D; i++) {
            rk[i] = rk[i + 1];
            rs[i] = rs[i + 1];
            rkey[i] = rkey[i + 1];
            radd[i] = radd[i + 1];
        }
        valid >>= 1;
        if (have) {
            const unsigned long long s = tr.home(kmer);
            rk[D - 1] = kmer;
            rs[D - 1] = s;
            radd[D - 1] = add;
            rkey[D - 1] = ld_cg_u64(&table[s].key);  // in flight until this entry reaches the front
            valid |= 1u << (D - 1);
        }
    }
    __device__ __forceinline__ void push(unsigned long long kmer, uint32_t add = 1) { step(true, kmer, add); }
    __device__ __forceinline__ void drain() {
#pragma unroll
        for (int i = 0; i < D; i++) step(false, 0, 0);
        if (kHisto && o_valid) hs.move(o_old, o_add);
        o_valid = false;
    }
};

// Fused extract + insert ("direct" mode): k-mers never touch HBM.
template <int D, bool kHisto>
__global__ void __launch_bounds__(256)
extract_insert_kernel(const uint64_t *__restrict__ codes, const uint32_t *__restrict__ breaks,
                      uint64_t u_begin, uint64_t u_end, uint32_t k, TableRef table,
                      ChunkCounters *__restrict__ cc, GlobalCounters *__restrict__ gc,
                      unsigned long long *__restrict__ g_hist, unsigned long long histo_max) {
    __shared__ int s_low[kHisto ? kLowBins * 32 : 1];
    if (kHisto) histo_smem_init(s_low);
    const uint64_t u = u_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const UnitInput in = load_unit(codes, breaks, u, u_end);
    HistoSink hs{s_low, g_hist, histo_max, threadIdx.x & 31};
    InsertPipe<D, kHisto> pipe(table, hs);
    unsigned long long n_win = extract_unit(in, k, [&](uint64_t kmer, int) { pipe.push(kmer); });
    pipe.drain();
    if (pipe.part_full) gc->part_full = 1ull;
    block_add(&gc->n_distinct, pipe.n_new);
    block_add(&cc->n_windows, n_win);
    if (kHisto) histo_smem_flush(s_low, g_hist);
}

// Warp-cooperative probing (the variant BASELINE.json's north star names), kept as an A/B switch for the
// direct kernel (SKM_WARP_COOP=1): 8 lanes inspect the 8 slots from the home slot on (one 128-byte line when
// the home slot is line-aligned), ballot for a match or an empty slot, and the first in probe order wins.
// A warp counts 4 k-mers at a time instead of 32.  Measured slower than one thread per k-mer at loads
// 0.5 and 0.7 (profiles/experiments_r02.md): a probe sequence is 1.3-2.2 slots long, so seven of the eight
// loads are wasted, and the table is bound by sector throughput, not by the latency of a probe chain.
template <bool kHisto>
__global__ void __launch_bounds__(256)
extract_insert_coop_kernel(const uint64_t *__restrict__ codes, const uint32_t *__restrict__ breaks,
                           uint64_t u_begin, uint64_t u_end, uint32_t k, TableRef table,
                           ChunkCounters *__restrict__ cc, GlobalCounters *__restrict__ gc,
                           unsigned long long *__restrict__ g_hist, unsigned long long histo_max) {
    __shared__ int s_low[kHisto ? kLowBins * 32 : 1];
    if (kHisto) histo_smem_init(s_low);
    const uint32_t lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
    const uint64_t u = u_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const UnitInput in = load_unit(codes, breaks, u, u_end);
    HistoSink hs{s_low, g_hist, histo_max, lane};
    Slot *slots = table.slots;
    unsigned long long n_new = 0, n_win = 0;
    bool part_full = false;
    // the rolling state of extract_unit, advanced for every lane in lockstep
    const uint64_t kmask = (1ull << (2 * k)) - 1;
    const uint32_t top = 2 * (k - 1);
    uint64_t fwd = in.prev;
    uint64_t rev = rc64(in.prev) >> (64 - 2 * k);
    uint32_t n_valid = in.prev_inv ? (uint32_t)(__ffs(in.prev_inv) - 1) : 32u;
    const bool unit_has_bases = in.inv != 0xFFFFFFFFu;
    for (int j = 0; j < 32; j++) {
        const uint64_t b = (in.cur >> (62 - 2 * j)) & 3ull;
        fwd = (fwd << 2) | b;
        rev = (rev >> 2) | ((3ull - b) << top);
        const bool brk = (in.inv >> (31 - j)) & 1u;
        n_valid = brk ? 0u : n_valid + 1u;
        const bool have = unit_has_bases && n_valid >= k;
        const uint64_t f = fwd & kmask;
        const unsigned long long mine = f < rev ? f : rev;
        n_win += have;
        const uint32_t have_mask = __ballot_sync(0xffffffffu, have);
        // eight rounds: round r serves the k-mers of lanes 4r .. 4r+3, one per 8-lane group
        for (uint32_t r = 0; r < 8; r++) {
            const uint32_t src = 4 * r + grp;
            if (!((have_mask >> (4 * r)) & 0xFu)) continue;   // (uniform: nobody in this round has a k-mer)
            const unsigned long long kmer = __shfl_sync(0xffffffffu, mine, src);
            bool active = (have_mask >> src) & 1u;
            uint64_t s = table.home(kmer);
            uint32_t probes = 0;
            while (__any_sync(0xffffffffu, active)) {
                uint64_t slot = s;
                for (uint32_t q = 0; q < sub; q++) slot = TableRef::next(slot);
                const unsigned long long key = active ? ld_cg_u64(&slots[slot].key) : 0ull;
                const uint32_t m_match = (__ballot_sync(0xffffffffu, active && key == kmer) >> (8 * grp)) & 0xFFu;
                const uint32_t m_empty = (__ballot_sync(0xffffffffu, active && key == SKM_EMPTY_KEY) >> (8 * grp)) & 0xFFu;
                if (!active) continue;
                const int i_match = m_match ? __ffs(m_match) - 1 : 8, i_empty = m_empty ? __ffs(m_empty) - 1 : 8;
                int target = -1;       // the slot (sub-lane) that gets the count
                bool retry = false;
                if (i_match < i_empty) {
                    target = i_match;
                } else if (i_empty < 8) {
                    unsigned long long prev = 0;
                    if ((int)sub == i_empty) prev = atomicCAS(&slots[slot].key, (unsigned long long)SKM_EMPTY_KEY, kmer);
                    prev = __shfl_sync(__activemask(), prev, 8 * grp + i_empty);
                    if (prev == SKM_EMPTY_KEY) {
                        target = i_empty;
                        if ((int)sub == i_empty) n_new++;
                    } else if (prev == kmer) {
                        target = i_empty;
                    } else {
                        retry = true;   // another k-mer took the slot: look at the same window again
                    }
                }
                if (target >= 0) {
                    if ((int)sub == target) {
                        if (kHisto) hs.move(atomicAdd(&slots[slot].count, 1ull), 1);
                        else red_add_u64(&slots[slot].count, 1ull);
                    }
                    active = false;
                } else if (!retry) {
                    probes += 8;
                    if (probes >= kPartSlots) {
                        part_full = true;
                        active = false;
                    }
                    for (uint32_t q = 0; q < 8; q++) s = TableRef::next(s);
                }
            }
        }
    }
    if (part_full) gc->part_full = 1ull;
    block_add(&gc->n_distinct, n_new);
    block_add(&cc->n_windows, n_win);
    if (kHisto) histo_smem_flush(s_low, g_hist);
}

// Count the windows of a unit range (conservation checks, list sizing).
__global__ void __launch_bounds__(256)
count_windows_kernel(const uint64_t *__restrict__ codes, const uint32_t *__restrict__ breaks,
                     uint64_t u_begin, uint64_t u_end, uint32_t k, ChunkCounters *__restrict__ cc) {
    const uint64_t u = u_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const UnitInput in = load_unit(codes, breaks, u, u_end);
    unsigned long long n_win = extract_unit(in, k, [&](uint64_t, int) {});
    block_add(&cc->n_windows, n_win);
}

// Insert a flat list of k-mers (received from other ranks, or a region-sorted
// bucket list).  CTA b owns the contiguous slice [b*kListTile, (b+1)*kListTile):
// CTAs walk a region-sorted list in order, so at any moment the chip works on a
// few neighbouring table regions that stay resident in L2.
static constexpr uint32_t kListPerThread = 16;
static constexpr uint32_t kListTile = 256 * kListPerThread;

// One run = a contiguous list of k-mers (optionally with counts).  A launch walks a sequence of
// runs in order; tiles never span runs.
struct RunDesc {
    const unsigned long long *kmers;
    const uint32_t *counts;          // null => every k-mer counts 1
    unsigned long long n;
    unsigned long long tile_begin;   // index of this run's first warp tile
};

// Zero-fill by kernel.  cudaMemsetAsync may be serviced by a copy engine, where it queues behind
// host-to-device batches already submitted (measured: a routing pass submitted after ten 151 MB
// copies did not start until the last copy had finished).
__global__ void zero_kernel(unsigned long long *__restrict__ p, uint64_t n_words) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words;
         i += (uint64_t)gridDim.x * blockDim.x)
        p[i] = 0ull;
}

// Small host <-> device transfers go through kernels that read / write pinned host memory: a
// cudaMemcpy would queue on a copy engine BEHIND the read batches still in flight (measured in
// round 1: the first insert could not start until every batch had arrived).
// Same reason, other direction: small results the host waits for (bucket totals) are stored to
// pinned host memory by a kernel instead of a device-to-host copy.
__global__ void copy_words_kernel(const unsigned long long *__restrict__ src, unsigned long long *__restrict__ dst,
                                  uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// Persistent kernel: every WARP pulls the next 512-k-mer tile from a global counter, so there is
// no CTA-wide barrier inside the loop (ncu on the one-tile-per-CTA version: 25-30 % of issue
// slots stalled on barriers because probe latencies differ between warps) and tiles are still
// handed out in list order, which keeps the chip on a few neighbouring table regions.
static constexpr uint32_t kWarpTile = 32 * kListPerThread;  // k-mers per warp tile

// 6 CTAs per SM caps the kernel at 40 registers (no spills): 5 resident CTAs then leave 14 K registers
// per SM, enough for one CTA of a bucketing kernel of the next chunk to run beside the inserts
// (profiles/experiments_r01.md #25: 43 registers, 5 resident, nothing beside them: 62.0 ms per step;
// 40 registers, 6 resident: 58.3 ms; 40 registers, 5 resident + bucketing beside them: 54.0 ms).
#ifndef SKM_INSERT_MIN_CTAS
#define SKM_INSERT_MIN_CTAS 6
#endif
template <int D, bool kHisto>
__global__ void __launch_bounds__(256, D == 1 ? SKM_INSERT_MIN_CTAS : 1)
insert_runs_kernel(const RunDesc *__restrict__ descs, uint32_t n_desc, RunDesc single,
                   const unsigned long long *__restrict__ n_dev, unsigned long long n_tiles,
                   unsigned long long *__restrict__ tile_counter, TableRef table,
                   GlobalCounters *__restrict__ gc, unsigned long long *__restrict__ g_hist,
                   unsigned long long histo_max) {
    __shared__ int s_low[kHisto ? kLowBins * 32 : 1];
    if (kHisto) histo_smem_init(s_low);
    const uint32_t lane = threadIdx.x & 31;
    if (descs == nullptr) {
        // single run; n_dev (optional): its length lives in device memory (written by the bucketing
        // scan) so the host can queue this launch without waiting for it
        if (n_dev && *n_dev < single.n) single.n = *n_dev;
        n_tiles = (single.n + kWarpTile - 1) / kWarpTile;
    }
    HistoSink hs{s_low, g_hist, histo_max, lane};
    InsertPipe<D, kHisto> pipe(table, hs);
    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(tile_counter, 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tiles) break;
        RunDesc d;
        if (descs == nullptr) {
            d = single;
        } else {
            uint32_t lo = 0, hi = n_desc;  // last run whose tile_begin <= t
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (descs[mid].tile_begin <= t) lo = mid; else hi = mid;
            }
            d = descs[lo];
            t -= d.tile_begin;
        }
        const uint64_t base = t * kWarpTile + lane;
#pragma unroll 4
        for (uint32_t j = 0; j < kListPerThread; j++) {
            const uint64_t i = base + (uint64_t)j * 32;
            if (i < d.n) pipe.push(d.kmers[i], d.counts ? d.counts[i] : 1u);
        }
    }
    pipe.drain();
    if (pipe.part_full) gc->part_full = 1ull;
    block_add(&gc->n_distinct, pipe.n_new);
    if (kHisto) histo_smem_flush(s_low, g_hist);
}

__global__ void __launch_bounds__(256) table_clear_kernel(Slot *__restrict__ table, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 v;
        v.x = v.y = 0xFFFFFFFFu;
        v.z = v.w = 0u;
        reinterpret_cast<uint4 *>(table)[i] = v;
    }
}

// Grow: re-insert every occupied slot of the old table into the new one.
__global__ void __launch_bounds__(256)
rehash_kernel(const Slot *__restrict__ old_table, uint64_t old_cap, TableRef nt, GlobalCounters *__restrict__ gc) {
    Slot *__restrict__ table = nt.slots;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < old_cap;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 v = ld_nc_v4(reinterpret_cast<const uint4 *>(old_table) + i);
        const unsigned long long key = ((unsigned long long)v.y << 32) | v.x;
        if (key == SKM_EMPTY_KEY) continue;
        const unsigned long long cnt = ((unsigned long long)v.w << 32) | v.z;
        // keys are unique in the old table: claim the first EMPTY slot of the probe sequence
        uint64_t sl = nt.home(key);
        uint32_t probes = 0;
        for (;;) {
            if (ld_cg_u64(&table[sl].key) == SKM_EMPTY_KEY &&
                atomicCAS(&table[sl].key, (unsigned long long)SKM_EMPTY_KEY, key) == SKM_EMPTY_KEY)
                break;
            if (++probes >= kPartSlots) break;
            sl = TableRef::next(sl);
        }
        if (probes >= kPartSlots) {
            gc->part_full = 1ull;
            continue;
        }
        table[sl].count = cnt;
    }
}

// ---------------------------------------------------------------------------
// (4) histogram of the whole table (one streaming pass, 16 B / slot).
//     Shared-memory privatised bins: counts 1..63 go to a lane-private copy
//     (address = bin*32 + lane: no bank conflicts, no same-address
//     serialisation on the singleton bin); 64 <= count < n_smem_bins go to one
//     shared copy; the rest (rare) straight to global.  Bin histo_max+1 collects
//     every count > histo_max (src/kmer/histogram.rs:80-84,125-134).
// ---------------------------------------------------------------------------

__global__ void __launch_bounds__(512)
histogram_kernel(const Slot *__restrict__ table, uint64_t capacity, uint64_t histo_max,
                 uint32_t n_smem_bins, unsigned long long *__restrict__ g_bins,
                 HistoTotals *__restrict__ totals, int want_digest) {
    extern __shared__ uint32_t smem[];
    uint32_t *low = smem;                      // kLowBins * 32
    uint32_t *bins = smem + kLowBins * 32;     // n_smem_bins
    for (uint32_t i = threadIdx.x; i < kLowBins * 32 + n_smem_bins; i += blockDim.x) smem[i] = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    unsigned long long n_distinct = 0, n_kmers = 0, n_sat = 0, digest = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < capacity;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 v = ld_nc_v4(reinterpret_cast<const uint4 *>(table) + i);
        const unsigned long long key = ((unsigned long long)v.y << 32) | v.x;
        if (key == SKM_EMPTY_KEY) continue;
        unsigned long long cnt = ((unsigned long long)v.w << 32) | v.z;
        if (cnt >= kU32Max) {
            cnt = kU32Max;
            n_sat++;
        }
        n_distinct++;
        n_kmers += cnt;
        if (want_digest) digest += skm_pair_digest(key, (uint32_t)cnt);
        const unsigned long long bin = cnt > histo_max ? histo_max + 1 : cnt;
        if (bin < (unsigned long long)kLowBins)
            atomicAdd(&low[(uint32_t)bin * 32 + lane], 1u);
        else if (bin < n_smem_bins)
            atomicAdd(&bins[(uint32_t)bin], 1u);
        else
            atomicAdd(&g_bins[bin], 1ull);
    }
    __syncthreads();
    // flush: low bins are summed over their 32 lane copies
    for (uint32_t b = threadIdx.x; b < (uint32_t)kLowBins; b += blockDim.x) {
        unsigned long long s = 0;
        for (int l = 0; l < 32; l++) s += low[b * 32 + ((l + b) & 31)];
        if (s) {
            const unsigned long long dst = (unsigned long long)b > histo_max ? histo_max + 1 : b;
            atomicAdd(&g_bins[dst], s);
        }
    }
    for (uint32_t b = kLowBins + threadIdx.x; b < n_smem_bins; b += blockDim.x) {
        const uint32_t s = bins[b];
        if (s) atomicAdd(&g_bins[b], (unsigned long long)s);
    }
    n_distinct = warp_sum(n_distinct);
    n_kmers = warp_sum(n_kmers);
    n_sat = warp_sum(n_sat);
    digest = warp_sum(digest);
    if (lane == 0) {
        if (n_distinct) atomicAdd(&totals->n_distinct, n_distinct);
        if (n_kmers) atomicAdd(&totals->n_kmers, n_kmers);
        if (n_sat) atomicAdd(&totals->n_saturated, n_sat);
        if (digest) atomicAdd(&totals->digest, digest);
    }
}

// ---------------------------------------------------------------------------
// read side: export (compaction) and batched lookups
// ---------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
export_kernel(const Slot *__restrict__ table, uint64_t capacity, unsigned long long *__restrict__ keys,
              uint32_t *__restrict__ counts, uint64_t out_cap, unsigned long long *__restrict__ cursor) {
    const uint32_t lane = threadIdx.x & 31;
    // uniform trip count per warp so the ballots below are well formed
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t first = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n_iter = (capacity + stride - 1) / stride;
    for (uint64_t it = 0; it < n_iter; it++) {
        const uint64_t i = first + it * stride;
        unsigned long long key = SKM_EMPTY_KEY, cnt = 0;
        if (i < capacity) {
            const uint4 v = ld_nc_v4(reinterpret_cast<const uint4 *>(table) + i);
            key = ((unsigned long long)v.y << 32) | v.x;
            cnt = ((unsigned long long)v.w << 32) | v.z;
        }
        const bool occ = key != SKM_EMPTY_KEY;
        const uint32_t m = __ballot_sync(0xffffffffu, occ);
        if (m == 0) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (occ) {
            const unsigned long long o = base + __popc(m & ((1u << lane) - 1));
            if (o < out_cap) {
                keys[o] = key;
                counts[o] = cnt >= kU32Max ? kU32Max : (uint32_t)cnt;
            }
        }
    }
}

// FilteredKmerCounts::get_canonical_count / KmerCounts::get_count / get_canonical
// (src/kmer/counting.rs:205-222,328-342).  mode 0: probe min(kmer, revcomp);
// mode 1: probe the k-mer as given; mode 2: probe as given, else the revcomp.
__device__ __forceinline__ bool table_find(const TableRef &tr, uint64_t q, uint32_t &count) {
    if (q == SKM_EMPTY_KEY) return false;
    const Slot *__restrict__ table = tr.slots;
    uint64_t s = tr.home(q);
    for (uint32_t probes = 0; probes < kPartSlots; probes++) {
        const uint4 v = *(reinterpret_cast<const uint4 *>(table) + s);
        const unsigned long long key = ((unsigned long long)v.y << 32) | v.x;
        if (key == q) {
            const unsigned long long cnt = ((unsigned long long)v.w << 32) | v.z;
            count = cnt >= kU32Max ? kU32Max : (uint32_t)cnt;
            return true;
        }
        if (key == SKM_EMPTY_KEY) return false;
        s = TableRef::next(s);
    }
    return false;
}

__global__ void __launch_bounds__(256)
lookup_kernel(TableRef table, uint32_t k,
              const unsigned long long *__restrict__ kmers, uint64_t n, uint32_t min_count,
              int mode, uint32_t *__restrict__ counts, uint8_t *__restrict__ found) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t q = kmers[i];
        const uint64_t rc = skm_revcomp_kmer(q, k);
        uint32_t c = 0;
        bool hit;
        if (mode == 0)
            hit = table_find(table, q < rc ? q : rc, c);
        else if (mode == 1)
            hit = table_find(table, q, c);
        else
            hit = table_find(table, q, c) || table_find(table, rc, c);
        if (hit && c < min_count) {
            hit = false;
            c = 0;
        }
        if (!hit) c = 0;
        if (counts) counts[i] = c;
        if (found) found[i] = hit ? 1 : 0;
    }
}

// find_oligos_in_kmers (src/pcr/primers.rs:163-226) as one streaming pass over the table:
// a k-mer with count >= min_count matches if its first oligo_length bases are in `fwd_set`
// (oligos shifted to the top of the k-mer; reported as is), else if its last oligo_length bases
// are in `rc_set` (reverse complements of the oligos; its reverse complement is reported).
// Both sets are sorted arrays searched by bisection (a few hundred entries, L1-resident).
__device__ __forceinline__ bool sorted_contains(const unsigned long long *__restrict__ a, uint32_t n,
                                                unsigned long long x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const unsigned long long v = __ldg(&a[mid]);
        if (v == x) return true;
        if (v < x) lo = mid + 1; else hi = mid;
    }
    return false;
}

// A 2^16-bit filter per set on the first (up to) 8 bases of the masked value rejects almost every slot
// with two shared-memory loads, so the scan streams the table instead of bisecting for every slot
// (6.2 ms -> per 8.6 GB table before, profiles/experiments_r02.md).
static constexpr uint32_t kScanFilterWords = 2048;   // 65536 bits
__global__ void __launch_bounds__(256)
scan_oligos_kernel(const Slot *__restrict__ table, uint64_t capacity, uint32_t k,
                   const unsigned long long *__restrict__ fwd_set, const unsigned long long *__restrict__ rc_set,
                   uint32_t n_set, unsigned long long mask, unsigned long long rc_mask, uint32_t min_count,
                   unsigned long long *__restrict__ out_keys, uint32_t *__restrict__ out_counts,
                   uint64_t out_cap, unsigned long long *__restrict__ cursor, uint32_t oligo_length) {
    __shared__ uint32_t f_fwd[kScanFilterWords], f_rc[kScanFilterWords];
    for (uint32_t i = threadIdx.x; i < kScanFilterWords; i += blockDim.x) {
        f_fwd[i] = 0;
        f_rc[i] = 0;
    }
    __syncthreads();
    // filter key = the leading min(8, oligo_length) bases of the oligo
    const uint32_t fbits = 2 * (oligo_length < 8 ? oligo_length : 8);
    const uint32_t sh_fwd = 2 * k - fbits;               // masked forward value: oligo sits at the top of the 2k bits
    const uint32_t sh_rc = 2 * oligo_length - fbits;     // masked suffix value: 2 * oligo_length bits
    for (uint32_t i = threadIdx.x; i < n_set; i += blockDim.x) {
        const uint32_t a = (uint32_t)(fwd_set[i] >> sh_fwd), b = (uint32_t)(rc_set[i] >> sh_rc);
        atomicOr(&f_fwd[a >> 5], 1u << (a & 31));
        atomicOr(&f_rc[b >> 5], 1u << (b & 31));
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t first = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n_iter = (capacity + stride - 1) / stride;
    for (uint64_t it = 0; it < n_iter; it++) {
        const uint64_t i = first + it * stride;
        unsigned long long key = SKM_EMPTY_KEY, out = 0;
        uint32_t c = 0;
        bool hit = false;
        if (i < capacity) {
            const uint4 v = ld_nc_v4(reinterpret_cast<const uint4 *>(table) + i);
            key = ((unsigned long long)v.y << 32) | v.x;
            const unsigned long long cnt = ((unsigned long long)v.w << 32) | v.z;
            c = cnt >= kU32Max ? kU32Max : (uint32_t)cnt;
        }
        if (key != SKM_EMPTY_KEY && c >= min_count) {
            const uint32_t a = (uint32_t)((key & mask) >> sh_fwd), b = (uint32_t)((key & rc_mask) >> sh_rc);
            if (((f_fwd[a >> 5] >> (a & 31)) & 1u) && sorted_contains(fwd_set, n_set, key & mask)) {
                hit = true;
                out = key;
            } else if (((f_rc[b >> 5] >> (b & 31)) & 1u) && sorted_contains(rc_set, n_set, key & rc_mask)) {
                hit = true;
                out = skm_revcomp_kmer(key, k);
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (m == 0) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (hit) {
            const unsigned long long o = base + __popc(m & ((1u << lane) - 1));
            if (o < out_cap) {
                out_keys[o] = out;
                out_counts[o] = c;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// bucketing by hash, one scheme for both uses: bucket = (owner rank, table region of that
// owner) = the top bits of (owner, local hash).  On one GPU it orders a chunk's k-mers by
// table region (partitioned insert); across GPUs the same pass groups them by destination
// rank AND leaves every destination's run region-sorted.  Exact two-pass scheme:
//   pass 1  per-bucket counts (shared-memory counters, one global add per
//           non-empty bucket per CTA)
//   scan    exclusive prefix sum over buckets (single CTA)
//   pass 2  each CTA reserves its slice of every bucket with one global atomic
//           per bucket, then its threads scatter their k-mers; neighbouring
//           CTAs fill neighbouring 8-byte cells, so L2 merges them into full
//           sectors before they reach DRAM.
// ---------------------------------------------------------------------------

struct BucketFn {
    uint32_t n_ranks;       // owners (1 on a single GPU)
    uint32_t log2_regions;  // table regions per owner = 2^log2_regions
    __host__ __device__ __forceinline__ uint32_t operator()(uint64_t kmer) const {
        const uint64_t h = skm_hash_kmer(kmer);
        // one owner: owner = 0 and the local hash is h itself (skips two 64-bit multiplies per call;
        // the bucketing kernels are instruction bound and hash every k-mer three times)
        if (n_ranks == 1) return log2_regions ? (uint32_t)(h >> (64u - log2_regions)) : 0u;
        const uint32_t owner = skm_owner_rank(h, n_ranks);
        if (log2_regions == 0) return owner;
        return (owner << log2_regions) | (uint32_t)(skm_local_hash(h, n_ranks) >> (64u - log2_regions));
    }
};

// Each CTA covers kBucketUnits * 256 units (= kBucketUnits * 8192 bases), so that it owns
// several k-mers per bucket and the per-bucket global reservations amortise.
static constexpr uint32_t kBucketUnits = 4;

__global__ void __launch_bounds__(256)
bucket_count_kernel(const uint64_t *__restrict__ codes, const uint32_t *__restrict__ breaks,
                    uint64_t u_begin, uint64_t u_end, uint32_t k, BucketFn fn, uint32_t n_buckets,
                    unsigned long long *__restrict__ g_counts, ChunkCounters *__restrict__ cc) {
    extern __shared__ uint32_t s_cnt[];
    for (uint32_t i = threadIdx.x; i < n_buckets; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    unsigned long long n_win = 0;
#pragma unroll 1
    for (uint32_t r = 0; r < kBucketUnits; r++) {
        const uint64_t u = u_begin + ((uint64_t)blockIdx.x * kBucketUnits + r) * blockDim.x + threadIdx.x;
        const UnitInput in = load_unit(codes, breaks, u, u_end);
        n_win += extract_unit(in, k, [&](uint64_t kmer, int) { atomicAdd(&s_cnt[fn(kmer)], 1u); });
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_buckets; i += blockDim.x) {
        const uint32_t c = s_cnt[i];
        if (c) atomicAdd(&g_counts[i], (unsigned long long)c);
    }
    block_add(&cc->n_windows, n_win);
}

// offsets[b] = sum of counts[0..b); cursors[b] = offsets[b]; offsets[n] = total
// One CTA of 256 threads, not 1024: beside the persistent insert kernel (5 x 256 threads per SM)
// a 1024-thread CTA fits on no SM and the whole bucketing of the next chunk waited for the inserts
// to finish (profiles/experiments_r01.md #30).
static constexpr uint32_t kScanThreads = 256;
__global__ void __launch_bounds__(kScanThreads)
bucket_scan_kernel(const unsigned long long *__restrict__ counts, uint32_t n_buckets,
                   unsigned long long *__restrict__ offsets, unsigned long long *__restrict__ cursors) {
    __shared__ unsigned long long part[kScanThreads];
    // (one CTA per array of n_buckets + 1 entries: the multi-GPU sender scans all owners' counts in one launch)
    counts += (size_t)blockIdx.x * (n_buckets + 1);
    offsets += (size_t)blockIdx.x * (n_buckets + 1);
    cursors += (size_t)blockIdx.x * (n_buckets + 1);
    const uint32_t per = (n_buckets + blockDim.x - 1) / blockDim.x;
    const uint32_t a = threadIdx.x * per;
    unsigned long long s = 0;
    for (uint32_t i = a; i < a + per && i < n_buckets; i++) s += counts[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (uint32_t i = 0; i < blockDim.x; i++) {
            const unsigned long long t = part[i];
            part[i] = run;
            run += t;
        }
        offsets[n_buckets] = run;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (uint32_t i = a; i < a + per && i < n_buckets; i++) {
        offsets[i] = run;
        cursors[i] = run;
        run += counts[i];
    }
}

// Scatter with shared-memory staging: the CTA's k-mers are first laid out in bucket order in
// shared memory, then copied out so that consecutive threads write consecutive cells of a
// bucket run (whole 32-byte sectors / 128-byte lines instead of lone 8-byte stores, which cost
// one L2 request each).  One CTA = kScatterThreads units = kScatterThreads*32 bases.
//
// The kernel is bound by instruction issue, so every k-mer is hashed ONCE: round 1 extracts the
// unit's windows, hashes them, and takes each k-mer's rank inside its bucket from the one
// shared-memory atomic that also counts the bucket; (bucket, rank) of the unit's 32 windows stay in
// registers.  After the scan of the bucket counts, round 2 re-extracts (a rolling update, no hash,
// no atomic) and drops every k-mer at start[bucket] + rank, with its bucket id beside it (u16), so
// the copy-out needs no hash either.
static constexpr uint32_t kMaxBuckets = 1024;  // scatter stages a CTA's k-mers in bucket order: runs stay >= 10 k-mers
static constexpr uint32_t kScatterThreads = 320;                    // 10240 positions per CTA
static constexpr uint32_t kScatterStage = kScatterThreads * 32;     // max k-mers per CTA

__host__ __device__ inline size_t scatter_smem_bytes(uint32_t n_buckets, bool capped = false) {
    // stage (u64) + bucket ids (u16) | per bucket 8 B (exact: s_gbase u64; capped: count / s_g, s_obase u32) +
    // s_room u16 + s_start u16.  112 KiB at 1024 buckets: two CTAs per SM.
    (void)capped;
    return (size_t)kScatterStage * 10 + (size_t)n_buckets * 12;
}

// Capped layout (single GPU, no counting pass): bucket b owns cells [b * cap, (b + 1) * cap) of the
// list; what does not fit goes to one overflow run of `ovf_cap` cells at `ovf_base`.  With a
// uniform hash the buckets of a batch differ by a few sigma, so `cap` = mean + 6 % and the overflow
// run stays empty; skewed inputs (one k-mer repeated millions of times) spill into it, and if even
// that is too small the host sees cursors[n_buckets] > ovf_cap and re-buckets the batch exactly.
struct CapLayout {
    unsigned long long cap, ovf_base, ovf_cap;
};

// The state machine of extract_unit with the loop fully unrolled (j is a compile-time constant in
// `emit`, so per-window values can live in a register array).  kValidity = false: the caller knows
// which windows exist (round 2) and only wants the canonical k-mer of every position.
// The reverse strand is kept top-aligned (newest complement at bits 63:62), so its update is a
// constant shift; one shift by 64 - 2k brings the window down when it is emitted.
template <bool kValidity, class Emit>
__device__ __forceinline__ void extract_unit_unrolled(const UnitInput &in, uint32_t k, Emit &&emit) {
    const uint64_t kmask = (1ull << (2 * k)) - 1;  // k <= 31
    const uint32_t down = 64 - 2 * k;
    uint64_t fwd = in.prev;
    uint64_t rev_top = rc64(in.prev);
    uint32_t n_valid = in.prev_inv ? (uint32_t)(__ffs(in.prev_inv) - 1) : 32u;
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const uint64_t b = (in.cur >> (62 - 2 * j)) & 3ull;
        fwd = (fwd << 2) | b;
        rev_top = (rev_top >> 2) | ((3ull - b) << 62);
        bool have = true;
        if (kValidity) {
            const bool brk = (in.inv >> (31 - j)) & 1u;
            n_valid = brk ? 0u : n_valid + 1u;
            have = n_valid >= k;
        }
        const uint64_t f = fwd & kmask, r = rev_top >> down;
        emit(f < r ? f : r, j, have);
    }
}

static constexpr uint32_t kScatterPer = (kMaxBuckets + kScatterThreads - 1) / kScatterThreads;  // buckets per thread
static constexpr uint32_t kScatterRankShift = 10;   // (bucket | rank << 10): kMaxBuckets = 2^10, ranks < kScatterStage < 2^14
static_assert(kMaxBuckets <= (1u << kScatterRankShift) && kScatterStage < (1u << 16), "bucket, rank and stage offsets share small words");

// kCapped = false: cursors[] start at the exact bucket offsets of the counting pass.
// kCapped = true : cursors[] start at 0 and end as the bucket totals (cursors[n_buckets] = overflow
//                  total); also counts the windows of the chunk (the exact path does that in pass 1).
template <bool kCapped>
__global__ void __launch_bounds__(kScatterThreads, 2)
bucket_scatter_kernel(const uint64_t *__restrict__ codes, const uint32_t *__restrict__ breaks,
                      uint64_t u_begin, uint64_t u_end, uint32_t k, BucketFn fn, uint32_t n_buckets,
                      unsigned long long *__restrict__ cursors, unsigned long long *__restrict__ out, CapLayout lay,
                      ChunkCounters *__restrict__ cc) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned long long *stage = reinterpret_cast<unsigned long long *>(s_raw);
    uint16_t *stage_b = reinterpret_cast<uint16_t *>(stage + kScatterStage);       // bucket of every staged k-mer
    unsigned char *per_bucket = reinterpret_cast<unsigned char *>(stage_b + kScatterStage);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(per_bucket);            // round 1: this CTA's count per bucket
    // (a thread reads the counts of ITS buckets into registers, everybody passes a barrier, and only then
    //  are the arrays below, which overlay the counts, written)
    unsigned long long *s_gbase = reinterpret_cast<unsigned long long *>(per_bucket);   // exact: first list cell of the run
    uint32_t *s_g = cnt;                        // capped: the run's first cell inside the bucket's region
    uint32_t *s_obase = cnt + n_buckets;        // capped: first overflow cell of what does not fit
    uint16_t *s_room = reinterpret_cast<uint16_t *>(per_bucket + (size_t)8 * n_buckets);  // capped: how many of the run fit
    uint16_t *s_start = s_room + n_buckets;     // first stage slot of the bucket
    __shared__ uint32_t s_warp_tot[kScatterThreads / 32];

    for (uint32_t i = threadIdx.x; i < n_buckets; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    const uint64_t u = u_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const UnitInput in = load_unit(codes, breaks, u, u_end);
    const bool has_bases = in.inv != 0xFFFFFFFFu;
    // round 1: hash, count, rank
    uint32_t br[32];
    if (has_bases) {
        extract_unit_unrolled<true>(in, k, [&](uint64_t kmer, int j, bool have) {
            uint32_t v = 0xFFFFFFFFu;
            if (have) {
                const uint32_t b = fn(kmer);
                v = b | (atomicAdd(&cnt[b], 1u) << kScatterRankShift);
            }
            br[j] = v;
        });
    }
    __syncthreads();
    // exclusive scan of the counts (bucket starts inside the stage) + global reservation; the
    // reservations of a thread's buckets are all issued before any result is used
    const uint32_t per = (n_buckets + blockDim.x - 1) / blockDim.x;  // <= kScatterPer
    const uint32_t b0 = threadIdx.x * per;
    uint32_t c_[kScatterPer];
    unsigned long long g[kScatterPer];
    uint32_t mine = 0;
#pragma unroll
    for (uint32_t j = 0; j < kScatterPer; j++) {
        c_[j] = (j < per && b0 + j < n_buckets) ? cnt[b0 + j] : 0u;
        mine += c_[j];
    }
#pragma unroll
    for (uint32_t j = 0; j < kScatterPer; j++)
        g[j] = c_[j] ? atomicAdd(&cursors[b0 + j], (unsigned long long)c_[j]) : 0ull;
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= (uint32_t)o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();   // (every count has been read: the overlays may be written)
    uint32_t warp_off = 0;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) warp_off += s_warp_tot[w];
    uint32_t run = warp_off + incl - mine;
#pragma unroll
    for (uint32_t j = 0; j < kScatterPer; j++) {
        const uint32_t i = b0 + j;
        if (j < per && i < n_buckets) {
            s_start[i] = (uint16_t)run;
            run += c_[j];
        }
    }
    __syncthreads();
    uint32_t total = 0;
    for (uint32_t w = 0; w < kScatterThreads / 32; w++) total += s_warp_tot[w];
    // round 2: re-extract and place every k-mer at its bucket-ordered stage slot.  (The global reservations
    // g[] are first used after this round: their round trip to L2 is hidden behind it.)
    if (has_bases) {
        extract_unit_unrolled<false>(in, k, [&](uint64_t kmer, int j, bool) {
            const uint32_t v = br[j];
            if (v != 0xFFFFFFFFu) {
                const uint32_t b = v & ((1u << kScatterRankShift) - 1u);
                const uint32_t at = (uint32_t)s_start[b] + (v >> kScatterRankShift);
                stage[at] = kmer;
                stage_b[at] = (uint16_t)b;
            }
        });
    }
#pragma unroll
    for (uint32_t j = 0; j < kScatterPer; j++) {
        const uint32_t i = b0 + j;
        if (j < per && i < n_buckets) {
            if (kCapped) {
                const uint32_t room = g[j] >= lay.cap ? 0u : (uint32_t)min((unsigned long long)c_[j], lay.cap - g[j]);
                s_g[i] = (uint32_t)min(g[j], lay.cap);
                s_room[i] = (uint16_t)room;
                s_obase[i] = c_[j] > room
                                 ? (uint32_t)min(atomicAdd(&cursors[n_buckets], (unsigned long long)(c_[j] - room)),
                                                 0xffff0000ull)
                                 : 0u;
            } else {
                s_gbase[i] = g[j];
            }
        }
    }
    __syncthreads();
    // copy-out: consecutive threads write consecutive cells of a bucket run
    for (uint32_t p = threadIdx.x; p < total; p += blockDim.x) {
        const unsigned long long kmer = stage[p];
        const uint32_t b = stage_b[p];
        const uint32_t rel = p - (uint32_t)s_start[b];
        if (!kCapped) {
            out[s_gbase[b] + rel] = kmer;
        } else if (rel < (uint32_t)s_room[b]) {
            out[(unsigned long long)b * lay.cap + s_g[b] + rel] = kmer;
        } else {
            const unsigned long long o = (unsigned long long)s_obase[b] + (rel - (uint32_t)s_room[b]);
            if (o < lay.ovf_cap) out[lay.ovf_base + o] = kmer;  // else: dropped, the host re-buckets the batch
        }
    }
    if (kCapped && threadIdx.x == 0 && total) atomicAdd(&cc->n_windows, (unsigned long long)total);
}

// ---------------------------------------------------------------------------
// (5) tile-sorted lists + shared-memory insert ("tiled" mode; the bulk path)
//
//   pass A  bucket_scatter_kernel above: k-mers grouped by (owner, table region) — <= 1024 buckets.
//   pass B  tile_sort_kernel: every bucket's cells are cut into tiles of kTile k-mers; each tile
//           is sorted IN PLACE by the next g2 hash bits (the "sub-bucket") inside shared memory,
//           and the tile's F+1 sub-bucket offsets are written beside it.  Reads and writes are whole
//           64 KiB tiles: pure streaming, 8 B in + 8 B out per k-mer.
//   insert  tile_insert_kernel: one CTA per table partition (kPartSlots slots = 64 KiB).  The
//           partition's k-mers are, in every tile of its bucket, one contiguous run (sub-buckets
//           are ordered by hash).  The CTA loads the partition into shared memory, walks the runs
//           chunk by chunk (a CTA barrier between chunks makes the per-chunk histogram columns
//           exact), counts with shared-memory atomics, and writes the partition back.
//           DRAM traffic: 16 B/slot in + 16 B/slot out + 8 B per k-mer, all sequential.
//
//   The list geometry (g1 region bits, g2 sub-bucket bits) is fixed when a list is built and
//   does not depend on the table size at insert time: a partition of a smaller table covers
//   several adjacent sub-buckets or whole buckets (still one run per tile), a partition of a
//   larger table shares a sub-bucket with its neighbours and filters by home slot.
// ---------------------------------------------------------------------------

static constexpr uint32_t kTileLog2 = 13;
static constexpr uint32_t kTile = 1u << kTileLog2;            // k-mers one CTA sorts (64 KiB)
static constexpr uint32_t kMaxTileLog2 = 16;                  // a cluster of 8 CTAs sorts 65536 k-mers as ONE tile
typedef uint32_t tile_off_t;                                  // sub-bucket offsets inside a tile (<= 2^16 inclusive)
static constexpr uint32_t kSortThreads = 512;
static constexpr uint32_t kSortPer = kTile / kSortThreads;    // k-mers per thread
// (the sort kernels are compiled for 2 and for 4 sub-buckets per thread in their scans: up to 2^10, and 2^11 sub-buckets)
static constexpr uint32_t kMaxSubLog2 = 11;                   // g2 <= 11: sub-bucket and rank (< 2^14) share a 32-bit word
static_assert(kTile <= (1u << 14), "rank must fit 14 bits");

struct ListGeom {
    uint32_t n_ranks;
    uint32_t g1;  // log2(table regions per owner): bucket = owner << g1 | region
    uint32_t g2;  // log2(sub-buckets per region)
    __host__ __device__ __forceinline__ uint32_t sub(uint64_t kmer) const {
        const uint64_t h = skm_hash_kmer(kmer);
        const uint64_t lh = n_ranks == 1 ? h : skm_local_hash(h, n_ranks);
        return g2 ? (uint32_t)((lh << g1) >> (64u - g2)) : 0u;
    }
};

// Where the tiles of a bucketed list are (device arrays, one set per list):
//   cell_begin[b]  first cell of bucket b          bucket_n[b]  k-mers in bucket b
//   tile_begin[b]  index of bucket b's first tile; tile_begin[nb] = number of tiles
// Tile t of bucket b covers cells [cell_begin[b] + j*T, +min(T, bucket_n[b] - j*T)), T = 2^tile_log2 (a
// property of the list: kTile when one CTA sorts a tile, up to 2^kMaxTileLog2 when a cluster does),
// j = t - tile_begin[b]; a tile past the end of its bucket is empty (capped layout: every bucket
// owns the same number of tile slots).
struct ListMeta {
    unsigned long long *cell_begin;
    uint32_t *tile_begin;
    uint32_t *bucket_n;
};
__host__ __device__ inline size_t list_meta_words(uint32_t nb) { return (size_t)nb + (2 * (size_t)nb + 2) / 2 + 1; }
__host__ __device__ inline ListMeta list_meta_at(unsigned long long *base, uint32_t nb) {
    ListMeta m;
    m.cell_begin = base;
    m.tile_begin = reinterpret_cast<uint32_t *>(base + nb);
    m.bucket_n = m.tile_begin + nb + 1;
    return m;
}

// src: capped layout (cap > 0) = the scatter kernel's per-bucket totals; exact layout (cap == 0) =
// the nb+1 bucket offsets.  One CTA of 1024 threads.
__device__ __forceinline__ void tile_plan_body(const unsigned long long *__restrict__ src, uint32_t nb,
                                               unsigned long long cap, uint32_t tiles_per_bucket, ListMeta m,
                                               uint32_t tile_log2);
__global__ void __launch_bounds__(1024)
tile_plan_kernel(const unsigned long long *__restrict__ src, uint32_t nb, unsigned long long cap,
                 uint32_t tiles_per_bucket, ListMeta m, uint32_t tile_log2) {
    tile_plan_body(src, nb, cap, tiles_per_bucket, m, tile_log2);
}

// The lists of all owners of one batch (multi-GPU sender): one launch serves them all.
static constexpr uint32_t kMaxOwners = 16;
struct OwnerArrays {
    unsigned long long *list[kMaxOwners];
    unsigned long long *meta[kMaxOwners];   // list_meta_words(nb) words each
    tile_off_t *tile_off[kMaxOwners];
};

// CTA o plans owner o's list from the o-th array of nb + 1 totals / offsets.
__global__ void __launch_bounds__(1024)
tile_plan_owners_kernel(const unsigned long long *__restrict__ src, uint32_t nb, unsigned long long cap,
                        uint32_t tiles_per_bucket, OwnerArrays oa) {
    tile_plan_body(src + (size_t)blockIdx.x * (nb + 1), nb, cap, tiles_per_bucket, list_meta_at(oa.meta[blockIdx.x], nb), kTileLog2);
}

__device__ __forceinline__ void tile_plan_body(const unsigned long long *__restrict__ src, uint32_t nb,
                                               unsigned long long cap, uint32_t tiles_per_bucket, ListMeta m,
                                               uint32_t tile_log2) {
    __shared__ uint32_t s_warp[32];
    const uint32_t b = threadIdx.x;
    uint32_t tiles = 0;
    unsigned long long n = 0;
    if (b < nb) {
        if (cap) {
            n = src[b] < cap ? src[b] : cap;
            m.cell_begin[b] = (unsigned long long)b * cap;
            tiles = tiles_per_bucket;
        } else {
            n = src[b + 1] - src[b];
            m.cell_begin[b] = src[b];
            tiles = (uint32_t)((n + (1ull << tile_log2) - 1) >> tile_log2);
        }
        m.bucket_n[b] = (uint32_t)n;
    }
    uint32_t incl = tiles;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= (uint32_t)o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t off = 0;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) off += s_warp[w];
    if (b < nb) m.tile_begin[b] = off + incl - tiles;
    if (b == nb - 1 || (nb == 0 && b == 0)) m.tile_begin[nb] = nb ? off + incl : 0;
}

// tile_off[t * (F + 1) + f] = first cell (relative to the tile) of sub-bucket f; [.. + F] = tile length
template <uint32_t kSubPer>
__device__ __forceinline__ void tile_sort_body(unsigned long long *__restrict__ list, ListMeta m, uint32_t nb,
                                               ListGeom geom, tile_off_t *__restrict__ tile_off, uint32_t t);
template <uint32_t kSubPer>   // sub-buckets per thread in the scan: 2 (g2 <= 10) or 4 (g2 = 11)
__global__ void __launch_bounds__(kSortThreads, 2)
tile_sort_kernel(unsigned long long *__restrict__ list, ListMeta m, uint32_t nb, ListGeom geom,
                 tile_off_t *__restrict__ tile_off) {
    tile_sort_body<kSubPer>(list, m, nb, geom, tile_off, blockIdx.x);
}
// grid (tile slots per owner, owners)
__global__ void __launch_bounds__(kSortThreads, 2)
tile_sort_owners_kernel(OwnerArrays oa, uint32_t nb, ListGeom geom) {
    const uint32_t o = blockIdx.y;
    tile_sort_body<2>(oa.list[o], list_meta_at(oa.meta[o], nb), nb, geom, oa.tile_off[o], blockIdx.x);   // owner lists: g2 <= 10
}

template <uint32_t kSubPer>
__device__ __forceinline__ void tile_sort_body(unsigned long long *__restrict__ list, ListMeta m, uint32_t nb,
                                               ListGeom geom, tile_off_t *__restrict__ tile_off, uint32_t t) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned long long *stage = reinterpret_cast<unsigned long long *>(s_raw);   // kTile
    uint32_t *cnt = reinterpret_cast<uint32_t *>(stage + kTile);                  // F (+1)
    __shared__ uint32_t s_warp[kSortThreads / 32];
    const uint32_t F = 1u << geom.g2;
    if (t >= m.tile_begin[nb]) return;
    uint32_t lo = 0, hi = nb;  // last bucket whose tile_begin <= t
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (m.tile_begin[mid] <= t) lo = mid; else hi = mid;
    }
    const uint32_t j = t - m.tile_begin[lo];
    const unsigned long long bn = m.bucket_n[lo];
    const unsigned long long first = (unsigned long long)j * kTile;
    const uint32_t n = first >= bn ? 0u : (uint32_t)(bn - first < kTile ? bn - first : kTile);
    tile_off_t *off = tile_off + (size_t)t * (F + 1);
    if (n == 0) {  // empty tile slot: all offsets zero, so every run read from it has length 0
        for (uint32_t f = threadIdx.x; f <= F; f += kSortThreads) off[f] = 0;
        return;
    }
    unsigned long long *cells = list + m.cell_begin[lo] + first;
    for (uint32_t f = threadIdx.x; f < F; f += kSortThreads) cnt[f] = 0;
    __syncthreads();
    unsigned long long km[kSortPer];
    uint32_t fr[kSortPer];  // sub-bucket | rank inside (tile, sub-bucket) << 10
#pragma unroll
    for (uint32_t r = 0; r < kSortPer; r++) {
        const uint32_t i = threadIdx.x + r * kSortThreads;
        km[r] = i < n ? cells[i] : 0ull;
    }
#pragma unroll
    for (uint32_t r = 0; r < kSortPer; r++) {
        const uint32_t i = threadIdx.x + r * kSortThreads;
        if (i < n) {
            const uint32_t f = geom.sub(km[r]);
            fr[r] = f | (atomicAdd(&cnt[f], 1u) << kMaxSubLog2);
        }
    }
    __syncthreads();
    // exclusive scan of cnt[0..F) (F <= 2^kMaxSubLog2 = kSubPer entries per thread)
    const uint32_t a = threadIdx.x * kSubPer;
    uint32_t cj[kSubPer];
    uint32_t mine = 0;
#pragma unroll
    for (uint32_t j = 0; j < kSubPer; j++) {
        cj[j] = a + j < F ? cnt[a + j] : 0u;
        mine += cj[j];
    }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= (uint32_t)o) incl += v;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) base += s_warp[w];
    uint32_t run = base + incl - mine;
#pragma unroll
    for (uint32_t j = 0; j < kSubPer; j++) {
        if (a + j < F) {
            cnt[a + j] = run;
            off[a + j] = run;
        }
        run += cj[j];
    }
    if (threadIdx.x == 0) off[F] = n;
    __syncthreads();
#pragma unroll
    for (uint32_t r = 0; r < kSortPer; r++) {
        const uint32_t i = threadIdx.x + r * kSortThreads;
        if (i < n) stage[cnt[fr[r] & ((1u << kMaxSubLog2) - 1)] + (fr[r] >> kMaxSubLog2)] = km[r];
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += kSortThreads) cells[i] = stage[i];
}

// Pass B with a CLUSTER per tile: C * kTile k-mers are sorted as ONE tile by the C CTAs of a
// thread-block cluster.  Each CTA loads its kTile cells and ranks its k-mers per sub-bucket in its
// own shared memory (as tile_sort_kernel does); after a cluster barrier every CTA reads the others'
// counts through distributed shared memory: the tile-wide start of every sub-bucket (a scan of the
// summed counts) plus the k-mers that lower-ranked CTAs hold for the same sub-bucket give the first
// cell of this CTA's piece.  The CTA lays its k-mers out by sub-bucket in its stage and copies every
// piece to its place in the tile — in place: nobody writes before everybody has loaded (the barrier).
// Why: a tile C times larger gives the insert kernel runs C times longer, or C times more
// sub-buckets at the same run length — which is what lets a multi-GPU sender sort an owner's coarse
// bucket straight down to that owner's table partitions, without a re-bucketing pass.
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int C, uint32_t kSubPer>
__global__ void __launch_bounds__(kSortThreads, 2)
tile_sort_cluster_kernel(unsigned long long *__restrict__ list, ListMeta m, uint32_t nb, ListGeom geom,
                         tile_off_t *__restrict__ tile_off, unsigned int *__restrict__ tile_counter) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char s_raw[];
    constexpr uint32_t kMaxF = 1u << kMaxSubLog2;
    unsigned long long *stage = reinterpret_cast<unsigned long long *>(s_raw);   // kTile
    uint32_t *cnt = reinterpret_cast<uint32_t *>(stage + kTile);                  // this CTA's counts (the cluster reads them)
    uint32_t *lstart = cnt + kMaxF;                                               // first stage slot of sub-bucket f
    uint32_t *gbase = lstart + kMaxF;                                             // first tile cell of this CTA's piece of f
    uint16_t *stage_d = reinterpret_cast<uint16_t *>(gbase + kMaxF);              // kTile: tile cell every staged k-mer goes to
    __shared__ unsigned long long s_warp[kSortThreads / 32];
    __shared__ uint32_t s_next[2];   // (CTA 0 of the cluster) the cluster's next tile, double-buffered
    constexpr uint32_t T = (uint32_t)C * kTile;
    static_assert(T <= 65536u, "tile-relative cell indices are kept in 16 bits");
    const uint32_t F = 1u << geom.g2;
    const uint32_t cr = cluster.block_rank();
    const uint32_t n_tiles = m.tile_begin[nb];
    const uint32_t n_clusters = gridDim.x / C;
    // PERSISTENT clusters: as many as the chip holds; cluster g starts with tile g and then takes tiles from a
    // counter.  (One cluster per tile left CTA slots idle: a new cluster starts only when C slots of one GPC are
    // free at the same time, and the CTAs of a cluster leave together.)
    uint32_t t = blockIdx.x / C;
    for (uint32_t it = 0;; it++) {
        if (t >= n_tiles) {   // (the same in every CTA of the cluster)
            cluster.sync();   // nobody leaves while s_next of CTA 0 may still be read
            break;
        }
        uint32_t lo = 0, hi = nb;  // last bucket whose tile_begin <= t
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (m.tile_begin[mid] <= t) lo = mid; else hi = mid;
        }
        const uint32_t j = t - m.tile_begin[lo];
        const unsigned long long bn = m.bucket_n[lo];
        const unsigned long long first = (unsigned long long)j * T;
        // (an empty tile slot takes the same path with zero cells everywhere: all its offsets come out zero)
        const uint32_t n_tile = first >= bn ? 0u : (uint32_t)(bn - first < T ? bn - first : T);
        tile_off_t *off = tile_off + (size_t)t * (F + 1);
        const uint32_t my0 = cr * kTile;
        const uint32_t n = my0 >= n_tile ? 0u : (n_tile - my0 < kTile ? n_tile - my0 : kTile);
        unsigned long long *tile_cells = list + m.cell_begin[lo] + first;
        const unsigned long long *cells = tile_cells + my0;
        // (cnt was last read by the cluster before the arrive that the wait at the end of the previous trip matched)
        for (uint32_t f = threadIdx.x; f < F; f += kSortThreads) cnt[f] = 0;
        __syncthreads();
        unsigned long long km[kSortPer];
        uint32_t fr[kSortPer];  // sub-bucket | rank inside (CTA, sub-bucket) << 10
#pragma unroll
        for (uint32_t r = 0; r < kSortPer; r++) {
            const uint32_t i = threadIdx.x + r * kSortThreads;
            km[r] = i < n ? cells[i] : 0ull;
        }
#pragma unroll
        for (uint32_t r = 0; r < kSortPer; r++) {
            const uint32_t i = threadIdx.x + r * kSortThreads;
            if (i < n) {
                const uint32_t f = geom.sub(km[r]);
                fr[r] = f | (atomicAdd(&cnt[f], 1u) << kMaxSubLog2);
            }
        }
        cluster.sync();   // every CTA's counts are complete, and every CTA holds its cells in registers
        if (cr == 0 && threadIdx.x == 0) s_next[it & 1] = n_clusters + atomicAdd(tile_counter, 1u);
        // sub-buckets a .. a + kSubPer - 1 of this thread: tile-wide totals, and what lower-ranked CTAs hold
        // (parked in this thread's entries of lstart / gbase until the scan is done: no registers held across it)
        const uint32_t a = threadIdx.x * kSubPer;
        uint32_t tot_sum = 0, own_sum = 0;
        if (a < F) {   // (F is a power of two: a thread's kSubPer sub-buckets are all inside or all outside, F >= kSubPer,
                       //  or only the first F of thread 0's are inside)
            uint32_t tot[kSubPer], bef[kSubPer];
#pragma unroll
            for (uint32_t j = 0; j < kSubPer; j++) tot[j] = bef[j] = 0;
            // all remote loads first (independent), then the stores
#pragma unroll
            for (uint32_t c = 0; c < (uint32_t)C; c++) {
                const uint32_t *rc = cluster.map_shared_rank(cnt, c);
#pragma unroll
                for (uint32_t j = 0; j < kSubPer; j++) {
                    const uint32_t x = a + j < F ? rc[a + j] : 0u;
                    tot[j] += x;
                    if (c < cr) bef[j] += x;
                }
            }
#pragma unroll
            for (uint32_t j = 0; j < kSubPer; j++) {
                if (a + j < F) {
                    lstart[a + j] = tot[j];
                    gbase[a + j] = bef[j];
                    tot_sum += tot[j];
                    own_sum += cnt[a + j];
                }
            }
        }
        cluster_arrive();   // this CTA has read the others' counts; CTA 0 has published the next tile
        // ONE exclusive scan for both: high word = tile-wide totals, low word = this CTA's counts
        const unsigned long long mine = ((unsigned long long)tot_sum << 32) | (unsigned long long)own_sum;
        unsigned long long incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= (uint32_t)o) incl += v;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned long long base = 0;
        for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) base += s_warp[w];
        const unsigned long long start = base + incl - mine;
        uint32_t gs = (uint32_t)(start >> 32), ls = (uint32_t)start;
#pragma unroll
        for (uint32_t j = 0; j < kSubPer; j++) {
            if (a + j < F) {
                const uint32_t tot = lstart[a + j], bef = gbase[a + j];
                lstart[a + j] = ls;
                gbase[a + j] = gs + bef;
                if (cr == 0) off[a + j] = gs;
                gs += tot;
                ls += cnt[a + j];
            }
        }
        if (cr == 0 && threadIdx.x == 0) off[F] = n_tile;
        __syncthreads();
#pragma unroll
        for (uint32_t r = 0; r < kSortPer; r++) {
            const uint32_t i = threadIdx.x + r * kSortThreads;
            if (i < n) {
                const uint32_t f = fr[r] & ((1u << kMaxSubLog2) - 1), rank = fr[r] >> kMaxSubLog2;
                const uint32_t at = lstart[f] + rank;
                stage[at] = km[r];
                stage_d[at] = (uint16_t)(gbase[f] + rank);
            }
        }
        __syncthreads();
        // consecutive threads write consecutive cells of a piece
        for (uint32_t p = threadIdx.x; p < n; p += kSortThreads) tile_cells[stage_d[p]] = stage[p];
        cluster_wait();   // everybody has read this CTA's counts; s_next of CTA 0 is visible
        t = *cluster.map_shared_rank(&s_next[it & 1], 0);
        __syncthreads();  // (s_warp, lstart, gbase and the stage are rewritten by the next trip)
    }
}
__host__ __device__ inline size_t tile_sort_cluster_smem_bytes() {
    return (size_t)kTile * 10 + ((size_t)3 << kMaxSubLog2) * 4 + 16;
}

// Multi-GPU senders: pass A buckets a batch by (owner, coarse region) — at most 1024 buckets in
// all, so only 1024 / n_ranks regions per owner — and a table partition of an owner would find its
// k-mers scattered over n_ranks times more, n_ranks times shorter runs than on one GPU.  This
// kernel re-buckets ONE owner's slice of such a list into that owner's kFineRegions table
// regions (tile by tile: a tile of a coarse bucket only feeds kFineRegions / coarse regions
// fine buckets, so its k-mers leave in long runs).  The result is a list with exactly the
// geometry a single GPU builds for itself; it is tile-sorted and shipped as a whole.
//   MODE 0: capped output (bucket f owns cells [f * cap, (f + 1) * cap); what does not fit is counted in
//           cursors[kFineRegions] and dropped: the host rebuilds the batch exactly)
//   MODE 1: count only (cursors[f] += k-mers of fine bucket f)
//   MODE 2: exact output (cursors[] start at the bucket offsets)
static constexpr uint32_t kFineLog2 = 10;
static constexpr uint32_t kFineRegions = 1u << kFineLog2;

template <int MODE>
__global__ void __launch_bounds__(kSortThreads, 2)
tile_rebucket_kernel(const unsigned long long *__restrict__ in_list, ListMeta m, uint32_t nb_in, uint32_t log2_coarse,
                     uint32_t n_ranks, unsigned long long *__restrict__ cursors_all, OwnerArrays oa,
                     unsigned long long cap) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned long long *stage = reinterpret_cast<unsigned long long *>(s_raw);     // kTile
    unsigned long long *s_g = stage + kTile;                                        // kFineRegions: first cell of this tile's run
    uint32_t *cnt = reinterpret_cast<uint32_t *>(s_g + kFineRegions);               // kFineRegions: counts, then stage starts
    uint32_t *room = cnt + kFineRegions;                                            // kFineRegions (capped)
    uint16_t *stage_f = reinterpret_cast<uint16_t *>(room + kFineRegions);          // kTile: fine bucket of every staged k-mer
    __shared__ uint32_t s_warp[kSortThreads / 32];
    // every tile of the coarse list; its bucket says whose k-mers it holds (bucket = owner << log2_coarse | region)
    const uint32_t t = blockIdx.x;
    if (t >= m.tile_begin[nb_in]) return;
    uint32_t lo = 0, hi = nb_in;  // last bucket whose tile_begin <= t
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (m.tile_begin[mid] <= t) lo = mid; else hi = mid;
    }
    const uint32_t owner = lo >> log2_coarse;
    unsigned long long *cursors = cursors_all + (size_t)owner * (kFineRegions + 1);
    unsigned long long *out = oa.list[owner];
    const uint32_t j = t - m.tile_begin[lo];
    const unsigned long long bn = m.bucket_n[lo];
    const unsigned long long first = (unsigned long long)j * kTile;
    const uint32_t n = first >= bn ? 0u : (uint32_t)(bn - first < kTile ? bn - first : kTile);
    if (n == 0) return;
    const unsigned long long *cells = in_list + m.cell_begin[lo] + first;
    for (uint32_t f = threadIdx.x; f < kFineRegions; f += kSortThreads) cnt[f] = 0;
    __syncthreads();
    unsigned long long km[kSortPer];
    uint32_t fr[kSortPer];
#pragma unroll
    for (uint32_t r = 0; r < kSortPer; r++) {
        const uint32_t i = threadIdx.x + r * kSortThreads;
        km[r] = i < n ? cells[i] : 0ull;
    }
    // A tile of one coarse bucket feeds only kFineRegions / (coarse regions per owner) = n_ranks fine buckets, so
    // the lanes of a warp collide on a handful of counters: the warp groups its lanes by fine bucket
    // (match.any) and one lane per group reserves the group's ranks with a single shared-memory atomic.
    const uint32_t lane = threadIdx.x & 31;
#pragma unroll
    for (uint32_t r = 0; r < kSortPer; r++) {
        const uint32_t i = threadIdx.x + r * kSortThreads;
        uint32_t f = 0xffffffffu;   // lanes past the end of the tile form a group of their own
        if (i < n) f = (uint32_t)(skm_local_hash(skm_hash_kmer(km[r]), n_ranks) >> (64u - kFineLog2));
        const uint32_t peers = __match_any_sync(0xffffffffu, f);
        const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
        uint32_t base = 0;
        if (lane == leader && i < n) base = atomicAdd(&cnt[f], (uint32_t)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        fr[r] = f | ((base + (uint32_t)__popc(peers & ((1u << lane) - 1u))) << kMaxSubLog2);
    }
    __syncthreads();
    const uint32_t a = threadIdx.x * 2;   // kFineRegions = 2 * kSortThreads
    const uint32_t c0 = cnt[a], c1 = cnt[a + 1];
    if (MODE == 1) {
        if (c0) atomicAdd(&cursors[a], (unsigned long long)c0);
        if (c1) atomicAdd(&cursors[a + 1], (unsigned long long)c1);
        return;
    }
    // reserve this tile's run in every fine bucket it feeds, and lay the k-mers out in bucket order in the stage
    const unsigned long long g0 = c0 ? atomicAdd(&cursors[a], (unsigned long long)c0) : 0ull;
    const unsigned long long g1 = c1 ? atomicAdd(&cursors[a + 1], (unsigned long long)c1) : 0ull;
    uint32_t incl = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= (uint32_t)o) incl += v;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) base += s_warp[w];
    const uint32_t start0 = base + incl - c0 - c1;
    cnt[a] = start0;
    cnt[a + 1] = start0 + c0;
    if (MODE == 0) {
        const uint32_t r0 = g0 >= cap ? 0u : (uint32_t)(c0 < cap - g0 ? c0 : cap - g0);
        const uint32_t r1 = g1 >= cap ? 0u : (uint32_t)(c1 < cap - g1 ? c1 : cap - g1);
        room[a] = r0;
        room[a + 1] = r1;
        if (c0 + c1 > r0 + r1) atomicAdd(&cursors[kFineRegions], (unsigned long long)(c0 + c1 - r0 - r1));
        s_g[a] = (unsigned long long)a * cap + (g0 < cap ? g0 : cap);
        s_g[a + 1] = (unsigned long long)(a + 1) * cap + (g1 < cap ? g1 : cap);
    } else {
        s_g[a] = g0;
        s_g[a + 1] = g1;
    }
    __syncthreads();
#pragma unroll
    for (uint32_t r = 0; r < kSortPer; r++) {
        const uint32_t i = threadIdx.x + r * kSortThreads;
        if (i < n) {
            const uint32_t f = fr[r] & ((1u << kMaxSubLog2) - 1);
            const uint32_t at = cnt[f] + (fr[r] >> kMaxSubLog2);
            stage[at] = km[r];
            stage_f[at] = (uint16_t)f;
        }
    }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < n; p += kSortThreads) {
        const unsigned long long kmer = stage[p];
        const uint32_t f = stage_f[p];
        const uint32_t rel = p - cnt[f];
        if (MODE == 2 || rel < room[f]) out[s_g[f] + rel] = kmer;   // (capped: the rest was counted as overflow)
    }
}
__host__ __device__ inline size_t tile_rebucket_smem_bytes() { return (size_t)kTile * 10 + (size_t)kFineRegions * 16 + 16; }

__host__ __device__ inline size_t tile_sort_smem_bytes(uint32_t g2) { return (size_t)kTile * 8 + ((size_t)1 << g2) * 4 + 16; }

// One bucketed, tile-sorted list as the insert kernel sees it.  `list` and `tile_off` are biased by
// the receiver so that the sender's own numbering (cell_begin / tile_begin of bucket0 + region)
// indexes them directly: on one GPU bucket0 = 0 and nothing is biased; across GPUs the list is the
// owner's block inside a receive arena.
struct SegDesc {
    const unsigned long long *list;
    const tile_off_t *tile_off;
    const uint32_t *tile_begin;
    const unsigned long long *cell_begin;
    uint32_t bucket0;   // first bucket of this table's owner in the list's numbering (rank << g1)
    uint32_t chunk;     // chunk index relative to the launch's first chunk
    uint32_t rel;       // 1: the arrays are an owner's SLICE of the sender's metadata (bucket0 = 0) and the
                        //    list / tile_off pointers address the slice: subtract the slice's first entries
    uint32_t pad;
};

#ifndef SKM_INS_THREADS
#define SKM_INS_THREADS 512
#endif
#ifndef SKM_INS_CTAS
#define SKM_INS_CTAS 2
#endif
static constexpr uint32_t kInsThreads = SKM_INS_THREADS;
static constexpr uint32_t kInsCtasPerSm = SKM_INS_CTAS;
static constexpr uint32_t kMaxVseg = 256;    // (list, bucket) pairs per launch that one partition reads
static constexpr uint32_t kStagedRuns = 256; // runs whose descriptors are staged in shared memory at a time
#ifndef SKM_INS_STAGE
#define SKM_INS_STAGE 1024
#endif
#ifndef SKM_INS_DEPTH
#define SKM_INS_DEPTH 4
#endif
static constexpr uint32_t kStageCap = SKM_INS_STAGE;    // k-mers per staged span
static constexpr uint32_t kStageDepth = SKM_INS_DEPTH;  // spans in flight (being copied or counted), 2..4
static_assert(kStageDepth >= 2 && kStageDepth <= 4, "cp.async wait depth");
static constexpr uint32_t kMaxSpans = 48;    // spans planned at a time
static constexpr uint32_t kHlogCap = 512;    // pending moves of histogram bins >= k_low, per partition
static constexpr uint32_t kBigCount = 0x80000000u;

struct InsertLaunch {
    Slot *table;
    uint32_t log2cap, n_ranks;
    uint32_t g1, g2;
    uint32_t tile_log2;                  // the lists' tiles hold 2^tile_log2 cells
    const SegDesc *segs;                 // sorted by chunk
    const uint32_t *chunk_first_seg;     // [n_chunks + 1]
    uint32_t n_segs, n_chunks;
    uint32_t k_low;                      // histogram bins kept in shared memory (per chunk)
    uint32_t max_occupied;               // a partition holding more keys than this fails (grow + retry)
    const uint32_t *part_ids;            // null: partitions 0 .. n_parts-1
    unsigned long long n_parts;
    int fresh;                           // table logically empty: partitions are not loaded
    unsigned long long *g_delta;         // [n_chunks][histo_max + 2] histogram moves per chunk
    unsigned long long *g_recount;       // [histo_max + 2] histogram of the partitions written back
    unsigned long long histo_max;
    GlobalCounters *gc;
    HistoTotals *tot;                    // n_distinct / n_kmers / n_saturated of the partitions written back
    unsigned long long *part_counter;
    uint32_t *fail_list;
    uint32_t fail_cap;
    const unsigned long long *abort_if;  // speculative launch: non-null and != 0 on the device => the launch does nothing
};

// Speculative insert launches (skm_finalize queues the insert behind the lists' events before the host has seen
// their bucket totals): a capped list that overflowed makes the launch a no-op — the flag is set on the device
// from the overflow words the bucketing kernels left in pinned host memory — and the host takes the slow path.
static constexpr uint32_t kMaxGuardWords = 64;
struct GuardWords {
    const unsigned long long *w[kMaxGuardWords];
};
__global__ void spec_guard_kernel(GuardWords g, uint32_t n, unsigned long long *flag) {
    unsigned long long any = 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) any |= *reinterpret_cast<const volatile unsigned long long *>(g.w[i]);
    any = __reduce_or_sync(0xffffffffu, (unsigned)(any != 0));
    if (threadIdx.x == 0) *flag = any;
}

__host__ __device__ inline size_t tile_insert_smem_bytes(uint32_t n_chunks, uint32_t k_low, bool histo) {
    size_t b = (size_t)kPartSlots * 12;                    // keys + counts
    b += (size_t)kMaxVseg * 16 + 8;                        // vs_cell (u64), vs_tb, vs_first (u32)
    b += (size_t)kStagedRuns * 16 + 8;                     // run_src (u64), run_len, run_pos (u32)
    b += (size_t)kStageDepth * kStageCap * 8;              // staged spans of k-mers
    if (histo) b += (size_t)n_chunks * k_low * 8 + (size_t)k_low * 4 + (size_t)kHlogCap * 4;  // chist + phist, fhist, hlog
    else b += (size_t)k_low * 4;
    return b + 64;
}

// A histogram move of a bin >= k_low: logged per partition (so that a failed partition can be rolled
// back) and published when the partition commits; a full log publishes at once and marks the
// partition dirty.
__device__ __forceinline__ void hlog_push(uint32_t *hlog, uint32_t *n, uint32_t *dirty, const InsertLaunch &L, uint32_t c,
                                          uint32_t bin, bool minus) {
    const uint32_t e = atomicAdd(n, 1u);
    if (e < kHlogCap) {
        hlog[e] = (c << 24) | (minus ? 0x800000u : 0u) | bin;
    } else {
        *dirty = 1;
        atomicAdd(&L.g_delta[(size_t)c * (L.histo_max + 2) + bin], minus ? ~0ull : 1ull);
    }
}

template <bool kHisto>
__global__ void __launch_bounds__(kInsThreads, kInsCtasPerSm)
tile_insert_kernel(const InsertLaunch L) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(s_raw);
    unsigned long long *stage = keys + kPartSlots;           // kStageDepth * kStageCap
    unsigned long long *vs_cell = stage + kStageDepth * kStageCap;
    unsigned long long *run_src = vs_cell + kMaxVseg;
    uint32_t *counts = reinterpret_cast<uint32_t *>(run_src + kStagedRuns);
    uint32_t *vs_tb = counts + kPartSlots;
    uint32_t *vs_first = vs_tb + kMaxVseg;          // kMaxVseg + 1
    uint32_t *run_len = vs_first + kMaxVseg + 2;
    uint32_t *run_pos = run_len + kStagedRuns;      // kStagedRuns + 1: first k-mer of each staged run, window-relative
    int *fhist = reinterpret_cast<int *>(run_pos + kStagedRuns + 2);   // k_low: histogram of the written-back partition(s)
    int *chist = fhist + L.k_low;                   // n_chunks * k_low: moves of committed partitions
    int *phist = chist + (kHisto ? L.n_chunks * L.k_low : 0);   // n_chunks * k_low: moves of the current partition
    uint32_t *hlog = reinterpret_cast<uint32_t *>(phist + (kHisto ? L.n_chunks * L.k_low : 0));
    __shared__ unsigned long long s_q;
    __shared__ uint32_t s_occ, s_fail, s_big, s_hlog_n, s_dirty;
    __shared__ uint32_t s_span[kMaxSpans][5];     // the plan: chunk, first run, end run, first k-mer, k-mers of each span
    __shared__ uint32_t s_nspans, s_more, s_next_run;

    if (L.abort_if && *L.abort_if) return;   // (speculative launch called off: the same in every thread)
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = kInsThreads / 32;
    uint32_t keys_s = (uint32_t)__cvta_generic_to_shared(keys), counts_s = (uint32_t)__cvta_generic_to_shared(counts);
    uint32_t phist_s = (uint32_t)__cvta_generic_to_shared(phist), stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    // (opaque to the compiler, which otherwise re-derives the window base, S2UR + ULEA, inside the probe loop)
    asm volatile("" : "+r"(keys_s), "+r"(counts_s), "+r"(phist_s), "+r"(stage_s));
    const uint32_t home_shift = 64u - L.log2cap;   // log2cap >= kPartLog2: a table has at least one partition
    const uint32_t pbits = L.log2cap - kPartLog2;
    const uint32_t F = 1u << L.g2;
    const unsigned long long top = L.histo_max + 1;
    for (uint32_t i = threadIdx.x; i < L.k_low; i += kInsThreads) fhist[i] = 0;
    if (kHisto)
        for (uint32_t i = threadIdx.x; i < 2 * L.n_chunks * L.k_low; i += kInsThreads) chist[i] = 0;
    unsigned long long n_new_cta = 0, n_kmers_cta = 0, n_dist_cta = 0, n_sat_cta = 0;

    // (thread 0) the partition counter is read one partition ahead: the atomic's round trip to L2 is over
    // long before the value is needed, instead of holding the whole CTA at the top of every trip
    unsigned long long next_i = 0;
    if (threadIdx.x == 0) next_i = atomicAdd(L.part_counter, 1ull);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long i = next_i;
            s_q = i < L.n_parts ? (L.part_ids ? (unsigned long long)L.part_ids[i] : i) : ~0ull;
            s_occ = 0;
            s_fail = 0;
            s_big = 0;
            s_hlog_n = 0;
            s_dirty = 0;
            if (i < L.n_parts) next_i = atomicAdd(L.part_counter, 1ull);
        }
        __syncthreads();
        const unsigned long long q = s_q;
        if (q == ~0ull) break;
        // which buckets / sub-buckets of the lists hold this partition's k-mers
        uint32_t b_lo, nbr, f0, f1;
        bool filter = false;
        if (pbits <= L.g1) {
            nbr = 1u << (L.g1 - pbits);
            b_lo = (uint32_t)q << (L.g1 - pbits);
            f0 = 0;
            f1 = F;
        } else if (pbits <= L.g1 + L.g2) {
            const uint32_t d = pbits - L.g1, e = L.g1 + L.g2 - pbits;
            nbr = 1;
            b_lo = (uint32_t)(q >> d);
            f0 = ((uint32_t)q & ((1u << d) - 1)) << e;
            f1 = f0 + (1u << e);
        } else {
            const uint32_t d = pbits - L.g1;
            nbr = 1;
            b_lo = (uint32_t)(q >> d);
            f0 = (uint32_t)(q >> (d - L.g2)) & (F - 1);
            f1 = f0 + 1;
            filter = true;
        }
        Slot *tp = L.table + (q << kPartLog2);
        // ---- load the partition (or start empty): slots [first, first + n) by `nthr` threads (whole warps).
        //      Done in two halves by warps 1.., each beside one of warp 0's two single-warp jobs below (the scan
        //      of the tile counts; the scan of the run lengths + the span plan), which otherwise leave them idle.
        auto load_part = [&](uint32_t first, uint32_t n, uint32_t tid, uint32_t nthr) {
            uint32_t occ = 0, big = 0;
            if (!L.fresh) {
#pragma unroll 4
                for (uint32_t i = first + tid; i < first + n; i += nthr) {
                    const uint4 v = ld_nc_v4(reinterpret_cast<const uint4 *>(tp) + i);
                    const unsigned long long key = ((unsigned long long)v.y << 32) | v.x;
                    keys[i] = key;
                    const bool is_big = v.w != 0u || v.z >= kBigCount;
                    counts[i] = is_big ? kBigCount : v.z;
                    occ += key != SKM_EMPTY_KEY;
                    big |= is_big && key != SKM_EMPTY_KEY;
                }
            } else {
                for (uint32_t i = first + tid; i < first + n; i += nthr) {
                    keys[i] = SKM_EMPTY_KEY;
                    counts[i] = 0;
                }
            }
            occ = __reduce_add_sync(0xffffffffu, occ);
            if (lane == 0 && occ) atomicAdd(&s_occ, occ);
            if (big) s_big = 1;
        };
        bool part_ready = false;   // the second half is in shared memory
        // ---- which tiles: one "virtual segment" per (list, bucket) ----
        const uint32_t n_vseg = L.n_segs * nbr;   // <= kMaxVseg (the host splits launches)
        for (uint32_t v = threadIdx.x; v < n_vseg; v += kInsThreads) {
            const SegDesc &sg = L.segs[v / nbr];
            const uint32_t bkt = sg.bucket0 + b_lo + v % nbr;
            const uint32_t tb = sg.tile_begin[bkt];
            vs_tb[v] = tb - (sg.rel ? sg.tile_begin[0] : 0u);
            vs_first[v] = sg.tile_begin[bkt + 1] - tb;   // tiles; scanned below
            vs_cell[v] = sg.cell_begin[bkt] - (sg.rel ? sg.cell_begin[0] : 0ull);
        }
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the tile counts -> first run of each virtual segment
            uint32_t run = 0;
            for (uint32_t v0 = 0; v0 < n_vseg; v0 += 32) {
                const uint32_t v = v0 + lane;
                const uint32_t x = v < n_vseg ? vs_first[v] : 0u;
                uint32_t incl = x;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= (uint32_t)o) incl += t;
                }
                if (v < n_vseg) vs_first[v] = run + incl - x;
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) vs_first[n_vseg] = run;
        } else {
            load_part(0, kPartSlots / 2, threadIdx.x - 32, kInsThreads - 32);
        }
        __syncthreads();
        const uint32_t total_runs = vs_first[n_vseg];
        unsigned long long n_new = 0;
        const uint32_t hm = (uint32_t)L.histo_max, top32 = hm + 1;
        const uint32_t fast_lim = min(L.k_low, hm + 1);   // old + 1 < fast_lim: both bins are plain shared-memory bins
        // Runs [w0, w1) have their descriptors staged (run_src / run_len / run_pos).  The k-mers are
        // walked in SPANS of <= kStageCap consecutive k-mers of one chunk.  PLAN: warp 0 alone cuts
        // the staged runs into spans (a table in shared memory) — run by every warp, that control
        // code was as many instructions as the counting itself.  EXECUTE: all warps walk the table;
        // the copies of the next kStageDepth-1 spans (cp.async, no registers held) are in flight
        // while a span is counted, and ONE barrier per span both publishes the next span's copies
        // and orders the chunks (chunk c is complete before any k-mer of chunk c+1 is counted).
        uint32_t w0 = 0, w1 = 0;
        // span cursor (meaningful in warp 0 only): chunk, run, k-mers done in the current group
        uint32_t sc = 0, spos = vs_first[L.chunk_first_seg[0] * nbr], soff = 0;
        uint32_t scr1 = vs_first[L.chunk_first_seg[1] * nbr];
        struct Span { uint32_t c, r0, r1, a, n; };
        // 0: span produced; 1: descriptors of run `spos` are not staged; 2: no k-mers left
        auto next_span = [&](Span &sp) -> int {
            for (;;) {
                if (sc >= L.n_chunks) return 2;
                if (spos >= scr1) {
                    sc++;
                    if (sc >= L.n_chunks) return 2;
                    spos = vs_first[L.chunk_first_seg[sc] * nbr];
                    scr1 = vs_first[L.chunk_first_seg[sc + 1] * nbr];
                    soff = 0;
                    continue;
                }
                if (spos >= w1 || spos < w0) return 1;
                const uint32_t gend = min(scr1, w1);
                const uint32_t gbase = run_pos[spos - w0], glen = run_pos[gend - w0] - gbase;
                if (soff >= glen) {
                    spos = gend;
                    soff = 0;
                    continue;
                }
                sp.c = sc;
                sp.r0 = spos;
                sp.r1 = gend;
                sp.a = gbase + soff;
                sp.n = min(kStageCap, glen - soff);
                soff += sp.n;
                return 0;
            }
        };
        auto stage_window = [&](uint32_t from) {
            w0 = from;
            w1 = min(total_runs, w0 + kStagedRuns);
            for (uint32_t r = w0 + threadIdx.x; r < w1; r += kInsThreads) {
                uint32_t lo = 0, hi = n_vseg;  // last virtual segment whose first run <= r
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (vs_first[mid] <= r) lo = mid; else hi = mid;
                }
                const SegDesc &sg = L.segs[lo / nbr];
                const uint32_t jt = r - vs_first[lo];
                const tile_off_t *off = sg.tile_off + (size_t)(vs_tb[lo] + jt) * (F + 1);
                const uint32_t o0 = off[f0], o1 = off[f1];
                run_src[r - w0] = reinterpret_cast<unsigned long long>(sg.list + vs_cell[lo] + ((unsigned long long)jt << L.tile_log2) + o0);
                run_len[r - w0] = o1 - o0;
            }
            __syncthreads();
            if (warp == 0) {
                // run_pos = exclusive scan of run_len over the window
                uint32_t run = 0;
                const uint32_t nw = w1 - w0;
                for (uint32_t v0 = 0; v0 < nw; v0 += 32) {
                    const uint32_t v = v0 + lane;
                    const uint32_t x = v < nw ? run_len[v] : 0u;
                    uint32_t incl = x;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= (uint32_t)o) incl += t;
                    }
                    if (v < nw) run_pos[v] = run + incl - x;
                    run += __shfl_sync(0xffffffffu, incl, 31);
                }
                if (lane == 0) run_pos[nw] = run;
                __syncwarp();
                // the plan: spans of this window, in order.  The usual case — every run of the partition is in this
                // window, at most 32 chunks — is planned by one lane per chunk; else warp 0 walks the runs serially.
                bool planned = false;
                if (w0 == 0 && w1 == total_runs && L.n_chunks <= 32) {
                    uint32_t r0 = 0, r1 = 0, gb = 0, gl = 0, ns_c = 0;
                    if (lane < L.n_chunks) {
                        r0 = vs_first[L.chunk_first_seg[lane] * nbr];
                        r1 = vs_first[L.chunk_first_seg[lane + 1] * nbr];
                        gb = run_pos[r0];
                        gl = run_pos[r1] - gb;
                        ns_c = (gl + kStageCap - 1) / kStageCap;
                    }
                    uint32_t incl = ns_c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= (uint32_t)o) incl += t;
                    }
                    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                    if (total <= kMaxSpans) {
                        uint32_t at = incl - ns_c;
                        for (uint32_t j = 0; j < ns_c; j++, at++) {
                            s_span[at][0] = lane;
                            s_span[at][1] = r0;
                            s_span[at][2] = r1;
                            s_span[at][3] = gb + j * kStageCap;
                            s_span[at][4] = min(kStageCap, gl - j * kStageCap);
                        }
                        if (lane == 0) {
                            s_nspans = total;
                            s_more = 2;
                            s_next_run = total_runs;
                        }
                        planned = true;
                    }
                }
                if (!planned) {
                    uint32_t ns = 0;
                    int st = 0;
                    Span sp;
                    while (ns < kMaxSpans && (st = next_span(sp)) == 0) {
                        if (lane == 0) {
                            s_span[ns][0] = sp.c;
                            s_span[ns][1] = sp.r0;
                            s_span[ns][2] = sp.r1;
                            s_span[ns][3] = sp.a;
                            s_span[ns][4] = sp.n;
                        }
                        ns++;
                    }
                    if (lane == 0) {
                        s_nspans = ns;
                        s_more = ns == kMaxSpans ? 0 : st;   // 0: this window has more spans; 1: next window; 2: done
                        s_next_run = spos;
                    }
                }
            } else if (!part_ready) {
                load_part(kPartSlots / 2, kPartSlots / 2, threadIdx.x - 32, kInsThreads - 32);
            }
            part_ready = true;
            __syncthreads();
        };
        auto issue_copy = [&](uint32_t i) {   // span i of the table -> stage buffer i % kStageDepth
            const uint32_t r0 = s_span[i][1], r1 = s_span[i][2], a = s_span[i][3], n = s_span[i][4];
            for (uint32_t r = r0 + warp; r < r1; r += n_warps) {
                const uint32_t rp = run_pos[r - w0], len = run_len[r - w0];
                const uint32_t lo = max(rp, a), hi = min(rp + len, a + n);
                const unsigned long long *src = reinterpret_cast<const unsigned long long *>(run_src[r - w0]);
                // (runs are ~64 k-mers = two trips: the 4x-unrolled form the compiler builds, with its remainder
                //  ladders, costs more instructions than the copies)
                const uint32_t dst_s = stage_s + ((i % kStageDepth) * kStageCap - a) * 8u;
#pragma unroll 1
                for (uint32_t k = lo + lane; k < hi; k += 32)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_s + k * 8u), "l"(src + (k - rp)) : "memory");
            }
        };
        uint32_t from = spos;
        bool replan_same_window = false;
        for (;;) {
            if (!total_runs) {   // nothing for this partition: it is only loaded and written back
                load_part(kPartSlots / 2, kPartSlots / 2, threadIdx.x, kInsThreads);
                break;
            }
            if (replan_same_window) {
                // the span table was full: warp 0 continues from its cursor over the same window
                __syncthreads();
                if (warp == 0) {
                    uint32_t ns = 0;
                    int st = 0;
                    Span sp;
                    while (ns < kMaxSpans && (st = next_span(sp)) == 0) {
                        if (lane == 0) {
                            s_span[ns][0] = sp.c;
                            s_span[ns][1] = sp.r0;
                            s_span[ns][2] = sp.r1;
                            s_span[ns][3] = sp.a;
                            s_span[ns][4] = sp.n;
                        }
                        ns++;
                    }
                    if (lane == 0) {
                        s_nspans = ns;
                        s_more = ns == kMaxSpans ? 0 : st;
                        s_next_run = spos;
                    }
                }
                __syncthreads();
            } else {
                __syncthreads();   // the previous table / window is no longer read
                stage_window(from);
            }
            const uint32_t n_spans = s_nspans;
            // ---- execute ----
            for (uint32_t i = 0; i < kStageDepth - 1; i++) {
                if (i < n_spans) issue_copy(i);
                cp_async_commit();
            }
            if (n_spans) {
                cp_async_wait<kStageDepth - 2>();   // span 0 has landed (this thread's part)
                __syncthreads();
            }
            for (uint32_t i = 0; i < n_spans; i++) {
                if (i + kStageDepth - 1 < n_spans) issue_copy(i + kStageDepth - 1);   // into the buffer span i-1 used
                cp_async_commit();
                const uint32_t c = s_span[i][0], span_n = s_span[i][4];
                const unsigned long long n_new_before = n_new;
                // Fewer occupied slots + k-mers in this span than slots: an empty slot exists throughout the span, so
                // every probe sequence ends and the loop needs no probe counter.  (s_occ only grows while it is read.)
                const bool roomy = s_occ + span_n < kPartSlots;
                if (!s_fail && roomy) {
                    const uint32_t src_s = stage_s + (i % kStageDepth) * kStageCap * 8u;
                    const uint32_t ph_s = phist_s + c * L.k_low * 4u;   // this chunk's histogram moves
                    uint32_t new_here = 0;
                    for (uint32_t i0 = 0; i0 < span_n; i0 += kInsThreads) {
                        const uint32_t k = i0 + threadIdx.x;
                        const unsigned long long kmer = k < span_n ? lds_u64(src_s + k * 8u) : SKM_EMPTY_KEY;
                        uint32_t active = kmer != SKM_EMPTY_KEY;
                        uint32_t s = 0;
                        if (active) {
                            uint64_t h = skm_hash_kmer(kmer);
                            if (L.n_ranks != 1) h = skm_local_hash(h, L.n_ranks);
                            const uint64_t home = h >> home_shift;   // skm_home_slot, log2cap > 0
                            if (filter && (home >> kPartLog2) != q) active = 0;
                            s = (uint32_t)home & (kPartSlots - 1);
                        }
                        const uint32_t counted = active;
                        uint32_t off = s * 8u;   // byte offset of the probed slot's key
                        // warp-uniform probe loop (lanes that are done idle, predicated off)
                        for (;;) {
                            if (active) {
                                const uint32_t sa = keys_s + off;
                                unsigned long long key = lds_u64(sa);
                                if (key == SKM_EMPTY_KEY) {
                                    key = atoms_cas_u64(sa, (unsigned long long)SKM_EMPTY_KEY, kmer);
                                    if (key == SKM_EMPTY_KEY) {
                                        key = kmer;
                                        new_here++;
                                    }
                                }
                                if (key == kmer) active = 0;
                                else off = (off + 8u) & (kPartSlots * 8u - 8u);
                            }
                            if (!__any_sync(0xffffffffu, active)) break;
                        }
                        if (!counted) continue;
                        s = off >> 3;
                        if (!kHisto) {
                            reds_add_u32(counts_s + s * 4u, 1u);
                        } else {
                            const uint32_t old = atoms_add_u32(counts_s + s * 4u, 1u);
                            if (old + 1 < fast_lim) {   // (see the guarded loop below)
                                reds_add_u32(ph_s + old * 4u, 0xFFFFFFFFu);
                                reds_add_u32(ph_s + old * 4u + 4u, 1u);
                            } else {
                                const uint32_t ob = old > hm ? top32 : old;
                                const uint32_t nb2 = old >= hm ? top32 : old + 1;
                                if (ob != nb2) {
                                    if (old) {
                                        if (ob < L.k_low) reds_add_u32(ph_s + ob * 4u, 0xFFFFFFFFu);
                                        else hlog_push(hlog, &s_hlog_n, &s_dirty, L, c, ob, true);
                                    }
                                    if (nb2 < L.k_low) reds_add_u32(ph_s + nb2 * 4u, 1u);
                                    else hlog_push(hlog, &s_hlog_n, &s_dirty, L, c, nb2, false);
                                }
                            }
                        }
                    }
                    n_new += new_here;
                } else if (!s_fail) {
                    const unsigned long long *src = stage + (i % kStageDepth) * kStageCap;
                    int *ph = phist + c * L.k_low;  // this chunk's histogram moves
                    (void)ph;
                    for (uint32_t i0 = 0; i0 < span_n; i0 += kInsThreads) {
                        const uint32_t k = i0 + threadIdx.x;
                        const unsigned long long kmer = k < span_n ? src[k] : SKM_EMPTY_KEY;
                        uint32_t active = kmer != SKM_EMPTY_KEY;
                        uint32_t s = 0;
                        if (active) {
                            const uint64_t h = skm_hash_kmer(kmer);
                            const uint64_t home = skm_home_slot(L.n_ranks == 1 ? h : skm_local_hash(h, L.n_ranks), L.log2cap);
                            if (filter && (home >> kPartLog2) != q) active = 0;
                            s = (uint32_t)home & (kPartSlots - 1);
                        }
                        // The probe loop is warp-uniform (lanes that are done idle, predicated off): with
                        // per-lane `break`s the compiler kept the lanes apart for the rest of the iteration
                        // and the count update ran with 7 of 32 lanes on average.
                        uint32_t counted = active;
                        uint32_t probes = 0;
                        while (__any_sync(0xffffffffu, active)) {
                            if (active) {
                                unsigned long long key = keys[s];
                                if (key == SKM_EMPTY_KEY) {
                                    key = atomicCAS(&keys[s], (unsigned long long)SKM_EMPTY_KEY, kmer);
                                    if (key == SKM_EMPTY_KEY) {
                                        key = kmer;
                                        n_new++;
                                    }
                                }
                                if (key == kmer) {
                                    active = 0;
                                } else if (++probes >= kPartSlots) {  // the partition holds only other keys
                                    active = 0;
                                    counted = 0;
                                    s_fail = 1;
                                } else {
                                    s = (s + 1) & (kPartSlots - 1);
                                }
                            }
                        }
                        if (!counted) continue;
                        if (!kHisto) {
                            atomicAdd(&counts[s], 1u);
                        } else {
                            const uint32_t old = atomicAdd(&counts[s], 1u);
                            // Histogram::move_count (src/kmer/histogram.rs:51-85): one unit of mass from bin old to old+1
                            if (old + 1 < fast_lim) {
                                // both bins are shared-memory bins below the clamp.  Bin 0 is a scratch cell
                                // (the histogram has no bin 0: it takes the -1 of a new key and is never
                                // published), so the common case is two unconditional atomics.
                                atomicAdd(&ph[old], -1);
                                atomicAdd(&ph[old + 1], 1);
                            } else {
                                const uint32_t ob = old > hm ? top32 : old;
                                const uint32_t nb2 = old >= hm ? top32 : old + 1;
                                if (ob != nb2) {
                                    if (old) {
                                        if (ob < L.k_low) atomicAdd(&ph[ob], -1);
                                        else hlog_push(hlog, &s_hlog_n, &s_dirty, L, c, ob, true);
                                    }
                                    if (nb2 < L.k_low) atomicAdd(&ph[nb2], 1);
                                    else hlog_push(hlog, &s_hlog_n, &s_dirty, L, c, nb2, false);
                                }
                            }
                        }
                    }
                }
                {   // occupancy: one shared-memory atomic per warp and span
                    const uint32_t add = __reduce_add_sync(0xffffffffu, (uint32_t)(n_new - n_new_before));
                    if (lane == 0 && add) atomicAdd(&s_occ, add);
                }

                cp_async_wait<kStageDepth - 2>();   // span i+1 has landed (this thread's part)
                __syncthreads();   // span i is counted everywhere; span i+1 is visible to everybody
                if (s_occ > L.max_occupied) s_fail = 1;   // too full to go on: roll the partition back, grow, retry
            }
            const uint32_t more = s_more;
            if (more == 2) break;
            replan_same_window = more == 0;
            from = s_next_run;
        }
        __syncthreads();
        // ---- commit or roll back ----
        if (!s_fail) {
            unsigned long long nk = 0, nd = 0, ns = 0;
#pragma unroll 4
            for (uint32_t i = threadIdx.x; i < kPartSlots; i += kInsThreads) {
                const unsigned long long key = keys[i];
                unsigned long long cnt = counts[i];
                if (s_big && counts[i] >= kBigCount && !L.fresh && key != SKM_EMPTY_KEY) {
                    // the slot may have held a count >= 2^31 before the launch (kept as the marker kBigCount + adds)
                    const uint4 o = *(reinterpret_cast<const uint4 *>(tp) + i);
                    const unsigned long long old64 = ((unsigned long long)o.w << 32) | o.z;
                    const unsigned long long oldkey = ((unsigned long long)o.y << 32) | o.x;
                    if (oldkey == key && old64 >= kBigCount) cnt = old64 + (cnt - kBigCount);
                }
                uint4 v;
                v.x = (uint32_t)key;
                v.y = (uint32_t)(key >> 32);
                v.z = (uint32_t)cnt;
                v.w = (uint32_t)(cnt >> 32);
                reinterpret_cast<uint4 *>(tp)[i] = v;
                if (key != SKM_EMPTY_KEY) {
                    if (cnt >= kU32Max) {
                        cnt = kU32Max;
                        ns++;
                    }
                    nd++;
                    nk += cnt;
                    const unsigned long long bin = cnt > L.histo_max ? top : cnt;
                    if (bin < L.k_low) atomicAdd(&fhist[(uint32_t)bin], 1);
                    else atomicAdd(&L.g_recount[bin], 1ull);
                }
            }
            n_kmers_cta += nk;
            n_dist_cta += nd;
            n_sat_cta += ns;
            n_new_cta += n_new;
            if (kHisto) {
                for (uint32_t i = threadIdx.x; i < L.n_chunks * L.k_low; i += kInsThreads) {
                    const int d = phist[i];
                    if (d) {
                        if (i % L.k_low) chist[i] += d;   // (bin 0 is the scratch cell)
                        phist[i] = 0;
                    }
                }
                const uint32_t nh = min(s_hlog_n, kHlogCap);
                for (uint32_t i = threadIdx.x; i < nh; i += kInsThreads) {
                    const uint32_t e = hlog[i];
                    atomicAdd(&L.g_delta[(size_t)(e >> 24) * (L.histo_max + 2) + (e & 0x7FFFFFu)],
                              (e & 0x800000u) ? ~0ull : 1ull);
                }
            }
        } else {
            // rolled back: nothing of this partition is published.  (A launch on a logically empty
            // table owns the partition's memory: leave it physically empty for the retry.)
            if (L.fresh)
                for (uint32_t i = threadIdx.x; i < kPartSlots; i += kInsThreads) {
                    uint4 v;
                    v.x = v.y = 0xFFFFFFFFu;
                    v.z = v.w = 0u;
                    reinterpret_cast<uint4 *>(tp)[i] = v;
                }
            if (kHisto)
                for (uint32_t i = threadIdx.x; i < L.n_chunks * L.k_low; i += kInsThreads) phist[i] = 0;
            if (threadIdx.x == 0) {
                const unsigned long long e = atomicAdd(&L.gc->n_failed, 1ull);
                if (e < L.fail_cap) L.fail_list[e] = (uint32_t)q;
                if (s_dirty) L.gc->fatal = 1ull;
            }
        }
    }
    __syncthreads();
    // ---- publish this CTA's accumulators ----
    for (uint32_t i = threadIdx.x; i < L.k_low; i += kInsThreads)
        if (fhist[i]) atomicAdd(&L.g_recount[i > L.histo_max ? top : i], (unsigned long long)(long long)fhist[i]);
    if (kHisto)
        for (uint32_t i = threadIdx.x; i < L.n_chunks * L.k_low; i += kInsThreads)
            if (chist[i]) {
                const unsigned long long bin = i % L.k_low;
                atomicAdd(&L.g_delta[(size_t)(i / L.k_low) * (L.histo_max + 2) + (bin > L.histo_max ? top : bin)],
                          (unsigned long long)(long long)chist[i]);
            }
    block_add(&L.gc->n_distinct, n_new_cta);
    block_add(&L.tot->n_kmers, n_kmers_cta);
    block_add(&L.tot->n_distinct, n_dist_cta);
    block_add(&L.tot->n_saturated, n_sat_cta);
}

// Columns of the incremental histogram from the per-chunk moves: column c = base + sum of the
// moves of chunks <= c (src/io.rs:1023-1028: chunks are merged in index order and the histogram
// is read after each).  `base` (the running histogram) ends as the last column.
__global__ void __launch_bounds__(256)
hist_columns_kernel(const unsigned long long *delta, uint32_t n_chunks, unsigned long long n_bins,
                    unsigned long long *__restrict__ base, unsigned long long *cols /* may be null or == delta */) {
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < n_bins;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long run = base[b];
        for (uint32_t c = 0; c < n_chunks; c++) {
            run += delta[(size_t)c * n_bins + b];
            if (cols) cols[(size_t)c * n_bins + b] = run;
        }
        base[b] = run;
    }
}

// *p += delta (delta may be "negative" in two's complement)
__global__ void adjust_counter_kernel(unsigned long long *p, unsigned long long delta) { *p += delta; }

// ---------------------------------------------------------------------------
// synthetic reads on the device (bench / tests)
// ---------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
synth_kernel(skm_synth_params p, uint32_t chunk_index, uint32_t n_chunks, uint64_t first, uint64_t n,
             uint8_t *__restrict__ out) {
    const uint64_t line = (uint64_t)p.read_len + 1;
    const uint64_t total = n * line;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r_local = first + i / line;
        const uint32_t j = (uint32_t)(i % line);
        const uint64_t r = skm_chunk_read_to_global(r_local, chunk_index, n_chunks);
        out[i] = j == p.read_len ? (uint8_t)'\n' : skm_synth_read_base(&p, r, j);
    }
}

// ---------------------------------------------------------------------------
// random-access roofline probe
// ---------------------------------------------------------------------------

// variant 0: key load + RED.ADD.64 (what an insert hit does); 1: RED only; 2: load only.
// variants 3..5: the same three, but update i goes to a random slot of region
// i / per_region (regions of 2^region_log2 slots visited in order) — the access
// pattern of the partitioned insert, whose working set stays in L2.
// variant 6: region-local key load + 32-bit RED.
__global__ void __launch_bounds__(256)
gups_kernel(Slot *__restrict__ table, uint32_t log2cap, uint64_t n, uint64_t salt, int variant,
            uint32_t region_log2, unsigned long long *__restrict__ sink) {
    unsigned long long acc = 0;
    const bool local = variant >= 3;
    const int op = variant == 6 ? 0 : (local ? variant - 3 : variant);
    const uint64_t n_regions = 1ull << (log2cap - region_log2);
    const uint64_t per_region = (n + n_regions - 1) / n_regions;
    if (!local) {
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
             i += (uint64_t)gridDim.x * blockDim.x) {
            const uint64_t s = skm_home_slot(skm_mix64(i + salt), log2cap);
            if (op != 1) acc += ld_cg_u64(&table[s].key);
            if (op != 2) red_add_u64(&table[s].count, 1ull);
        }
    } else {
        // one update per thread, in list order (like insert_sorted_list_kernel)
        const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n) {
            const uint64_t region = i / per_region;
            const uint64_t s = (region << region_log2) | (skm_mix64(i + salt) >> (64 - region_log2));
            if (op != 1) acc += ld_cg_u64(&table[s].key);
            if (op != 2) {
                if (variant == 6)
                    atomicAdd(reinterpret_cast<unsigned int *>(&table[s].count), 1u);
                else
                    red_add_u64(&table[s].count, 1ull);
            }
        }
    }
    if (acc == 0x123456789ull) *sink = acc;
}

}  // namespace skm
