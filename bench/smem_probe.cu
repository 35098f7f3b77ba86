// smem_probe.cu — round-2 experiment: how fast can a CTA count k-mers into a table partition that
// lives in SHARED memory (table streamed through the SMs once, no L2/DRAM atomics)?
//
// Models the C2 workload per partition: S slots, n_chunks runs of `run` k-mers each, 81 % of the
// occurrences drawn from a pool of "genomic" keys (many repeats), 19 % unique "error" keys.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o bench/smem_probe bench/smem_probe.cu
// Run  : bench/smem_probe            (prints one line per variant: ms, G k-mers/s)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

static constexpr unsigned long long EMPTY = ~0ull;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x;
}

struct __align__(16) Slot { unsigned long long key, count; };

// lists[(p * n_chunks + c) * run + i]
__global__ void gen_kernel(unsigned long long *lists, uint64_t n_part, uint32_t n_chunks, uint32_t run, uint32_t pool,
                           uint32_t genomic_permille) {
    const uint64_t total = n_part * n_chunks * run;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t p = i / ((uint64_t)n_chunks * run);
        const uint64_t r = mix64(i * 0x9e3779b97f4a7c15ull + 12345);
        unsigned long long key;
        if ((r % 1000) < genomic_permille) key = mix64(p * 1000003ull + ((r >> 20) % pool)) >> 2;   // repeated
        else key = (mix64(i ^ 0xabcdef1234567ull) >> 2) | (1ull << 61);                            // unique
        lists[i] = key;
    }
}

__global__ void clear_kernel(Slot *t, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 v; v.x = v.y = 0xFFFFFFFFu; v.z = v.w = 0; reinterpret_cast<uint4 *>(t)[i] = v;
    }
}

enum { HIST_NONE = 0, HIST_ATOMS = 1, HIST_PRIVATE = 2 };
static constexpr int kLow = 64;

template <int SLOG2, int T, int HIST>
__global__ void __launch_bounds__(T)
tile_insert_kernel(Slot *__restrict__ table, const unsigned long long *__restrict__ lists, uint64_t n_part,
                   uint32_t n_chunks, uint32_t run, unsigned long long *__restrict__ part_counter,
                   unsigned long long *__restrict__ g_hist /* n_chunks * kLow */, unsigned long long *__restrict__ g_new,
                   int load_table) {
    constexpr uint32_t S = 1u << SLOG2;
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem);
    uint32_t *counts = reinterpret_cast<uint32_t *>(keys + S);
    int *chist = reinterpret_cast<int *>(counts + S);                 // n_chunks * kLow
    short *priv = reinterpret_cast<short *>(chist + 16 * kLow);       // HIST_PRIVATE: kLow * T int16
    __shared__ unsigned long long s_p;
    __shared__ int s_maxbin;
    for (uint32_t i = threadIdx.x; i < 16 * kLow; i += T) chist[i] = 0;
    if (HIST == HIST_PRIVATE)
        for (uint32_t i = threadIdx.x; i < kLow * T / 2; i += T) reinterpret_cast<int *>(priv)[i] = 0;
    unsigned long long n_new = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_p = atomicAdd(part_counter, 1ull);
        __syncthreads();
        const uint64_t p = s_p;
        if (p >= n_part) break;
        Slot *tp = table + p * S;
        if (load_table) {
            for (uint32_t i = threadIdx.x; i < S; i += T) {
                const uint4 v = reinterpret_cast<const uint4 *>(tp)[i];
                keys[i] = ((unsigned long long)v.y << 32) | v.x;
                counts[i] = v.z;
            }
        } else {
            for (uint32_t i = threadIdx.x; i < S; i += T) { keys[i] = EMPTY; counts[i] = 0; }
        }
        __syncthreads();
        for (uint32_t c = 0; c < n_chunks; c++) {
            const unsigned long long *lp = lists + (p * n_chunks + c) * (uint64_t)run;
            int my_max = 0;
            for (uint32_t i = threadIdx.x; i < run; i += T) {
                const unsigned long long kmer = lp[i];
                uint32_t s = (uint32_t)(mix64(kmer) >> (64 - SLOG2));
                for (;;) {
                    unsigned long long k = keys[s];
                    if (k == kmer) break;
                    if (k == EMPTY) {
                        k = atomicCAS(&keys[s], EMPTY, kmer);
                        if (k == EMPTY) { n_new++; break; }
                        if (k == kmer) break;
                    }
                    s = (s + 1) & (S - 1);
                }
                const uint32_t old = atomicAdd(&counts[s], 1u);
                if (HIST == HIST_ATOMS) {
                    if (old < kLow - 1) {
                        if (old) atomicAdd(&chist[c * kLow + old], -1);
                        atomicAdd(&chist[c * kLow + old + 1], 1);
                    }
                } else if (HIST == HIST_PRIVATE) {
                    if (old < kLow - 1) {
                        if (old) priv[old * T + threadIdx.x] -= 1;
                        priv[(old + 1) * T + threadIdx.x] += 1;
                        my_max = max(my_max, (int)old + 1);
                    }
                }
            }
            if (HIST == HIST_PRIVATE) {
                // fold the thread-private deltas of this chunk into the CTA's per-chunk histogram
                if (threadIdx.x == 0) s_maxbin = 0;
                __syncthreads();
                my_max = __reduce_max_sync(0xffffffffu, my_max);
                if ((threadIdx.x & 31) == 0 && my_max) atomicMax(&s_maxbin, my_max);
                __syncthreads();
                const int nb = s_maxbin + 1;
                const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = T / 32;
                for (int b = w; b < nb; b += nw) {
                    int sum = 0;
                    int *row = reinterpret_cast<int *>(priv + b * T);   // T/2 packed pairs
                    for (uint32_t j = lane; j < T / 2; j += 32) {
                        const int v = row[j];
                        sum += (int)(short)(v & 0xffff) + (v >> 16);
                        row[j] = 0;
                    }
                    sum = __reduce_add_sync(0xffffffffu, sum);
                    if (lane == 0 && sum) chist[c * kLow + b] += sum;
                }
            }
            __syncthreads();
        }
        for (uint32_t i = threadIdx.x; i < S; i += T) {
            const unsigned long long k = keys[i];
            uint4 v;
            v.x = (uint32_t)k; v.y = (uint32_t)(k >> 32); v.z = counts[i]; v.w = 0;
            reinterpret_cast<uint4 *>(tp)[i] = v;
        }
    }
    __syncthreads();
    if (HIST != HIST_NONE)
        for (uint32_t i = threadIdx.x; i < n_chunks * kLow; i += T)
            if (chist[i]) atomicAdd(&g_hist[i], (unsigned long long)(long long)chist[i]);
    n_new = __reduce_add_sync(0xffffffffu, (unsigned)n_new);
    if ((threadIdx.x & 31) == 0 && n_new) atomicAdd(g_new, n_new);
}

// raw shared-memory atomic throughput: each thread does `iters` atomics on pseudo-random words of a 64 KB array
template <int OP>
__global__ void __launch_bounds__(256) atoms_kernel(uint32_t iters, unsigned long long *sink) {
    __shared__ __align__(16) unsigned long long a[4096];
    for (uint32_t i = threadIdx.x; i < 4096; i += 256) a[i] = OP == 2 ? EMPTY : 0;
    __syncthreads();
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 1;
    unsigned long long acc = 0;
    for (uint32_t i = 0; i < iters; i++) {
        x = x * 1664525u + 1013904223u;
        const uint32_t s = (x >> 12) & 4095;
        if (OP == 0) acc += atomicAdd(reinterpret_cast<uint32_t *>(a) + s, 1u);          // ATOMS.ADD 32, result used
        else if (OP == 1) atomicAdd(reinterpret_cast<uint32_t *>(a) + s, 1u);            // no result (RED-like)
        else if (OP == 2) acc += atomicCAS(&a[s], EMPTY, (unsigned long long)x);         // CAS 64
        else if (OP == 3) acc += a[s];                                                   // LDS.64 random
        else if (OP == 4) acc += atomicAdd(&a[s], 1ull);                                 // ADD 64 with result
        else if (OP == 5) { uint32_t *q = reinterpret_cast<uint32_t *>(a) + ((s & ~255u) | threadIdx.x); *q += 1; }  // private RMW
    }
    if (acc == 0x1234567) *sink = acc;
}

template <int SLOG2, int T, int HIST>
void run_variant(const char *name, Slot *table, const unsigned long long *lists, uint64_t total_slots_log2, uint32_t n_chunks,
                 uint64_t kmers_per_slot_x1000, unsigned long long *d_ctr, unsigned long long *d_hist, unsigned long long *d_new,
                 int sm_count, int load_table) {
    constexpr uint32_t S = 1u << SLOG2;
    const uint64_t n_part = 1ull << (total_slots_log2 - SLOG2);
    const uint32_t run = (uint32_t)(S * kmers_per_slot_x1000 / 1000 / n_chunks);
    size_t smem = (size_t)S * 12 + 16 * kLow * 4 + (HIST == HIST_PRIVATE ? (size_t)kLow * T * 2 : 0);
    auto kern = tile_insert_kernel<SLOG2, T, HIST>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
    if (occ < 1) { printf("%-28s does not fit (smem %zu)\n", name, smem); return; }
    const int grid = sm_count * occ;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    unsigned long long h_new = 0;
    for (int it = 0; it < 3; it++) {
        CK(cudaMemset(d_ctr, 0, 8));
        CK(cudaMemset(d_hist, 0, 16 * kLow * 8));
        CK(cudaMemset(d_new, 0, 8));
        if (load_table) { clear_kernel<<<sm_count * 8, 256>>>(table, 1ull << total_slots_log2); CK(cudaDeviceSynchronize()); }
        CK(cudaEventRecord(a));
        kern<<<grid, T, smem>>>(table, lists, n_part, n_chunks, run, d_ctr, d_hist, d_new, load_table);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaGetLastError());
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
        CK(cudaMemcpy(&h_new, d_new, 8, cudaMemcpyDeviceToHost));
    }
    const double n = (double)n_part * n_chunks * run;
    unsigned long long hist[16 * kLow];
    CK(cudaMemcpy(hist, d_hist, sizeof hist, cudaMemcpyDeviceToHost));
    long long support = 0;
    for (uint32_t c = 0; c < n_chunks; c++) for (int b2 = 1; b2 < kLow; b2++) support += (long long)hist[c * kLow + b2];
    printf("%-28s S=2^%d T=%d occ=%d run=%u load_table=%d : %7.3f ms  %6.2f G kmers/s  distinct=%llu (load %.3f) hist_support=%lld\n",
           name, SLOG2, T, occ, run, load_table, best, n / best / 1e6, h_new, (double)h_new / (double)(1ull << total_slots_log2), support);
}

int main(int argc, char **argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    const int sm = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, sm);
    const uint32_t total_log2 = argc > 1 ? atoi(argv[1]) : 29;   // table slots
    const uint32_t n_chunks = 10;
    const uint64_t per_slot_x1000 = 2371;                         // C2: 1.273e9 k-mers / 2^29 slots
    unsigned long long *d_ctr, *d_hist, *d_new, *d_sink;
    CK(cudaMalloc(&d_ctr, 8)); CK(cudaMalloc(&d_hist, 16 * kLow * 8)); CK(cudaMalloc(&d_new, 8)); CK(cudaMalloc(&d_sink, 8));

    // raw ATOMS throughput
    {
        const char *names[] = {"ATOMS.ADD.32 (result)", "ATOMS.ADD.32 (no result)", "ATOMS.CAS.64", "LDS.64 random", "ATOMS.ADD.64 (result)", "private LDS+STS RMW"};
        for (int op = 0; op < 6; op++) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            const uint32_t iters = 4096;
            const int grid = sm * 8;
            float best = 1e30f;
            for (int it = 0; it < 3; it++) {
                cudaEventRecord(a);
                switch (op) {
                case 0: atoms_kernel<0><<<grid, 256>>>(iters, d_sink); break;
                case 1: atoms_kernel<1><<<grid, 256>>>(iters, d_sink); break;
                case 2: atoms_kernel<2><<<grid, 256>>>(iters, d_sink); break;
                case 3: atoms_kernel<3><<<grid, 256>>>(iters, d_sink); break;
                case 4: atoms_kernel<4><<<grid, 256>>>(iters, d_sink); break;
                default: atoms_kernel<5><<<grid, 256>>>(iters, d_sink); break;
                }
                cudaEventRecord(b); CK(cudaEventSynchronize(b));
                float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
            }
            const double ops = (double)grid * 256 * iters;
            printf("%-28s : %7.3f ms  %7.1f G ops/s chip  (%.2f ops/clk/SM at 1.9 GHz)\n", names[op], best, ops / best / 1e6,
                   ops / best / 1e6 / sm / 1.9);
        }
    }

    Slot *table; unsigned long long *lists;
    const uint64_t n_kmers = ((1ull << total_log2) * per_slot_x1000) / 1000;
    CK(cudaMalloc(&table, sizeof(Slot) << total_log2));
    CK(cudaMalloc(&lists, (n_kmers + (1 << 20)) * 8));
    // one generation per partition size (pool of repeated keys scales with the partition)
    auto gen = [&](int slog2) {
        const uint64_t S = 1ull << slog2;
        const uint64_t n_part = 1ull << (total_log2 - slog2);
        const uint32_t run = (uint32_t)(S * per_slot_x1000 / 1000 / n_chunks);
        const uint32_t pool = (uint32_t)(S * 93 / 1000);   // 5e7 genomic k-mers / 2^29 slots
        gen_kernel<<<sm * 16, 256>>>(lists, n_part, n_chunks, run, pool, 810);
        CK(cudaDeviceSynchronize());
    };
    for (int load = 0; load <= 1; load++) {
        gen(12);
        run_variant<12, 256, HIST_NONE>("s12 none", table, lists, total_log2, n_chunks, per_slot_x1000, d_ctr, d_hist, d_new, sm, load);
        run_variant<12, 256, HIST_ATOMS>("s12 atoms", table, lists, total_log2, n_chunks, per_slot_x1000, d_ctr, d_hist, d_new, sm, load);
        run_variant<12, 256, HIST_PRIVATE>("s12 private", table, lists, total_log2, n_chunks, per_slot_x1000, d_ctr, d_hist, d_new, sm, load);
        run_variant<12, 128, HIST_PRIVATE>("s12 private T128", table, lists, total_log2, n_chunks, per_slot_x1000, d_ctr, d_hist, d_new, sm, load);
        gen(13);
        run_variant<13, 512, HIST_NONE>("s13 none", table, lists, total_log2, n_chunks, per_slot_x1000, d_ctr, d_hist, d_new, sm, load);
        run_variant<13, 512, HIST_ATOMS>("s13 atoms", table, lists, total_log2, n_chunks, per_slot_x1000, d_ctr, d_hist, d_new, sm, load);
        run_variant<13, 512, HIST_PRIVATE>("s13 private", table, lists, total_log2, n_chunks, per_slot_x1000, d_ctr, d_hist, d_new, sm, load);
        run_variant<13, 256, HIST_PRIVATE>("s13 private T256", table, lists, total_log2, n_chunks, per_slot_x1000, d_ctr, d_hist, d_new, sm, load);
        gen(14);
        run_variant<14, 1024, HIST_NONE>("s14 none", table, lists, total_log2, n_chunks, per_slot_x1000, d_ctr, d_hist, d_new, sm, load);
        run_variant<14, 1024, HIST_ATOMS>("s14 atoms", table, lists, total_log2, n_chunks, per_slot_x1000, d_ctr, d_hist, d_new, sm, load);
        run_variant<14, 512, HIST_PRIVATE>("s14 private T512", table, lists, total_log2, n_chunks, per_slot_x1000, d_ctr, d_hist, d_new, sm, load);
    }
    return 0;
}
