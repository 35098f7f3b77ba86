"""Python statement of include/skm_common.h (hash, revcomp, digest) for host-side
helpers and tests.  Pure integer arithmetic; must stay bit-identical to the header."""
M64 = (1 << 64) - 1
EMPTY_KEY = M64


def mix64(x: int) -> int:
    x &= M64
    x ^= x >> 33
    x = (x * 0xFF51AFD7ED558CCD) & M64
    x ^= x >> 33
    x = (x * 0xC4CEB9FE1A85EC53) & M64
    x ^= x >> 33
    return x


def hash_kmer(kmer: int) -> int:
    return mix64(kmer)


def owner_rank(h: int, n_ranks: int) -> int:
    """floor(h * n_ranks / 2^64): contiguous hash ranges per rank."""
    return (h * n_ranks) >> 64


def local_hash(h: int, n_ranks: int) -> int:
    return (h * n_ranks) & M64


def home_slot(local_h: int, log2_capacity: int) -> int:
    return (local_h >> (64 - log2_capacity)) if log2_capacity else 0


def route_bucket(kmer: int, n_ranks: int, log2_regions: int) -> int:
    """Bucket of the routing / partition pass: owner rank, then the top bits of the local hash."""
    h = hash_kmer(kmer)
    o, lh = owner_rank(h, n_ranks), local_hash(h, n_ranks)
    return (o << log2_regions) | (lh >> (64 - log2_regions) if log2_regions else 0)


def pair_digest(kmer: int, count: int) -> int:
    return mix64(kmer ^ mix64((0x9E3779B97F4A7C15 + count) & M64))


def revcomp_kmer(kmer: int, k: int) -> int:
    x = ~kmer & M64
    for sh, m in ((2, 0x3333333333333333), (4, 0x0F0F0F0F0F0F0F0F), (8, 0x00FF00FF00FF00FF),
                  (16, 0x0000FFFF0000FFFF)):
        x = ((x >> sh) & m) | ((x & m) << sh)
    x = ((x >> 32) | (x << 32)) & M64
    return x >> (64 - 2 * k)


def rate_to_thresh(rate: float) -> int:
    """A probability as the 32-bit threshold skm_synth_params takes (include/skm_common.h)."""
    return min(0xFFFFFFFF, int(round(rate * 4294967296.0)))
