"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a,
loads, and exports every symbol include/sharkmer_b200.h declares.  No compute
calls (there is no GPU here); the engine must fail loudly without a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from sharkmer_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared():
    src = open(os.path.join(ROOT, "include", "sharkmer_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(skm_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from sharkmer_b200 import _lib
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes signature"
    assert sorted(_lib.SYMBOLS) == names


def test_struct_sizes_match_header(lib):
    from sharkmer_b200 import _lib
    assert C.sizeof(_lib.SkmParams) == 56
    assert C.sizeof(_lib.SkmTotals) == 56
    assert lib.skm_abi_version() == 1


def test_sm100a_cubin_present():
    import subprocess
    so = os.path.join(ROOT, "sharkmer_b200", "libsharkmer_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(lib):
    """Without a CUDA device the engine refuses to exist (and says why)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sharkmer_b200.kmer import Engine, SkmError
    with pytest.raises(SkmError) as e:
        Engine(21)
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)


def test_param_validation_messages(lib):
    """k / histo_max bounds (src/cli.rs:662-673) are checked before the device is touched."""
    from sharkmer_b200.kmer import Engine, SkmError
    for k in (0, 32, 33):
        with pytest.raises(SkmError) as e:
            Engine(k)
        assert e.value.code == 1 and "k must be less than 32" in str(e.value)
    with pytest.raises(SkmError) as e:
        Engine(20)
    assert "k must be odd" in str(e.value)
    for hm in (0, 1000001):
        with pytest.raises(SkmError) as e:
            Engine(21, histo_max=hm)
        assert e.value.code == 1


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under sharkmer_b200/ or include/ may reference it."""
    bad = []
    for base in ("sharkmer_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".c")):
                    txt = open(os.path.join(dp, fn), errors="replace").read()
                    if re.search(r"liboracle|skm_oracle|from oracle|import oracle|orc_", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_python_mirror_of_skm_common_matches_the_header(tmp_path):
    """sharkmer_b200/common.py must stay bit-identical to include/skm_common.h (hash, owner rank,
    local hash, home slot, digest, revcomp, routing bucket)."""
    import random
    import subprocess
    from sharkmer_b200 import common
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdlib.h>
#include "skm_common.h"
int main(int argc, char **argv) {
    unsigned long long x; unsigned n, k, l2;
    while (scanf("%llu %u %u %u", &x, &n, &k, &l2) == 4) {
        uint64_t h = skm_hash_kmer(x);
        uint64_t lh = skm_local_hash(h, n);
        uint32_t owner = skm_owner_rank(h, n);
        uint32_t bucket = (owner << l2) | (l2 ? (uint32_t)(lh >> (64 - l2)) : 0);
        printf("%llu %u %llu %llu %llu %llu %u\n", (unsigned long long)h, owner, (unsigned long long)lh,
               (unsigned long long)skm_home_slot(lh, 29), (unsigned long long)skm_pair_digest(x, k),
               (unsigned long long)skm_revcomp_kmer(x & ((1ull << (2 * k)) - 1), k), bucket);
    }
    return 0;
}
''')
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-O1", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    rng = random.Random(3)
    cases = [(rng.getrandbits(62), rng.choice([1, 2, 3, 7, 8, 16]), rng.randint(1, 31), rng.choice([0, 3, 7, 10]))
             for _ in range(500)]
    out = subprocess.run([str(exe)], input="".join(f"{x} {n} {k} {l2}\n" for x, n, k, l2 in cases),
                         capture_output=True, text=True, check=True).stdout.split("\n")
    for (x, n, k, l2), line in zip(cases, out):
        h, owner, lh, home, dig, rc, bucket = map(int, line.split())
        assert common.hash_kmer(x) == h
        assert common.owner_rank(h, n) == owner and owner < n
        assert common.local_hash(h, n) == lh
        assert common.home_slot(lh, 29) == home
        assert common.pair_digest(x, k) == dig
        assert common.revcomp_kmer(x & ((1 << (2 * k)) - 1), k) == rc
        assert common.route_bucket(x, n, l2) == bucket
