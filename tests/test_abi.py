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
