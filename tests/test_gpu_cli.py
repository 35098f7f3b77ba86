"""End-to-end on the GPU: FASTQ file(s) -> {sample}.histo / .final.histo / .stats.yaml
through the C++ host driver (sharkmer_b200/host) + CUDA engine, compared byte for
byte with the files the oracle CLI writes from the same input (formats of
src/io.rs:1049-1094 and src/stats.rs:26-45)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_CLI = os.path.join(ROOT, "sharkmer_b200", "host", "sharkmer_b200_cli")
ORC_CLI = os.path.join(ROOT, "oracle", "sharkmer_oracle")


def stats_fields(path):
    d = {}
    for line in open(path):
        k, _, v = line.partition(": ")
        d[k] = v.strip()
    d.pop("command", None)
    d.pop("peak_memory_bytes", None)
    return d


def run_both(tmp_path, args, inputs, gpu_flags=()):
    outs = []
    for name, cli, extra in (("gpu", GPU_CLI, list(gpu_flags)), ("orc", ORC_CLI, [])):
        d = tmp_path / name
        d.mkdir(exist_ok=True)
        r = subprocess.run([cli, *map(str, args), *extra, "-s", "smp", "-o", str(d) + "/", *map(str, inputs)],
                           capture_output=True, text=True)
        outs.append((r, d))
    return outs


# the multi-threaded reader (default) and the reference-shaped serial reader feed the same batches
@pytest.mark.parametrize("reader", [(), ("--serial",), ("-t", "3")])
@pytest.mark.parametrize("k,chunks,gz,n", [(21, 10, False, 25_500), (31, 1, True, 12_000), (25, 3, True, 7_777)])
def test_histo_files_byte_identical(oracle, tmp_path, k, chunks, gz, n, reader):
    assert os.path.exists(GPU_CLI) and os.path.exists(ORC_CLI)
    fq = tmp_path / ("reads.fastq.gz" if gz else "reads.fastq")
    oracle.synth_fastq(fq, seed=k, genome_len=80_000, read_len=150, sub_rate=0.01, n_rate=0.001, first=0, n=n, gzip=gz)
    (rg, dg), (ro, do) = run_both(tmp_path, ["-k", k, "--chunks", chunks, "--histo-max", 500], [fq], reader)
    assert rg.returncode == 0, rg.stderr
    assert ro.returncode == 0, ro.stderr
    for f in ("smp.histo", "smp.final.histo"):
        a, b = open(dg / f, "rb").read(), open(do / f, "rb").read()
        assert a == b, f
        assert a.startswith(f"# sharkmer 3.1.0 k={k} chunks={chunks}\n".encode())
    assert stats_fields(dg / "smp.stats.yaml") == stats_fields(do / "smp.stats.yaml")


def test_no_histogram_mode_and_max_reads(oracle, tmp_path):
    fq = tmp_path / "reads.fastq"
    oracle.synth_fastq(fq, seed=3, genome_len=50_000, read_len=100, sub_rate=0.01, n_rate=0.0, first=0, n=5_000)
    (rg, dg), (ro, do) = run_both(tmp_path, ["-k", 21, "-m", 3_333], [fq])
    assert rg.returncode == 0 and ro.returncode == 0, rg.stderr + ro.stderr
    assert not os.path.exists(dg / "smp.histo") and not os.path.exists(do / "smp.histo")  # chunks == 0
    sg = stats_fields(dg / "smp.stats.yaml")
    assert sg == stats_fields(do / "smp.stats.yaml")
    assert sg["n_reads_read"] == "3333" and "n_singleton_kmers" not in sg


def test_paired_files(oracle, tmp_path):
    f1, f2 = tmp_path / "R1.fastq.gz", tmp_path / "R2.fastq.gz"
    oracle.synth_fastq(f1, seed=5, genome_len=60_000, read_len=120, sub_rate=0.01, n_rate=0.001, first=0, n=4_100, gzip=True)
    oracle.synth_fastq(f2, seed=5, genome_len=60_000, read_len=120, sub_rate=0.01, n_rate=0.001, first=10_000, n=4_100, gzip=True)
    (rg, dg), (ro, do) = run_both(tmp_path, ["-k", 21, "--chunks", 4, "--histo-max", 200, "--paired", "-m", 6001], [f1, f2])
    assert rg.returncode == 0 and ro.returncode == 0, rg.stderr + ro.stderr
    for f in ("smp.histo", "smp.final.histo"):
        assert open(dg / f, "rb").read() == open(do / f, "rb").read()
    assert stats_fields(dg / "smp.stats.yaml") == stats_fields(do / "smp.stats.yaml")


def test_cli_errors_match_reference_text(oracle, tmp_path):
    bad = tmp_path / "bad.fastq"
    bad.write_text("@r0\nACGTACGTACGTACGTACGTACGTAC\n+\nIIIIIIIIIIIIIIIIIIIIIIIIII\n@r1\nACGTXCGT\n+\nIIIIIIII\n")
    (rg, _), (ro, _) = run_both(tmp_path, ["-k", 21], [bad])
    assert rg.returncode == 1 and ro.returncode == 1
    msg = "Invalid character 'X' in sequence. Only ACGTN allowed."
    assert msg in rg.stderr and msg in ro.stderr
    r = subprocess.run([GPU_CLI, "-k", "22", str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and "k must be odd" in r.stderr
    empty = tmp_path / "empty.fastq"
    empty.write_text("")
    (rg, _), (ro, _) = run_both(tmp_path, ["-k", 21], [empty])
    assert rg.returncode == 1 and "No reads were ingested" in rg.stderr and "No reads were ingested" in ro.stderr


@pytest.mark.parametrize("gpus,devices", [(2, "0,0"), (3, "0,0,0")])
def test_sharded_cli_writes_identical_files(oracle, tmp_path, gpus, devices):
    """`--gpus N` (skm_group_*: the multi-GPU path behind the C ABI, driven by the C++ host the way
    src/main.rs:112-131 drives one table): the same .histo / .final.histo / stats bytes as the oracle,
    with the N ranks sharing cuda:0."""
    fq = tmp_path / "reads.fastq.gz"
    oracle.synth_fastq(fq, seed=11, genome_len=80_000, read_len=150, sub_rate=0.01, n_rate=0.001, first=0, n=21_300, gzip=True)
    (rg, dg), (ro, do) = run_both(tmp_path, ["-k", 25, "--chunks", 5, "--histo-max", 300], [fq],
                                  ("--gpus", str(gpus), "--devices", devices, "--arena-mb", "128"))
    assert rg.returncode == 0, rg.stderr
    assert ro.returncode == 0, ro.stderr
    for f in ("smp.histo", "smp.final.histo"):
        assert open(dg / f, "rb").read() == open(do / f, "rb").read(), f
    assert stats_fields(dg / "smp.stats.yaml") == stats_fields(do / "smp.stats.yaml")
