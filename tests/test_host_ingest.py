"""Host-side logic of the hot path, on CPU: FASTQ framing, 1000-read batching and
round-robin chunk routing of sharkmer_b200/host/ingest.hpp (mirror of src/io.rs
:271-361, 541-543, 598-697), driven against a TEST-ONLY recording mock of the C ABI
(tests/host/mock_abi.cpp) and compared with (a) a direct Python statement of the
routing rule and (b) the oracle's per-chunk totals."""
import gzip
import os
import random
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def harness_bin(tmp_path_factory):
    out = tmp_path_factory.mktemp("host") / "host_harness"
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-o", str(out), os.path.join(HERE, "host", "mock_abi.cpp"),
                    "-lz"], check=True)
    return str(out)


# every test runs against the reference-shaped serial reader (ingest.hpp) and the multi-threaded one
# (fastq_parallel.hpp); "tiny" forces many windows, carried partial records and window growth
READERS = {"serial": ["--serial"], "parallel": ["--threads", "4"],
           "tiny": ["--threads", "3", "--window-bytes", "700"]}


@pytest.fixture(scope="module", params=list(READERS))
def harness(request, harness_bin):
    return [harness_bin] + READERS[request.param]


def write_fastq(path, seqs, gz=False, crlf=False, trailing_newline=True):
    nl = "\r\n" if crlf else "\n"
    txt = "".join(f"@r{i}{nl}{s}{nl}+{nl}{'I' * len(s)}{nl}" for i, s in enumerate(seqs))
    if not trailing_newline:
        txt = txt[:-len(nl)]
    data = txt.encode()
    if gz:
        with gzip.open(path, "wb") as f:
            f.write(data)
    else:
        with open(path, "wb") as f:
            f.write(data)


def rand_seqs(n, seed, lo=0, hi=160):
    rng = random.Random(seed)
    return ["".join(rng.choice("ACGTN" if rng.random() < 0.02 else "ACGT") for _ in range(rng.randint(lo, hi)))
            for _ in range(n)]


def route(seqs, n_chunks):
    """drain_batch: batch b (1000 reads) -> chunk b mod n; the partial last batch follows the same rule."""
    out = [[] for _ in range(n_chunks)]
    for i, s in enumerate(seqs):
        out[(i // 1000) % n_chunks].append(s)
    return out


def run(harness, tmp, args):
    d = tmp / "dump"
    d.mkdir(exist_ok=True)
    r = subprocess.run(harness + ["--dump", str(d)] + [str(a) for a in args], capture_output=True, text=True)
    return r, d


def read_chunks(d, n):
    return [open(d / f"chunk_{c}.txt").read().split("\n")[:-1] if os.path.getsize(d / f"chunk_{c}.txt") else []
            for c in range(n)]


@pytest.mark.parametrize("chunks,n,gz,buf", [(0, 2500, False, 0), (3, 7300, True, 0), (10, 25999, False, 4096),
                                            (4, 4000, True, 1 << 16), (7, 999, False, 0)])
def test_routing_matches_rule_and_oracle(harness, oracle, tmp_path, chunks, n, gz, buf):
    seqs = rand_seqs(n, seed=n)
    fq = tmp_path / ("r.fastq.gz" if gz else "r.fastq")
    write_fastq(fq, seqs, gz=gz)
    args = ["--chunks", chunks, fq] + (["--buffer-bytes", buf] if buf else [])
    r, d = run(harness, tmp_path, args)
    assert r.returncode == 0, r.stderr
    nc = max(1, chunks)
    got = read_chunks(d, nc)
    assert got == route(seqs, nc)
    n_reads, n_bases = map(int, open(d / "counts.txt").read().split())
    assert (n_reads, n_bases) == (n, sum(map(len, seqs)))
    # the oracle (restating src/io.rs) routes the same file the same way
    orc = oracle.Run(21, chunks, 100)
    orc.read_fastq(str(fq))
    orc.finish_ingest()
    for c in range(nc):
        assert orc.chunk_totals(c)[0] == len(got[c])
        assert orc.chunk_totals(c)[1] == sum(len(s) - s.count("N") for s in got[c])
    assert (orc.n_reads_read, orc.n_bases_read) == (n_reads, n_bases)


def test_max_reads_multiple_files_and_state_across_files(harness, oracle, tmp_path):
    a, b = rand_seqs(1500, 1), rand_seqs(1700, 2)
    fa, fb = tmp_path / "a.fastq", tmp_path / "b.fastq.gz"
    write_fastq(fa, a)
    write_fastq(fb, b, gz=True)
    r, d = run(harness, tmp_path, ["--chunks", 3, "-m", 2750, fa, fb])
    assert r.returncode == 0, r.stderr
    assert read_chunks(d, 3) == route((a + b)[:2750], 3)  # batches continue across files (io.rs:498-512)
    orc = oracle.Run(21, 3, 100)
    assert orc.read_fastq(str(fa), 2750) == 0 and orc.read_fastq(str(fb), 2750) == 1
    orc.finish_ingest()
    assert orc.n_reads_read == 2750


def test_crlf_and_missing_final_newline(harness, tmp_path):
    seqs = rand_seqs(50, 3, 1, 80)
    fq = tmp_path / "c.fastq"
    write_fastq(fq, seqs, crlf=True, trailing_newline=False)
    r, d = run(harness, tmp_path, [fq])
    assert r.returncode == 0, r.stderr
    assert read_chunks(d, 1) == [seqs]


def test_gzip_magic_without_extension_and_single_member(harness, tmp_path):
    a, b = rand_seqs(30, 4, 1, 50), rand_seqs(30, 5, 1, 50)
    fq = tmp_path / "noext"
    write_fastq(tmp_path / "a.gz", a, gz=True)
    write_fastq(tmp_path / "b.gz", b, gz=True)
    with open(fq, "wb") as f:  # two concatenated members: GzDecoder (not Multi) reads only the first (io.rs:619-621)
        f.write(open(tmp_path / "a.gz", "rb").read() + open(tmp_path / "b.gz", "rb").read())
    r, d = run(harness, tmp_path, [fq])
    assert r.returncode == 0, r.stderr
    assert read_chunks(d, 1) == [a]


def test_paired_interleave_and_quirks(harness, oracle, tmp_path):
    r1, r2 = rand_seqs(1200, 6, 1, 60), rand_seqs(1200, 7, 1, 60)
    f1, f2 = tmp_path / "R1.fastq", tmp_path / "R2.fastq"
    write_fastq(f1, r1)
    write_fastq(f2, r2)
    inter = [s for p in zip(r1, r2) for s in p]
    r, d = run(harness, tmp_path, ["--paired", "--chunks", 2, f1, f2])
    assert r.returncode == 0, r.stderr
    assert read_chunks(d, 2) == route(inter, 2)
    # odd max_reads is rounded up to even (io.rs:483-485)
    r, d = run(harness, tmp_path, ["--paired", "--chunks", 2, "-m", 1001, f1, f2])
    assert read_chunks(d, 2) == route(inter[:1002], 2)
    # R1 shorter than R2: the EOF probe of R2 ingests exactly one extra R2 record (io.rs:653-657)
    write_fastq(f1, r1[:10])
    r, d = run(harness, tmp_path, ["--paired", f1, f2])
    want = [s for p in zip(r1[:10], r2[:10]) for s in p] + [r2[10]]
    assert read_chunks(d, 1) == [want]
    orc = oracle.Run(21, 0, 100)
    orc.read_fastq_paired(str(f1), str(f2))
    orc.finish_ingest()
    assert orc.n_reads_read == 21
    # R2 shorter than R1: R1's record of the broken pair stays ingested
    write_fastq(f1, r1[:10])
    write_fastq(f2, r2[:7])
    r, d = run(harness, tmp_path, ["--paired", f1, f2])
    want = [s for p in zip(r1[:7], r2[:7]) for s in p] + [r1[7]]
    assert read_chunks(d, 1) == [want]


@pytest.mark.parametrize("content,msg", [
    (">seq1\nACGT\n>seq2\nGGGG\n", "Input appears to be FASTA format, not FASTQ (record 1 starts with '>')"),
    ("r0\nACGT\n+\nIIII\n", "FASTQ record 1 has invalid header (expected '@', got 'r'): r0"),
    ("@r0\nACGT\n-\nIIII\n", "FASTQ record 1 has invalid separator line (expected '+', got '-'): -"),
    ("@r0\nACGT\n+\nIII\n", "FASTQ record 1 has mismatched sequence (4) and quality (3) lengths"),
    ("@r0\nACGT\n+\nIIII\n@r1\nAC\n", "Truncated FASTQ record at record 2"),
    ("@r0\nACGT\n+\nIIII\n@r1\n", "missing sequence line"),
])
def test_fastq_errors(harness, oracle, tmp_path, content, msg):  # io.rs:161-198, 291-318
    fq = tmp_path / "bad.fastq"
    fq.write_text(content)
    r, _ = run(harness, tmp_path, [fq])
    assert r.returncode == 1 and msg in r.stderr, r.stderr
    orc = oracle.Run(21, 0, 100)
    with pytest.raises(oracle.OracleError) as e:
        orc.read_fastq(str(fq))
    assert msg in str(e.value)


def test_only_first_record_validated_by_default(harness, tmp_path):  # io.rs:321-332
    fq = tmp_path / "v.fastq"
    fq.write_text("@r0\nACGT\n+\nIIII\nXr1\nGGCC\n-\nII\n")
    r, d = run(harness, tmp_path, [fq])
    assert r.returncode == 0 and read_chunks(d, 1) == [["ACGT", "GGCC"]]
    r, _ = run(harness, tmp_path, ["--validate-every", 1, fq])
    assert r.returncode == 1 and "FASTQ record 2 has invalid header" in r.stderr


def test_record_longer_than_a_window_and_empty_lines(harness, tmp_path):
    # a read far longer than the "tiny" reader's 700-byte window (the window must grow, then shrink
    # back), an empty sequence line (a read of length 0 is still a read) and an empty header line
    seqs = ["ACGT" * 3, "G" * 5000, "", "TTAGC", "C" * 1500, "A"]
    fq = tmp_path / "long.fastq"
    txt = "".join(f"@r{i}\n{s}\n+\n{'I' * len(s)}\n" for i, s in enumerate(seqs))
    txt += "\nACGTAC\n+\nIIIIII\n"  # record 7: empty header; only record 1 is validated by default
    fq.write_text(txt)
    r, d = run(harness, tmp_path, ["--chunks", 2, fq])
    assert r.returncode == 0, r.stderr
    assert read_chunks(d, 2) == route(seqs + ["ACGTAC"], 2)
    n_reads, n_bases = map(int, open(d / "counts.txt").read().split())
    assert (n_reads, n_bases) == (7, sum(map(len, seqs)) + 6)


def test_empty_and_newline_only_files(harness, tmp_path):
    fq = tmp_path / "empty.fastq"
    fq.write_bytes(b"")
    r, d = run(harness, tmp_path, [fq])
    assert r.returncode == 0 and open(d / "counts.txt").read().split() == ["0", "0"]
    fq.write_bytes(b"\n")  # one empty line: a header without its record
    r, _ = run(harness, tmp_path, [fq])
    assert r.returncode == 1 and "Truncated FASTQ record at record 1" in r.stderr and "missing sequence line" in r.stderr
