#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 k-mer counting engine.

Metric (BASELINE.json): k-mers counted per second, whole job.
Workload at N=1 (BASELINE.json configs[1], SURVEY.md §8d "C2"): k=21, incremental
counting with 10 chunks over 10 M synthetic 150 bp reads (50 Mbp genome, 1 %
substitutions, 0.1 % N, seed 2), 1 B200.  At N>1 the same per-GPU load is kept
(weak scaling): 10 M reads and 50 Mbp of genome per GPU, table sharded by k-mer
hash, k-mers routed with an NCCL all-to-all.

A "step" = one full pass of the hot path over the whole input:
    reset table -> pack (ASCII -> 2-bit) -> per chunk: extract + insert -> histogram snapshot
`value`  : inputs already resident in HBM (device-generated reads).
`e2e`    : the same job through the C-ABI call a host makes (skm_ingest_batch with
           pinned HOST buffers): H2D copies and the D2H of the histograms are inside
           the timed region.
Timing: CUDA events on the stream the engine launches on (a torch stream handed to
the ctx), barrier + synchronize on both sides, max over ranks.  Inputs (1.5 GB) and
table (8.6 GB) are far larger than L2 (126 MB), so no explicit L2 flush is needed.

`--impl reference` times the reference's CPU algorithm (the C oracle port — the
Rust reference cannot be built here) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

K = 21
CHUNKS = 10
HISTO_MAX = 10000
READ_LEN = 150
READS_PER_GPU = 10_000_000
GENOME_PER_GPU = 50_000_000
SUB_RATE = 0.01
N_RATE = 0.001
SEED = 2
DISTINCT_HINT_PER_GPU = 320_000_000
CPU_SAMPLE_READS = 300_000
# BASELINE.json configs (SURVEY.md §8d fixes the generators).  `--config`; the default and the driver's runs are C2.
#   weak: reads / genome / distinct hint are PER GPU; strong: totals, divided over the GPUs
CONFIGS = {
    "C2": dict(k=21, chunks=10, scaling="weak", reads=10_000_000, genome=50_000_000, sub=0.01, n=0.001, seed=2,
               hint=320_000_000, bufs=10, golden="C2",
               what="k=21 incremental counting, 10 chunks over 10 M synthetic 150 bp reads with 1 % error (per GPU)"),
    "C1": dict(k=31, chunks=1, scaling="strong", reads=1_000_000, genome=10_000_000, sub=0.01, n=0.001, seed=1,
               hint=45_000_000, bufs=1, golden="C1_chunks1",
               what="sharkmer -k 31 --max-reads 1000000 on 1 M synthetic 150 bp reads (golden counts + histogram)"),
    "C3": dict(k=31, chunks=0, scaling="strong", reads=100_000_000, genome=500_000_000, sub=0.01, n=0.0, seed=3,
               hint=4_700_000_000, bufs=10, golden=None,
               what="k=31, 100 M synthetic 150 bp reads (~15 Gbp), hash-sharded table across the GPUs"),
    "C5": dict(k=31, chunks=0, scaling="strong", reads=1_000_000_000, genome=8_000_000_000, sub=0.0005, n=0.0, seed=5,
               hint=10_500_000_000, bufs=64, bufs_per_round=8, golden=None,
               what="high-diversity stress: 1 B synthetic 150 bp reads, ~1e10 distinct k-mers, tables near HBM capacity; "
                    "the input is counted in rounds (skm_mg_flush) because its k-mer lists do not fit the GPUs at once"),
}
CONFIG = "C2"
N_BUF = CHUNKS          # input buffers per GPU (one per chunk when chunks > 0)
BUFS_PER_ROUND = 0      # > 0: flush the sharded engine after this many buffers (inputs larger than memory)
SCALING = "weak"
GOLDEN = "C2"

B_SURVEY_PER_KMER = 64.0   # 32 B sector in + 32 B sector out per k-mer occurrence (random-sector model)
B_SURVEY_PER_BASE = 1.5    # 1 B ASCII read + 0.25 B packed write + 0.25 B packed read
# Algorithmic bytes of each kernel of the tiled path, per unit of work:
#   bucket_scatter_kernel : 0.375 B per position read (2-bit code + break bit) + 8 B per k-mer written
#   tile_sort_kernel      : 8 B per k-mer read + 8 B per k-mer written (in place, tile by tile)
#   tile_insert_kernel    : 8 B per k-mer read + 16 B per slot written (+ 16 B per slot read unless the table is new)
#   extract_insert_kernel : 64 B per k-mer (one random 32-byte sector in and out) + 0.375 B per position (direct mode)
# DRAM bytes per launch measured by `ncu --set full` (profiles/README.md), per unit, for `traffic`:
NCU_TRAFFIC = {}  # filled from profiles/r02_traffic.json when present


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def _nvml_loop(self, period_s):
        """In-process NVML sampling (the library nvidia-smi itself reads): the same fields, one handle kept open."""
        import pynvml as nv
        h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
        slow = getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8))
        therm_hw = getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40))
        therm_sw = getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20))
        pcap = getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                r = int(get_reasons(h))
                f = lambda bit: "Active" if r & bit else "Not Active"
                self.rows.append(f"{self.gpu}, {sm}, {mx}, {pw:.2f}, 0x{r:016x}, {f(slow)}, {f(therm_hw)}, {f(therm_sw)}, {f(pcap)}")
            except Exception:
                pass
            self._stop.wait(period_s)

    def start(self):
        period_ms = os.environ.get("SKM_SAMPLER_MS", "100")
        # default: NVML in-process (r2_43: the nvidia-smi child process perturbed 6 of 10 end-to-end steps by 2-24 ms,
        # the in-process reader none); SKM_SAMPLER=smi runs the recipe's nvidia-smi -lms loop instead
        if os.environ.get("SKM_SAMPLER", "nvml") == "nvml":
            try:
                import pynvml as nv
                nv.nvmlInit()
                self._stop = threading.Event()
                self.source = "nvml (in-process, the library behind nvidia-smi)"
                self.t = threading.Thread(target=self._nvml_loop, args=(float(period_ms) / 1e3,), daemon=True)
                self.t.start()
                self.proc = "nvml"
                return
            except Exception:
                self.proc = None
        try:
            self.source = "nvidia-smi -lms " + period_ms
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", os.environ.get("SKM_SAMPLER_MS", "100"),
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        if self.proc == "nvml":
            self._stop.set()
            self.t.join(timeout=2)
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples drawing the most power
        order = np.argsort(power)[len(power) // 2:]
        return {"sm_mhz": float(np.median(np.asarray(sm)[order])), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(power)), "samples": len(sm), "reasons": sorted(reasons),
                "source": getattr(self, "source", "nvidia-smi")}


# ----------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port), bounded sample
# ----------------------------------------------------------------------------

def sample_genome(n_reads):
    """Genome length for a bounded sample at the SAME depth as the full workload: a sample
    taken from the full-size genome would be almost all first-time inserts, a different regime."""
    return max(1000, GENOME_PER_GPU * n_reads // READS_PER_GPU)


def cpu_sample(n_reads, genome_len):
    """One bounded CPU pass over n_reads reads of the workload's generator, k=21, 10 chunks."""
    from oracle import oracle as o
    reads = o.synth_reads(SEED, genome_len, READ_LEN, SUB_RATE, N_RATE, 0, n_reads)
    t0 = time.perf_counter()
    run = o.Run(K, CHUNKS, HISTO_MAX)
    run.push_lines(reads)
    run.finish()
    dt = time.perf_counter() - t0
    return run, reads, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.cpu_sample_reads
    genome = sample_genome(n)
    times, kmers = [], 0
    for i in range(args.warmup + args.steps):
        run, _, dt = cpu_sample(n, genome)
        kmers = run.n_kmers_ingested
        if i >= args.warmup:
            times.append(dt)
        del run
    t = float(np.mean(times))
    v = kmers / t
    sample = (f"{n} reads of the workload's generator at the workload's depth ({genome} bp genome = 30x, same error "
              f"rates and seed; k={K}, chunks={CHUNKS}); full algorithm: per-chunk tables, ordered merge, incremental "
              f"histogram; one pass per step")
    cfg = workload_config(args.gpus)   # the repo arm's config (contract); what was actually run is `sample` below
    line = {
        "impl": "reference", "metric": "kmers_counted_per_sec", "value": v, "unit": "kmers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": SCALING, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": v, "unit": "kmers/s", "cores": 1, "kind": "port", "sample": sample,
                         "note": "C oracle port of sharkmer src/kmer + src/io.rs (Rust reference cannot be "
                                 "built here: no cargo); the reference counts on 1 thread regardless of -t"},
        "e2e": {"value": v, "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "sample": {"what": "each step runs a BOUNDED SAMPLE of config.workload, not the whole workload",
                   "reads": n, "genome_len": genome, "depth": "30x, as in the workload", "kmers_per_step": int(kmers)},
    }
    emit(line)
    return 0


def workload_config(n_gpus):
    per = 1 if SCALING == "strong" else n_gpus
    reads, genome = READS_PER_GPU * per if SCALING == "weak" else READS_PER_GPU, GENOME_PER_GPU * per if SCALING == "weak" else GENOME_PER_GPU
    return {"workload": f"{CONFIG}: k={K}, {CHUNKS} chunks, {reads} synthetic {READ_LEN} bp reads "
                        f"({genome} bp genome, {SUB_RATE} sub, {N_RATE} N, seed {SEED})",
            "baseline_config": CONFIGS[CONFIG]["what"],
            "k": K, "chunks": CHUNKS, "reads": reads, "read_len": READ_LEN,
            "genome_len": genome, "histo_max": HISTO_MAX,
            "l2": "inputs and table are GBs, far beyond the 126 MB L2; no flush needed",
            "parallelism": "1 GPU" if n_gpus == 1 else f"{n_gpus} GPUs: reads split, table sharded by k-mer hash range, "
                           "k-mers routed to their owners over NVLink by the copy engines"}


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    from sharkmer_b200 import _lib, kmer
    from sharkmer_b200.common import rate_to_thresh

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    if world > 1 and os.environ.get("SKM_NUMA_BIND", "1") != "0":
        # before the CUDA context and every pinned allocation: keep this rank's host buffers on the
        # NUMA node of its GPU
        from sharkmer_b200.multigpu import bind_to_gpu_numa
        cpus = bind_to_gpu_numa(local_rank)
        print(f"[rank {rank}] bound to {len(cpus) if cpus else 0} CPUs local to GPU {local_rank}"
              + (f" ({cpus[0]}..{cpus[-1]})" if cpus else ""), file=sys.stderr)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    scale = args.reads_per_gpu / READS_PER_GPU           # (--reads-per-gpu: experiments at reduced size)
    if SCALING == "weak":
        n_reads_total = args.reads_per_gpu * world
        genome = int(GENOME_PER_GPU * world * scale)
        hint = int(DISTINCT_HINT_PER_GPU * scale)
    else:   # strong: the configuration fixes the totals
        n_reads_total = int(READS_PER_GPU * scale) // (1000 * world) * (1000 * world)
        genome = int(GENOME_PER_GPU * scale)
        hint = int(DISTINCT_HINT_PER_GPU * scale / world * 1.03)
    genome = max(genome, 1000)
    mode = {"auto": _lib.INSERT_AUTO, "direct": _lib.INSERT_DIRECT, "partitioned": _lib.INSERT_PARTITIONED}[args.mode]

    if world > 1 and os.environ.get("SKM_TRACE"):
        os.environ["SKM_TRACE"] += f".rank{rank}"   # one timeline file per rank
    # high priority: the inserts (this stream) outrank the engine's bucketing streams
    stream = torch.cuda.Stream(device=dev, priority=-1 if os.environ.get("SKM_PRIO", "1") != "0" else 0)
    eng = kmer.Engine(K, CHUNKS, HISTO_MAX, capacity_hint=hint, device=local_rank, insert_mode=mode,
                      n_ranks=world, rank=rank, stream=stream.cuda_stream)
    st, nt = rate_to_thresh(SUB_RATE), rate_to_thresh(N_RATE)
    line = READ_LEN + 1

    # ---- inputs: chunk c holds the reads of batches b = c (mod CHUNKS) (src/io.rs:355-361);
    #      rank r takes the r-th contiguous slice of every chunk's 1000-read batches ---------------
    def make_inputs(engine, n_total, genome_len, n_buf=None):
        """This rank's input as device buffers.  chunks > 0: buffer c = this rank's slice of chunk c's 1000-read
        batches; chunks == 0: this rank's contiguous share of the reads, cut into n_buf buffers."""
        bufs, counts = [], []
        if CHUNKS > 0:
            for c in range(CHUNKS):
                n_batches_c = len(range(c, (n_total + 999) // 1000, CHUNKS))
                lo, hi = n_batches_c * rank // world, n_batches_c * (rank + 1) // world
                first, n = lo * 1000, (hi - lo) * 1000
                # (n_total is a multiple of 1000 in every configuration used here)
                t = torch.empty(max(n * line, 1), dtype=torch.uint8, device=dev)[:n * line]
                if n:
                    engine.synth_device(SEED, genome_len, READ_LEN, st, nt, c, CHUNKS, first, n, t.data_ptr())
                bufs.append(t)
                counts.append(n)
        else:
            n_buf = n_buf or N_BUF
            n_batches = (n_total + 999) // 1000
            lo, hi = n_batches * rank // world, n_batches * (rank + 1) // world
            for b in range(n_buf):
                b0, b1 = lo + (hi - lo) * b // n_buf, lo + (hi - lo) * (b + 1) // n_buf
                first, n = b0 * 1000, (b1 - b0) * 1000
                t = torch.empty(max(n * line, 1), dtype=torch.uint8, device=dev)[:n * line]
                if n:
                    engine.synth_device(SEED, genome_len, READ_LEN, st, nt, 0, 1, first, n, t.data_ptr())
                bufs.append(t)
                counts.append(n)
        torch.cuda.synchronize()
        return bufs, counts

    d_bufs, n_local = make_inputs(eng, n_reads_total, genome)
    h_bufs = []
    if not args.no_e2e:
        for t in d_bufs:
            h = torch.empty(t.numel(), dtype=torch.uint8, pin_memory=True)
            h.copy_(t)
            h_bufs.append(h)
        torch.cuda.synchronize()
    in_bytes = sum(t.numel() for t in d_bufs)

    # ---- one step -------------------------------------------------------------------------------
    sharded = None
    cpu_group = None
    if world > 1:
        from sharkmer_b200.multigpu import ShardedCounter
        cpu_group = dist.new_group(backend="gloo")   # carries the library's all-gathers (host bytes); NCCL: barriers only
        # receive arena: this rank's share of all ranks' k-mers = ~its own input x 8 B, +6 % capped-region
        # slack, +3 % tile offsets, + margin for imbalance
        # Rounds: when lists + arena of the whole input do not fit beside the table (chunks == 0 only), the input is
        # counted in rounds with a collective flush in between (a flush empties lists and arenas).
        global BUFS_PER_ROUND
        if CHUNKS == 0:
            table_bytes = 16 * (1 << int(np.ceil(np.log2(max(hint, 1) / 0.6))))
            free_b = torch.cuda.mem_get_info(dev)[0]
            rounds = 1 if not BUFS_PER_ROUND else N_BUF // BUFS_PER_ROUND
            while rounds < N_BUF and in_bytes / rounds * (9.7 + 9.7 + 10.4) + (8 << 30) > free_b - table_bytes:
                rounds += 1
                while N_BUF % rounds:
                    rounds += 1
            BUFS_PER_ROUND = N_BUF // rounds if rounds > 1 else 0
        per_round = in_bytes * (BUFS_PER_ROUND / N_BUF if BUFS_PER_ROUND else 1.0)
        # (every source's list for this rank: 1/N of its positions x 8 B, +12 % owner imbalance, +6 % bucket slack,
        #  +3 % tile offsets, + the fixed slack of a capped list per batch)
        arena = int(per_round * 8 * 1.3) + (256 << 20) + (10 << 20) * len(d_bufs) * world
        sharded = ShardedCounter(eng, dev, cpu_group=cpu_group, arena_bytes=arena)

    dbg = os.environ.get("SKM_BENCH_DEBUG") and rank == 0

    def step(host_buffers: bool):
        t0 = time.perf_counter()
        eng.reset()
        t1 = time.perf_counter()
        r = _step(host_buffers)
        if dbg:
            print(f"[bench] reset {1e3 * (t1 - t0):.1f} ms, ingest+finalize {1e3 * (time.perf_counter() - t1):.1f} ms "
                  f"(host={host_buffers})", file=sys.stderr)
        return r

    def _step(host_buffers: bool):
        t_in = time.perf_counter()
        for i in range(len(d_bufs)):
            c = i if CHUNKS > 0 else 0
            if d_bufs[i].numel() == 0:
                continue
            if host_buffers:
                eng.ingest_ptr(c, h_bufs[i].data_ptr(), h_bufs[i].numel(), _lib.INGEST_ASYNC)
            else:
                eng.ingest_device(c, d_bufs[i].data_ptr(), d_bufs[i].numel())
            if sharded is not None and BUFS_PER_ROUND and (i + 1) % BUFS_PER_ROUND == 0 and i + 1 < len(d_bufs):
                sharded.flush()    # collective: count this round, free its lists and the arenas
        if dbg:
            print(f"[bench]   ingest calls {1e3 * (time.perf_counter() - t_in):.1f} ms", file=sys.stderr)
        if world == 1:
            eng.finalize()
            return eng.histogram(CHUNKS - 1) if CHUNKS else None
        cols = sharded.finalize()
        return cols[-1] if cols is not None else None

    per_step = []   # wall ms of every timed step of the last timed() call (rank 0's view; for spotting outliers)

    def timed(host_buffers: bool, steps: int, warmup: int):
        for _ in range(warmup):
            step(host_buffers)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        per_step.clear()
        for _ in range(steps):
            t_s = time.perf_counter()
            hist = step(host_buffers)          # (ends with a blocking read of the result)
            per_step.append(round((time.perf_counter() - t_s) * 1e3, 3))
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            tm = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms, wall = float(tm[0]), float(tm[1]) / 1e3
        return ms / steps, wall * 1e3 / steps, hist

    # ---- parity guard before timing: GPU vs oracle on a bounded sample, through the SAME path as the
    #      timed region (tiled insert; at N > 1 the sharded path: bucketing by owner, copy-engine exchange,
    #      collective finalize) ----------------------------------------------------------------------------
    cpu = None
    parity = None
    if not args.no_cpu:
        unit = 1000 * max(1, CHUNKS) * world
        n_s = max(unit, args.cpu_sample_reads // unit * unit)
        g_s = sample_genome(n_s)
        chk = kmer.Engine(K, CHUNKS, HISTO_MAX, capacity_hint=int(hint * n_s * world / max(n_reads_total, 1)) + 1000,
                          device=local_rank, insert_mode=_lib.INSERT_PARTITIONED if args.mode == "auto" else mode,
                          n_ranks=world, rank=rank)
        s_bufs, _ = make_inputs(chk, n_s, g_s, n_buf=3)
        chk_sh = None
        if world > 1:
            chk_sh = ShardedCounter(chk, dev, cpu_group=cpu_group,
                                    arena_bytes=int(sum(t.numel() for t in s_bufs) * 8 * 1.5) + (64 << 20) * world)
        for i, t in enumerate(s_bufs):
            if t.numel():
                chk.ingest_device(i if CHUNKS > 0 else 0, t.data_ptr(), t.numel())
        if world == 1:
            chk.finalize()
        else:
            chk_sh.finalize()
        digest = chk.digest()
        cols = np.stack([chk.histogram(c) for c in range(CHUNKS)]) if CHUNKS else None
        if world > 1:   # a table digest is a wrapping sum over entries: the global one is the sum of the partitions'
            dg = torch.tensor([np.uint64(digest).astype(np.int64)], device=dev, dtype=torch.int64)
            dist.all_reduce(dg)
            digest = int(np.int64(int(dg[0])).astype(np.uint64))
        if rank == 0:
            run, _, dt = cpu_sample(n_s, g_s)   # the oracle: checker + CPU baseline
            ok_digest = digest == run.table().digest()
            ok_cols = all((cols[c] == run.histogram(c)).all() for c in range(CHUNKS))   # (chunks == 0: no columns)
            parity = {"sample_reads": n_s, "ranks": world, "table_digest_equals_oracle": bool(ok_digest),
                      "histogram_columns_equal_oracle": bool(ok_cols)}
            if not (ok_digest and ok_cols):
                raise SystemExit(f"PARITY FAILURE: GPU table/histograms differ from the oracle on the sample: {parity}")
            cpu = {"value": run.n_kmers_ingested / dt, "unit": "kmers/s", "cores": 1, "kind": "port",
                   "sample": f"{n_s} reads of the workload's generator at the workload's depth ({g_s} bp genome), same "
                             f"k/chunks; {dt:.1f} s; the GPU result on this sample (same path as the timed region, "
                             f"{world} rank(s)) was checked bit-exact: table digest + {CHUNKS} histogram columns"}
            del run
        chk.close()
        del s_bufs

    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("SKM_NO_SAMPLER"):
        sampler.start()
    def global_state():
        """Totals and the table digest of the last run, summed over the partitions."""
        t = eng.totals()
        kc = sum(eng.chunk_totals(c).n_kmers for c in range(max(1, CHUNKS)))
        v = torch.tensor([int(t.n_kmers), int(t.n_unique), int(kc), int(np.uint64(eng.digest()).astype(np.int64))],
                         device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(v)
        return {"table_mass": int(v[0]), "distinct": int(v[1]), "windows_extracted": int(v[2]),
                "digest": int(np.int64(int(v[3])).astype(np.uint64))}

    ms_dev, ms_wall, hist_dev = timed(False, args.steps, args.warmup)
    steps_dev_ms = list(per_step)
    stt = eng.stage_times()
    tot = eng.totals()
    gs_dev = global_state()
    full = {"table_mass_equals_windows_extracted": gs_dev["table_mass"] == gs_dev["windows_extracted"],   # src/io.rs:1042-1047
            "histogram_support_equals_distinct": hist_dev is None or int(hist_dev[1:].sum()) == gs_dev["distinct"]}  # :1120-1132
    if not args.no_e2e:
        ms_e2e, ms_e2e_wall, hist_e2e = timed(True, args.steps, args.warmup)
        steps_e2e_ms = list(per_step)
        gs_e2e = global_state()
        full["host_buffer_run_equals_device_run"] = bool((hist_dev is None or (hist_e2e == hist_dev).all())
                                                         and gs_e2e == gs_dev)
        st2 = eng.stage_times()
        if world > 1:
            print(f"[rank {rank}] e2e h2d {st2.h2d:.1f} ms, step {ms_e2e:.1f} ms", file=sys.stderr)
    if GOLDEN and scale == 1.0 and (SCALING == "strong" or world == 1) and args.k is None and args.chunks is None:
        # the oracle's result for exactly this configuration, committed as a fixture (tests/golden/make_golden_full.py)
        try:
            gold = json.load(open(os.path.join(ROOT, "tests", "golden", "full_cases.json"))).get(GOLDEN)
        except Exception:
            gold = None
        if gold:
            full["table_digest_equals_oracle_golden"] = gs_dev["digest"] == int(gold["digest"])
            full["totals_equal_oracle_golden"] = (gs_dev["table_mass"], gs_dev["distinct"]) == (gold["n_kmers"], gold["n_unique"])
            if CHUNKS and hist_dev is not None:
                want = np.zeros(HISTO_MAX + 2, dtype=np.uint64)
                for b, v in gold["histograms"][CHUNKS - 1].items():
                    want[int(b)] = v
                full["final_histogram_equals_oracle_golden"] = bool((hist_dev == want).all())
    if not all(full.values()):
        raise SystemExit(f"PARITY FAILURE on the full-size run: {full} {gs_dev}")
    if parity is None:
        parity = {}
    parity["full_size_run"] = dict(full, **{"n_ranks": world, "table_digest": f"{gs_dev['digest']:016x}",
                                            "note": "conservation identities of src/io.rs:1042-1047,1120-1132 on the "
                                                    "totals summed over all partitions; the digest is the wrapping sum of "
                                                    "skm_pair_digest over every (k-mer, count) of every partition"})
    clocks = sampler.stop() if rank == 0 else None

    # ---- totals over ranks ------------------------------------------------------------------------
    n_kmers_local = sum(eng.chunk_totals(c).n_kmers for c in range(max(1, CHUNKS)))
    n_bases_local = in_bytes - sum(n_local)
    n_kmers = gs_dev["windows_extracted"]
    if world > 1:
        tt = torch.tensor([n_bases_local], device=dev, dtype=torch.int64)
        dist.all_reduce(tt)
        n_bases = int(tt[0])
    else:
        n_bases = n_bases_local

    if rank == 0:
        peak, peak_src = load_peaks()
        value = n_kmers / (ms_dev / 1e3)
        # per-kernel roofline: algorithmic bytes over the CUDA-event time of the kernel's launches
        slots = int(stt.table_capacity)
        kernels = {}

        def add_kernel(name, ms, launches, nbytes, what, survey_bytes=None, survey_what=None):
            """nbytes: what the kernel has to move in THIS design (streaming passes).  survey_bytes: the same work on
            SURVEY.md §8d's per-unit figure (the contract's roofline.achieved), where §8d has one for the kernel."""
            if launches and ms > 0:
                kernels[name] = {"ms_per_step": ms, "launches_per_step": int(launches), "ms_per_launch": ms / launches,
                                 "streaming_bytes_per_launch": nbytes / launches, "streaming_GBps": nbytes / (ms * 1e-3) / 1e9,
                                 "streaming_frac": nbytes / (ms * 1e-3) / 1e9 / peak, "streaming_bytes": what}
                sb, sw = (survey_bytes, survey_what) if survey_bytes is not None else (nbytes, what)
                kernels[name].update({"algorithmic_bytes_per_launch": sb / launches, "achieved": sb / (ms * 1e-3) / 1e9,
                                      "frac": sb / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": sw})
        if stt.insert_bases:   # direct mode
            add_kernel("extract_insert_kernel", stt.insert, stt.launches[4],
                       n_kmers_local * 64.0 + stt.insert_bases * 0.375,
                       "64 B per k-mer (one random 32 B sector in and out) + 0.375 B per packed position")
        else:
            add_kernel("tile_insert_kernel", stt.insert, stt.launches[4], n_kmers_local * 8.0 + slots * 16.0,
                       "8 B per k-mer (list read) + 16 B per table slot written (new table: nothing to load)",
                       survey_bytes=n_kmers_local * B_SURVEY_PER_KMER,
                       survey_what="SURVEY.md §8d: 64 B per k-mer occurrence (one 32 B sector in and out per update) x the "
                                   "k-mers one launch counts; the kernel counts in shared memory and moves far fewer DRAM "
                                   "bytes (streaming_* and traffic)")
            add_kernel("bucket_scatter_kernel", stt.partition, stt.launches[3], in_bytes * 0.375 + n_kmers_local * 8.0,
                       "0.375 B per position read (2-bit code + break bit) + 8 B per k-mer written")
            add_kernel("tile_sort_kernel", stt.sort, stt.sort_launches, n_kmers_local * 16.0,
                       "8 B per k-mer read + 8 B per k-mer written (in place)")
        add_kernel("pack_kernel", stt.pack, stt.launches[1], in_bytes * 1.375, "1 B per position read + 0.375 B written")
        kernel_name = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
        dom = kernels[kernel_name]
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            traffic = tr["dram_bytes_per_launch"].get(kernel_name)
        except Exception:
            pass
        survey_bytes = n_kmers * B_SURVEY_PER_KMER + n_bases * B_SURVEY_PER_BASE
        out = {
            "metric": "kmers_counted_per_sec", "value": value, "unit": "kmers/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": SCALING, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(world),
            "bases_per_sec": n_bases / (ms_dev / 1e3),
            "ms_per_step_wall": ms_wall, "steps_ms_wall": steps_dev_ms,
            "insert_mode": args.mode,
            "n_kmers": n_kmers, "n_distinct_rank0": int(tot.n_unique),
            "stage_ms": {"pack": stt.pack, "count": stt.count, "partition": stt.partition, "sort": stt.sort,
                         "insert": stt.insert, "histogram": stt.histogram, "grow": stt.grow,
                         "finalize": stt.total_finalize},
            "table": {"slots": slots, "bytes": int(stt.table_bytes),
                      "load": float(tot.n_unique) / float(stt.table_capacity), "grows": int(stt.n_grows),
                      "tiled_launches": int(stt.tiled_launches), "tiled_retries": int(stt.tiled_retries)},
            "roofline": {"bound": "hbm", "kernel": kernel_name,
                         "achieved": dom["achieved"], "peak": peak, "unit": "GB/s", "frac": dom["frac"],
                         "streaming_achieved": dom["streaming_GBps"], "streaming_frac": dom["streaming_frac"],
                         "streaming_bytes": dom["streaming_bytes"],
                         "traffic": traffic,
                         "traffic_GBps": (traffic / (dom["ms_per_launch"] * 1e-3) / 1e9) if traffic else None,
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel from the "
                                         "committed ncu --set full capture of this workload (profiles/r02_traffic.json)",
                         "peak_source": peak_src,
                         "launches_per_step": dom["launches_per_step"], "ms_per_launch": dom["ms_per_launch"],
                         "algorithmic_bytes": dom["algorithmic_bytes"],
                         "kernels": kernels,
                         "step_survey_model": {
                             "what": "whole step on SURVEY.md §8d's figure: 64 B per k-mer occurrence + 1.5 B per input base "
                                     "(the random-sector model the tiled design replaces with streaming passes)",
                             "achieved": survey_bytes / (ms_dev * 1e-3) / 1e9 / world,
                             "frac": survey_bytes / (ms_dev * 1e-3) / 1e9 / world / peak},
                         "kmers_per_sec_in_kernel": n_kmers_local / (max(stt.insert, 1e-6) * 1e-3)},
            "gpu_launches": int(stt.kernel_launches) * args.steps,
            "nvlink_bytes_sent_per_step_rank0": sharded.bytes_sent if sharded else 0,
            "rounds": (N_BUF // BUFS_PER_ROUND) if BUFS_PER_ROUND else 1,
            "exchange": None if world == 1 else "copy-engine peer copies over NVLink into per-source sub-arenas at ingest time "
                        "(no SM time); NCCL carries only the bench's barriers and all-reduces, gloo the library's all-gathers",
            "clocks": clocks,
        }
        if not args.no_gups:
            g = {}
            for name, var in (("load+red", 0), ("red_only", 1), ("load_only", 2), ("local_load+red", 3),
                              ("local_red_only", 4), ("local_load_only", 5), ("local_load+red32", 6)):
                ms = eng.bench_gups(int(np.log2(stt.table_capacity)), 1 << 28, 3, var)
                g[name] = (1 << 28) / (ms * 1e-3)
            out["gups"] = g
            # north-star ratio: whole-job k-mers/s per GPU over the part's measured random-access update rates
            out["roofline"]["random_access_frac"] = value / world / g["load+red"]
            out["roofline"]["random_access_frac_l2_local"] = value / world / g["local_load+red"]
            out["roofline"]["random_access_note"] = ("k-mers/s per GPU / GUPS probe (key load + RED.ADD on 16 B slots): DRAM-resident "
                                                      "uniform-random updates, and region-local (L2-resident) updates")
        out["parity"] = parity
        if not args.no_e2e:
            out["e2e"] = {"value": n_kmers / (ms_e2e / 1e3), "unit": "kmers/s",
                          "h2d_bytes_per_step": int(in_bytes) * world,
                          "d2h_bytes_per_step": int(CHUNKS * (HISTO_MAX + 2) * 8 + 64) * world,   # histogram columns + totals
                          "ms_per_step": ms_e2e, "ms_per_step_wall": ms_e2e_wall, "steps_ms_wall": steps_e2e_ms,
                          "stage_ms": {"h2d": st2.h2d, "pack": st2.pack, "insert": st2.insert, "histogram": st2.histogram},
                          "api": "skm_ingest_batch(pinned host buffers) x10 -> skm_finalize -> skm_histogram"}
        if world == 1 and not args.no_services:
            # the table services the consumers call: one oligo scan = one streaming pass over the table
            # (16 B per slot); batched canonical lookups = random 16 B probes, through the C ABI with host buffers
            rng = np.random.default_rng(1)
            oligos = rng.integers(0, 1 << 30, size=64, dtype=np.uint64)
            eng.scan_oligos(oligos, 15, 2)
            s0 = eng.stage_times().scan
            for _ in range(3):
                eng.scan_oligos(oligos, 15, 2)
            scan_ms = (eng.stage_times().scan - s0) / 3
            sample = d_bufs[0][:30_000 * line].cpu().numpy()
            q = eng.extract_kmers(sample)
            q = q[q != np.uint64(0xFFFFFFFFFFFFFFFF)]
            eng.lookup(q[:1000], 0, 0)
            t0 = time.perf_counter()
            cnt, found = eng.lookup(q, 0, 0)
            lk_s = time.perf_counter() - t0
            out["services"] = {
                "scan_oligos": {"ms_per_scan": scan_ms, "GB_per_s": 16.0 * slots / (scan_ms * 1e-3) / 1e9,
                                "frac_of_hbm_peak": 16.0 * slots / (scan_ms * 1e-3) / 1e9 / peak,
                                "what": "find_oligos_in_kmers (src/pcr/primers.rs:163-226) as one table pass, 64 oligos of 15 bases, "
                                        "kernel time (CUDA events)"},
                "lookup_batch": {"queries": int(q.size), "found": int(found.sum()), "lookups_per_s": q.size / lk_s,
                                 "what": "get_canonical_count (src/kmer/counting.rs:205-209) for every k-mer of 30 k reads in ONE call, "
                                         "host buffers in and out, wall time of the call"}}
            if "gups" in out:
                out["services"]["lookup_batch"]["frac_of_random_load_rate"] = q.size / lk_s / out["gups"]["load_only"]
        if cpu:
            out["cpu_baseline"] = cpu
        emit(out)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


class _CudaArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}


def _as_tensor(ptr, n, dev):
    import torch
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=dev)
    return torch.as_tensor(_CudaArray(ptr, n), device=dev)


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries (NCCL's version banner, torchrun notices) print to stdout; keep fd 1 clean for the
    # single JSON line by pointing it at stderr until the result is ready.
    global _REAL_STDOUT
    global CHUNKS, K, CONFIG, READS_PER_GPU, GENOME_PER_GPU, SUB_RATE, N_RATE, SEED, DISTINCT_HINT_PER_GPU, N_BUF
    global BUFS_PER_ROUND, SCALING, GOLDEN
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "direct", "partitioned"])
    ap.add_argument("--reads-per-gpu", type=int, default=READS_PER_GPU)
    ap.add_argument("--cpu-sample-reads", type=int, default=CPU_SAMPLE_READS)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gups", action="store_true", help="skip the random-access roofline probe")
    ap.add_argument("--no-services", action="store_true", help="skip timing the table services (oligo scan, batched lookups)")
    ap.add_argument("--gups", action="store_true", help="(default now; kept for old command lines)")
    ap.add_argument("--exchange", default="dma", help="(ignored: the exchange is the copy-engine path; kept for old command lines)")
    ap.add_argument("--chunks", type=int, default=None, help="EXPERIMENT ONLY: override the workload's chunk count")
    ap.add_argument("--k", type=int, default=None, help="EXPERIMENT ONLY: override k")
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS),
                    help="BASELINE.json configuration (default C2: the metric's configuration and the driver's runs)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    CONFIG, K, CHUNKS, SCALING, GOLDEN = args.config, cfg["k"], cfg["chunks"], cfg["scaling"], cfg["golden"]
    READS_PER_GPU, GENOME_PER_GPU, DISTINCT_HINT_PER_GPU = cfg["reads"], cfg["genome"], cfg["hint"]
    SUB_RATE, N_RATE, SEED = cfg["sub"], cfg["n"], cfg["seed"]
    N_BUF, BUFS_PER_ROUND = cfg["bufs"], cfg.get("bufs_per_round", 0)
    if args.reads_per_gpu == 10_000_000 and args.config != "C2":
        args.reads_per_gpu = READS_PER_GPU      # (the option's default is C2's)
    if args.chunks is not None:
        CHUNKS = args.chunks
        N_BUF = max(N_BUF, CHUNKS)
    if args.k is not None:
        K = args.k
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
