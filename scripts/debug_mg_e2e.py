"""Debug: timeline of the multi-GPU e2e step (host timestamps on rank 0)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from sharkmer_b200 import kmer, _lib, multigpu
from oracle import oracle as o

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(device=dev)
K, CH, HM, L = 21, 10, 10000, 150
n_total = 10_000_000 * world
eng = kmer.Engine(K, CH, HM, capacity_hint=320_000_000, device=lr, n_ranks=world, rank=rank, stream=stream.cuda_stream)
st, nt = o.rate_to_thresh(0.01), o.rate_to_thresh(0.001)
d_bufs, h_bufs = [], []
for c in range(CH):
    nb = len(range(c, n_total // 1000, CH)); lo, hi = nb * rank // world, nb * (rank + 1) // world
    t = torch.empty((hi - lo) * 1000 * (L + 1), dtype=torch.uint8, device=dev)
    eng.synth_device(2, 50_000_000 * world, L, st, nt, c, CH, lo * 1000, (hi - lo) * 1000, t.data_ptr())
    d_bufs.append(t)
    h = torch.empty(t.numel(), dtype=torch.uint8, pin_memory=True); h.copy_(t); h_bufs.append(h)
torch.cuda.synchronize()
arena = int(1.3 * max(t.numel() for t in d_bufs)) + (1 << 20)
sc = multigpu.ShardedCounter(eng, CH, CH, HM, dev, stream=stream, exchange="p2p", arena_entries=arena)

orig_snap = eng.snapshot_histogram
marks = []
def snap(c):
    orig_snap(c); marks.append((f"snap{c}", time.perf_counter()))
eng.snapshot_histogram = snap
orig_rcd = eng.route_count_device
def rcd(c):
    marks.append((f"route{c}_begin", time.perf_counter()))
    r = orig_rcd(c)
    if c == 0:
        marks.append(("r0_count_queued", time.perf_counter()))
        torch.cuda.current_stream().synchronize()
        marks.append(("r0_count_done", time.perf_counter()))
    return r
eng.route_count_device = rcd
orig_ag = dist.all_gather_into_tensor
def ag(*a, **k):
    r = orig_ag(*a, **k)
    marks.append(("ag_queued", time.perf_counter()))
    torch.cuda.current_stream().synchronize()
    marks.append(("ag_done", time.perf_counter()))
    return r
dist.all_gather_into_tensor = ag
orig_sc = eng.route_scatter_p2p
def scp(c, slot, off):
    r = orig_sc(c, slot, off)
    if c == 0:
        marks.append(("r0_scatter_queued", time.perf_counter()))
        torch.cuda.current_stream().synchronize()
        marks.append(("r0_scatter_done", time.perf_counter()))
    return r
eng.route_scatter_p2p = scp

for it in range(4):
    marks.clear()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.reset()
    marks.append(("reset", time.perf_counter()))
    for c in range(CH):
        eng.ingest_ptr(c, h_bufs[c].data_ptr(), h_bufs[c].numel(), _lib.INGEST_ASYNC)
    marks.append(("ingest_queued", time.perf_counter()))
    sc.finalize()
    marks.append(("done", time.perf_counter()))
    if rank == 0 and it >= 2:
        print(" ".join(f"{n}={1e3*(t-t0):.1f}" for n, t in marks), flush=True)
dist.destroy_process_group()
