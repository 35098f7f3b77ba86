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
sc = multigpu.ShardedCounter(eng, CH, CH, HM, dev, stream=stream, exchange="dma", arena_entries=arena)

marks = []
def wrap(obj, name, label):
    orig = getattr(obj, name)
    def f(*a, **k):
        t = time.perf_counter(); r = orig(*a, **k); marks.append((label, t, time.perf_counter())); return r
    setattr(obj, name, f)
wrap(eng, "route_count", "cnt"); wrap(eng, "route_scatter_dma", "dma"); wrap(eng, "dma_wait", "wait")
wrap(eng, "insert_runs_device", "ins"); wrap(dist, "barrier", "bar"); wrap(dist, "all_gather_into_tensor", "ag")
MODE = os.environ.get("DBG_MODE", "value")
for it in range(4):
    marks.clear()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.reset()
    for c in range(CH):
        if MODE == "e2e":
            eng.ingest_ptr(c, h_bufs[c].data_ptr(), h_bufs[c].numel(), _lib.INGEST_ASYNC)
        else:
            eng.ingest_device(c, d_bufs[c].data_ptr(), d_bufs[c].numel())
    marks.clear()
    sc.finalize()
    t1 = time.perf_counter()
    if rank == 0 and it >= 2:
        print(f"total={1e3*(t1-t0):.1f} " + " ".join(f"{n}@{1e3*(a-t0):.1f}+{1e3*(b-a):.2f}" for n, a, b in marks), flush=True)
dist.destroy_process_group()
