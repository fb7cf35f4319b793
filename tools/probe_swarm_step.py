"""Under torchrun: where the time of a SwarmLoopClosureMatching.step goes (rank 0), with CUDA
events between the sub-stages and host wall clocks around the whole call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
import argparse

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from cslam_b200.swarm import SwarmExchange, SwarmLoopClosureMatching
args = argparse.Namespace(k=30, mac_budget=1000, dim=512, backbone="resnet18", precision="fp32")
params = bench.frontend_params(args, rank, world)
sw = SwarmLoopClosureMatching(params, SwarmExchange(), exchange_k=30)
g = torch.Generator(device=dev).manual_seed(rank)
shard = 1000000 // world
for s in range(0, shard, 125000):
    m = min(125000, shard - s)
    x = torch.rand((m, 512), generator=g, device=dev)
    sw.local_nnsm.add_items_device(x / x.norm(dim=1, keepdim=True), range(s, s + m))
    sw._append_ids(range(s, s + m), dev)
B = 64
emb = torch.rand((B, 512), generator=g, device=dev)
emb = emb / emb.norm(dim=1, keepdim=True)

# monkeypatch pieces with event timers
ex = sw.exchange
marks = []
def timed(name, fn):
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); r = fn(*a, **k); e1.record(); t1 = time.perf_counter()
        marks.append((name, e0, e1, (t1 - t0) * 1e3)); return r
    return w
ex.all_gather_descriptors = timed("all_gather_descriptors", ex.all_gather_descriptors)
ex.all_gather_topk = timed("all_gather_topk", ex.all_gather_topk)
sw.local_nnsm.add_items_device = timed("add_items_device", sw.local_nnsm.add_items_device)
sw.local_nnsm.search_batch_device = timed("search_batch_device", sw.local_nnsm.search_batch_device)
sw._filter_round = timed("filter_round", sw._filter_round)
sw.candidate_selector.add_matches = timed("add_matches", sw.candidate_selector.add_matches)
kf = shard
for it in range(8):
    marks.clear()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    sw.step(emb, list(range(kf, kf + B))); kf += B
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    if rank == 0 and it >= 5:
        print(f"step: {e0.elapsed_time(e1):.3f} ms on the stream, {(t1 - t0) * 1e3:.3f} ms wall | " +
              " | ".join(f"{n} {a.elapsed_time(b):.3f} (host {h:.3f})" for n, a, b, h in marks), flush=True)
dist.destroy_process_group()
