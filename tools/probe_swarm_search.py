"""One-GPU probe of the search a rank runs per swarm round at N robots: R*B queries against a
1M/R-row shard with k = nb_best_matches + B (the intra-robot window), CUDA-event time per call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cslam_b200.nns_matching import NearestNeighborsMatching

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for R in (2, 4, 8):
    rows, B, dim = 1000000 // R, 64, 512
    nn = NearestNeighborsMatching(device=0)
    for s in range(0, rows, 125000):
        x = torch.rand((min(125000, rows - s), dim), generator=g, device=dev)
        nn.add_items_device(x / x.norm(dim=1, keepdim=True))
    q = torch.rand((R * B, dim), generator=g, device=dev)
    q = q / q.norm(dim=1, keepdim=True)
    for k in (30, 94):
        for _ in range(3):
            nn.search_batch_device(q, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            nn.search_batch_device(q, k)
        e1.record()
        torch.cuda.synchronize()
        print(f"R={R}: {R * B} queries x {rows} rows, k={k}: {e0.elapsed_time(e1) / 10:.3f} ms per search, "
              f"coarse {nn.last_timing()[0]:.3f} ms, info {nn.last_info.tolist()}", flush=True)
