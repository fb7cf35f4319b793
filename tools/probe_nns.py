"""GPU probe: NNS at a given pool size; prints timings and cross-checks the tensor-core
path against the exact fp64 scan kernel."""
import argparse, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cslam_b200.nns_matching import NearestNeighborsMatching

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--d", type=int, default=512)
ap.add_argument("--q", type=int, default=64)
ap.add_argument("--k", type=int, default=30)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--check", type=int, default=4)
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(2)
nn = NearestNeighborsMatching(device=0)
chunk = 100000
for s in range(0, a.n, chunk):
    m = min(chunk, a.n - s)
    x = torch.rand((m, a.d), generator=g, device=dev)
    x = x / x.norm(dim=1, keepdim=True)
    nn.add_items_device(x, items=None if s == 0 else None)
q = torch.rand((a.q, a.d), generator=g, device=dev, dtype=torch.float64)
q = q / q.norm(dim=1, keepdim=True)
torch.cuda.synchronize()
for r in range(a.reps):
    idx, sims = nn.search_batch_device(q, a.k)
    torch.cuda.synchronize()
    c, nl, t = nn.last_timing()
    byts = a.n * nn.dim * 2 if nn.dim % 64 == 0 else a.n * ((nn.dim + 63) // 64 * 64) * 2
    print(f"rep {r}: coarse {c*1e3:.1f} us ({byts/c/1e6:.0f} GB/s)  total {t*1e3:.1f} us  coarse_launches {nl} info {nn.last_info.tolist()}")
if a.check:
    nn.set_mode(1)
    t0 = time.time()
    idx2, sims2 = nn.search_batch_device(q[:a.check], a.k)
    torch.cuda.synchronize()
    print("exact scan time", time.time() - t0, "match", bool((idx[:a.check] == idx2).all()),
          "max dsim", float((sims[:a.check] - sims2).abs().max()))
