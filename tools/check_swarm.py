"""Run under torchrun on >= 2 GPUs: the NCCL multi-robot path (cslam_b200/swarm.py, replacing the
DDS broadcasts of cslam/global_descriptor_loop_closure_detection.py:198-289,407-433) against the
ORACLE: one `oracle.nns.NNSOracle` pool per robot on rank 0, driven with the reference's
per-keyframe sequence (cslam/loop_closure_sparse_matching.py:36-92).  Also checks that every rank
ends up with the identical candidate table.  Prints SWARM_OK on success."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from oracle.nns import NNSOracle
from cslam_b200.swarm import SwarmExchange, SwarmLoopClosureMatching

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, D, rounds, thr = 16, 128, 4, 0.8
params = {'robot_id': rank, 'max_nb_robots': world, 'frontend.similarity_threshold': thr,
          'frontend.nb_best_matches': 5, 'frontend.enable_intra_robot_loop_closures': True,
          'frontend.enable_sparsification': True, 'frontend.sensor_type': 'stereo',
          'evaluation.enable_sparsification_comparison': False}
rng = np.random.default_rng(5)
places = rng.random((10, D))
stream = []
for t in range(rounds):
    x = places[rng.integers(0, 10, (world, B))] + 0.4 * rng.random((world, B, D))
    x /= np.linalg.norm(x, axis=2, keepdims=True)
    stream.append(x.astype(np.float32))

sw = SwarmLoopClosureMatching(params, SwarmExchange(), exchange_k=3)
edges, intra = [], []
for t, x in enumerate(stream):
    e, i = sw.step(torch.from_numpy(x[rank]).to(dev), list(range(t * B, (t + 1) * B)))
    edges.extend(tuple(v) for v in e)
    intra.extend((kf, ids) for kf, ids, _ in i)

ok = True
if rank == 0:
    pools = [NNSOracle() for _ in range(world)]
    exp_edges, exp_intra = [], []
    for t, x in enumerate(stream):
        before = pools[0].n
        for r in range(world):
            for b in range(B):
                pools[r].add_item(x[r, b], t * B + b)
        for q in range(world):
            for b in range(B):
                for g in range(world):
                    if g == q:
                        continue
                    kf, sim = pools[g].search_best(x[q, b])
                    if sim >= thr:
                        exp_edges.append((q, t * B + b, g, kf, float(sim)))
        for b in range(B):
            ids, sims = pools[0].search(x[0, b], 5 + B)
            keep = [i for i in ids if i < before + b][:5]
            exp_intra.append((t * B + b, keep))
    ok = [e[:4] for e in edges] == [e[:4] for e in exp_edges] and \
        np.allclose([e[4] for e in edges], [e[4] for e in exp_edges], atol=1e-6) and intra == exp_intra
    print(f"edges {len(edges)} expected {len(exp_edges)} intra {len(intra)}")
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
# identical candidate tables on every rank: a digest of (key, weight) in key order
import hashlib
cand = sw.candidate_selector.candidate_edges
keys = sorted(cand.keys())
h = hashlib.sha256(repr([(k, float(cand[k].weight).hex()) for k in keys]).encode()).digest()
dig = torch.tensor([int.from_bytes(h[:7], "big"), len(keys)], device=dev, dtype=torch.int64)
lo, hi = dig.clone(), dig.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    same = bool((lo == hi).all())
    print(f"world {world}: candidate tables identical on all ranks: {same} ({len(keys)} candidates)")
    print("SWARM_OK" if int(flag) == 1 and same and len(edges) > 0 else "SWARM_FAIL")
dist.destroy_process_group()
