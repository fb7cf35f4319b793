"""Per-kernel timings on one B200 with CUDA events (20 reps after warm-up), algorithmic bytes per
SURVEY.md section 8(d), and the fraction of the measured HBM copy peak.  One JSON line per kernel.
Covers BASELINE.json configs[1] (NetVLAD/VGG16 640x480 batch 64, descriptor extraction only)
and the NNS regimes of configs[2]."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cslam_b200.nns_matching import NearestNeighborsMatching
from cslam_b200.vpr._common import Preprocessor, backbone_precision
from cslam_b200.vpr.cosplace import GemHead, get_backbone
from cslam_b200.vpr.netvlad import NetVLADLayer, PCAProjection, build_vgg16_encoder

dev = torch.device("cuda:0")
peak = 6550.0
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p)).get("hbm_gbs", peak)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > L2


def timed(fn, reps=20, warm=3, flush_l2=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        if flush_l2:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def report(name, ms, alg_bytes=None, flops=None, extra=None):
    line = {"kernel": name, "ms": ms}
    if alg_bytes is not None:
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        line.update({"algorithmic_bytes": int(alg_bytes), "GB/s": gbs, "frac_of_hbm_peak": gbs / peak})
    if flops is not None:
        line["TFLOP/s"] = flops / (ms * 1e-3) / 1e12
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


B = 64
g = torch.Generator(device=dev).manual_seed(0)
imgs = torch.randint(0, 256, (B, 480, 640, 3), generator=g, device=dev, dtype=torch.uint8)
pre = Preprocessor(376, 224, 0)
x = pre(imgs)
# A1: reads the 376x376 crop of every image once (uint8), writes float32 [B,3,224,224]
report("A1 preprocess (crop 376 -> bicubic 224 -> normalise), batch 64",
       timed(lambda: pre(imgs)), B * (376 * 376 * 3 + 3 * 224 * 224 * 4))

with torch.no_grad():
    vgg = build_vgg16_encoder().to(dev).eval()
    for prec in ("fp32", "tf32", "bf16"):
        with backbone_precision(prec):
            ms = timed(lambda: vgg(x), reps=5, flush_l2=False)
        report(f"A2 VGG16 conv stack to conv5_3 (cuDNN, {prec}), batch 64 [library, not graded]", ms,
               flops=B * 30.7e9, extra={"images_per_s": B / (ms * 1e-3)})
    feat = vgg(x).float().contiguous()                      # [64, 512, 14, 14]
    r18, _ = get_backbone("resnet18")
    r18 = r18.to(dev).eval()
    for prec in ("fp32", "tf32", "bf16"):
        with backbone_precision(prec):
            ms = timed(lambda: r18(x), reps=5, flush_l2=False)
        report(f"A2 ResNet-18 trunk (cuDNN, {prec}), batch 64 [library, not graded]", ms,
               flops=B * 3.6e9, extra={"images_per_s": B / (ms * 1e-3)})
    feat18 = r18(x).float().contiguous()                    # [64, 512, 7, 7]

vlad = NetVLADLayer(device=0)
vlad.load_state(torch.randn(64, 512, generator=torch.Generator().manual_seed(1)) * 0.05, torch.rand(64, 512))
# A3: reads the feature map (401 KB/img) + conv weight + centroids, writes 131 KB/img
report("A3 NetVLAD layer (fused normalise/1x1 conv/softmax/aggregate/intra-norm/L2), batch 64",
       timed(lambda: vlad(feat)), B * (512 * 196 * 4 + 32768 * 4) + 2 * 64 * 512 * 4,
       flops=B * 25.7e6)
v = vlad(feat)
rng = np.random.default_rng(1)
comp = torch.randn(4096, 32768, generator=torch.Generator().manual_seed(2)) / np.sqrt(32768)
pca = PCAProjection(comp.numpy(), 0.01 * rng.standard_normal(32768), rng.uniform(0.5, 1.5, 4096), True, device=0)
# A4: W [4096, 32768] fp32 = 537 MB read once per batch + activations
report("A4 PCA 32768 -> 4096 + whiten + L2, batch 64", timed(lambda: pca(v), reps=10),
       4096 * 32768 * 4 + B * (32768 + 4096) * 4, flops=2.0 * B * 32768 * 4096)
gem = GemHead(512, 512, device=0)
report("A5 CosPlace head (L2Norm/GeM/Linear 512->512/L2), batch 64", timed(lambda: gem(feat18)),
       B * (512 * 49 * 4 + 512 * 4) + 512 * 512 * 4 + 512 * 4)

# A6 regimes
for (n, d, q, k) in ((1000000, 512, 64, 30), (1000000, 512, 64, 1), (1000000, 512, 128, 30),
                     (1000000, 512, 512, 30), (250000, 4096, 64, 30)):
    nn = NearestNeighborsMatching(device=0)
    gg = torch.Generator(device=dev).manual_seed(3)
    for s in range(0, n, 50000):
        r = torch.rand((min(50000, n - s), d), generator=gg, device=dev)
        nn.add_items_device(r / r.norm(dim=1, keepdim=True))
    qs = torch.rand((q, d), generator=gg, device=dev)
    qs = qs / qs.norm(dim=1, keepdim=True)
    total = timed(lambda: nn.search_batch_device(qs, k), reps=10, flush_l2=False)
    coarse = []
    for _ in range(5):
        nn.search_batch_device(qs, k)
        torch.cuda.synchronize()
        coarse.append(nn.last_timing()[0])
    cms = float(np.median(coarse))
    dpad = (d + 63) // 64 * 64
    group = min(q, 256)          # queries served by the timed launch (one SM-pair sweep)
    tiles = (group + 127) // 128
    # useful FLOPs count the REAL queries of the timed launch; the MMA also multiplies the zero
    # rows that pad the last 128-query tile (reported separately, not as throughput)
    report(f"A6 k_nns_coarse_tc last group: pool {n}x{d}, {q} queries ({tiles} tile(s) per pool sweep), k={k}", cms,
           n * dpad * 2 + tiles * 128 * dpad * 2, flops=2.0 * group * n * dpad,
           extra={"issued_TFLOP/s_incl_padding": 2.0 * tiles * 128 * n * dpad / (cms * 1e-3) / 1e12,
                  "search_total_ms": total, "queries_per_s": q / (total * 1e-3),
                  "pool_sweeps": (q + 255) // 256, "info": nn.last_info.tolist()})
    del nn
