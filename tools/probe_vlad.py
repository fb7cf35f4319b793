"""NetVLAD layer at BASELINE.json configs[1] shape (batch 64, 512 x 14 x 14): CUDA-event time of the
whole layer (three kernels on the tensor-core path, one on the fused fp32 path: CSLAM_VLAD_TC=0)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cslam_b200.vpr.netvlad import NetVLADLayer

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator().manual_seed(0)
feat = torch.randn(B, 512, 14, 14, generator=g).to(dev)
layer = NetVLADLayer(device=0)
layer.load_state(torch.randn(64, 512, generator=g) * 0.05, torch.rand(64, 512, generator=g))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    layer(feat)
torch.cuda.synchronize()
ms = []
for _ in range(20):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); layer(feat); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
t = float(np.median(ms))
bytes_alg = B * (512 * 196 * 4 + 32768 * 4) + 2 * 64 * 512 * 4
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6551.4)
print(json.dumps({"kernel": "A3 NetVLAD layer", "path": "fused fp32" if os.environ.get("CSLAM_VLAD_TC") == "0" else "fp32 assign + tcgen05 tf32 aggregation + finish",
                  "batch": B, "ms": round(t, 4), "algorithmic_bytes": bytes_alg, "GB/s": round(bytes_alg / t / 1e6, 1),
                  "frac_of_hbm_peak": round(bytes_alg / t / 1e6 / peak, 4),
                  "TFLOP/s": round(B * 2 * 2 * 64 * 512 * 196 / t / 1e9, 2)}))
