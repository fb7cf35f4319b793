import sys
sys.path.insert(0, ".")
import torch
from cslam_b200.vpr.cosplace import GemHead
dev=torch.device("cuda:0")
feat=torch.randn(64,512,7,7,device=dev)
gem=GemHead(512,512,device=0)
for _ in range(3): gem(feat)
torch.cuda.synchronize()
