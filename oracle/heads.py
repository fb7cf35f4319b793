"""Oracle for the descriptor-extraction stages (TEST INFRASTRUCTURE ONLY).

torch-CPU / numpy restatements of
  * the preprocessing transform                      cslam/vpr/netvlad.py:202-208, :223-226
  * NetVLADLayer.forward                             cslam/vpr/netvlad.py:94-130
  * pca.transform + sklearn normalize                cslam/vpr/netvlad.py:234-237
    (sklearn.decomposition._base._BasePCA.transform: X @ components_.T - mean_ @ components_.T,
     divided by sqrt(explained_variance_) when whiten; preprocessing.normalize: row / ||row||_2)
  * GeoLocalizationNet.aggregation                   cslam/vpr/cosplace_utils/network.py:23-29,
                                                     cslam/vpr/cosplace_utils/layers.py:8-36
Pinned against outputs of the reference's own modules in tests/golden/heads.npz
(oracle/make_golden_heads.py).
"""
import numpy as np

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def preprocess(keyframe, crop):
    """The reference's transform on one HxWx3 uint8 image -> float32 [3, 224, 224]."""
    import torchvision.transforms as transforms
    from PIL import Image
    t = transforms.Compose([
        transforms.CenterCrop(crop),
        transforms.Resize(224, interpolation=3),
        transforms.ToTensor(),
        transforms.Normalize(IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD),
    ])
    return t(Image.fromarray(keyframe)).numpy()


def netvlad_layer(x, conv_w, centroids):
    """NetVLADLayer.forward (vladv2=False, normalize_input=True), per-cluster residual sums
    like the reference loop (netvlad.py:119-124)."""
    import torch
    import torch.nn.functional as F
    x = torch.as_tensor(x, dtype=torch.float32)
    conv_w = torch.as_tensor(conv_w, dtype=torch.float32)
    centroids = torch.as_tensor(centroids, dtype=torch.float32)
    N, C = x.shape[:2]
    K = conv_w.shape[0]
    x = F.normalize(x, p=2, dim=1)
    soft = F.conv2d(x, conv_w.view(K, C, 1, 1)).view(N, K, -1)
    soft = F.softmax(soft, dim=1)
    xf = x.view(N, C, -1)
    vlad = torch.zeros([N, K, C], dtype=x.dtype)
    for k in range(K):
        residual = xf - centroids[k].view(1, C, 1)
        residual = residual * soft[:, k:k + 1, :]
        vlad[:, k, :] = residual.sum(dim=-1)
    vlad = F.normalize(vlad, p=2, dim=2)
    vlad = vlad.view(N, -1)
    vlad = F.normalize(vlad, p=2, dim=1)
    return vlad.numpy()


def pca_project_normalize(x, components, mean, explained_variance=None, whiten=False):
    """sklearn PCA.transform followed by sklearn.preprocessing.normalize (l2)."""
    x = np.asarray(x)
    xt = x @ components.T - (mean.reshape(1, -1) @ components.T)
    if whiten:
        xt = xt / np.sqrt(explained_variance)
    xt = xt.astype(x.dtype, copy=False)
    norms = np.sqrt((xt.astype(np.float64) ** 2).sum(axis=1))
    norms[norms == 0] = 1.0
    return (xt / norms[:, None]).astype(x.dtype)


def gem_head(x, p, eps, fc_w, fc_b):
    """L2Norm -> GeM -> Flatten -> Linear -> L2Norm."""
    import torch
    import torch.nn.functional as F
    x = torch.as_tensor(x, dtype=torch.float32)
    x = F.normalize(x, p=2, dim=1)
    pt = torch.ones(1) * p
    x = F.avg_pool2d(x.clamp(min=eps).pow(pt), (x.size(-2), x.size(-1))).pow(1. / pt)
    x = x[:, :, 0, 0]
    x = F.linear(x, torch.as_tensor(fc_w), torch.as_tensor(fc_b))
    x = F.normalize(x, p=2, dim=1)
    return x.numpy()


# ---- whole-pipeline oracles (torch CPU backbone, seeded weights) --------------------------
def build_netvlad_modules(seed=0):
    """VGG16[:-2] encoder + pool parameters, seeded; returns (encoder, state_dict) where the
    state dict uses the reference checkpoint keys (encoder.*, pool.conv.weight, pool.centroids)."""
    import torch
    import torch.nn as nn
    import torchvision.models as models
    torch.manual_seed(seed)
    encoder = nn.Sequential(*list(models.vgg16(weights=None).features.children())[:-2]).eval()
    conv_w = torch.randn(64, 512, 1, 1) * 0.2
    centroids = torch.rand(64, 512)
    sd = {"encoder." + k: v for k, v in encoder.state_dict().items()}
    sd["pool.conv.weight"] = conv_w
    sd["pool.centroids"] = centroids
    return encoder, sd


def build_cosplace_modules(seed=0, backbone="resnet18", dim=512):
    import torch
    import torch.nn as nn
    import torchvision
    torch.manual_seed(seed)
    net = getattr(torchvision.models, backbone)(weights=None)
    trunk = nn.Sequential(*list(net.children())[:-2]).eval()
    feat = 512 if backbone in ("resnet18", "vgg16") else 2048
    lin = nn.Linear(feat, dim)
    sd = {"backbone." + k: v for k, v in trunk.state_dict().items()}
    sd["aggregation.1.p"] = torch.ones(1) * 3
    sd["aggregation.3.weight"] = lin.weight.detach()
    sd["aggregation.3.bias"] = lin.bias.detach()
    return trunk, sd


def synthetic_pca(din=32768, dout=4096, seed=1):
    rng = np.random.default_rng(seed)
    comp = (rng.standard_normal((dout, din), dtype=np.float32) / np.float32(np.sqrt(din)))
    mean = (0.01 * rng.standard_normal(din)).astype(np.float32)
    ev = rng.uniform(0.5, 1.5, dout)
    return comp, mean, ev


def netvlad_embedding(keyframe, crop, encoder, sd, pca):
    """Body of NetVLAD.compute_embedding (netvlad.py:222-239) on torch CPU."""
    import torch
    with torch.no_grad():
        x = torch.from_numpy(preprocess(keyframe, crop)).unsqueeze(0)
        enc = encoder(x)
        vlad = netvlad_layer(enc, sd["pool.conv.weight"].reshape(64, 512), sd["pool.centroids"])
    comp, mean, ev, whiten = pca
    return pca_project_normalize(vlad, comp, mean, ev, whiten)[0]


def cosplace_embedding(keyframe, crop, trunk, sd):
    """Body of CosPlace.compute_embedding (cosplace.py:91-100) on torch CPU."""
    import torch
    with torch.no_grad():
        x = torch.from_numpy(preprocess(keyframe, crop)).unsqueeze(0)
        feat = trunk(x)
    return gem_head(feat, float(sd["aggregation.1.p"][0]), 1e-6, sd["aggregation.3.weight"],
                    sd["aggregation.3.bias"])[0]
