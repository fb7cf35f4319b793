"""Golden outputs of the reference's descriptor-extraction modules (imported from
/root/reference, see oracle/make_golden.py) on the seeded inputs of oracle/inputs.py."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def gen_heads():
    import torch
    import sklearn.preprocessing
    from sklearn.decomposition import PCA
    from PIL import Image
    from cslam.vpr.netvlad import NetVLADLayer  # the reference's layer
    from cslam.vpr.cosplace_utils.layers import Flatten, GeM, L2Norm
    from oracle.inputs import gem_case, keyframe_image, pca_case, subsample, vlad_case
    import torchvision.transforms as transforms
    out = {}
    # A1: the reference transform (netvlad.py:202-208) with the example crop size 376
    img = keyframe_image()
    tr = transforms.Compose([transforms.CenterCrop(376), transforms.Resize(224, interpolation=3),
                             transforms.ToTensor(),
                             transforms.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])
    pre = tr(Image.fromarray(img)).numpy()
    out["pre_sub"] = subsample(pre, 13)
    out["pre_sum"] = np.array([pre.astype(np.float64).sum(), (pre.astype(np.float64) ** 2).sum()])
    # A3: NetVLADLayer.forward
    x, conv_w, cent = vlad_case()
    layer = NetVLADLayer(num_clusters=64, dim=512, vladv2=False)
    with torch.no_grad():
        layer.conv.weight.copy_(torch.from_numpy(conv_w).view(64, 512, 1, 1))
        layer.centroids.copy_(torch.from_numpy(cent))
        v = layer(torch.from_numpy(x)).numpy()
    out["vlad_sub"] = subsample(v, 37)
    out["vlad_sum"] = np.array([v.astype(np.float64).sum(), (v.astype(np.float64) ** 2).sum()])
    # A4: sklearn PCA.transform + normalize, exactly the calls of netvlad.py:234-236
    px, comp, mean, ev, whiten = pca_case()
    pca = PCA(n_components=comp.shape[0], whiten=whiten)
    pca.components_, pca.mean_, pca.explained_variance_ = comp, mean, ev
    pca.n_components_ = comp.shape[0]
    pca.n_features_in_ = comp.shape[1]
    red = pca.transform(px)
    out["pca_out"] = sklearn.preprocessing.normalize(red)
    # A5: aggregation of GeoLocalizationNet (network.py:23-29)
    gx, p, eps, w, b = gem_case()
    lin = torch.nn.Linear(w.shape[1], w.shape[0])
    with torch.no_grad():
        lin.weight.copy_(torch.from_numpy(w))
        lin.bias.copy_(torch.from_numpy(b))
        agg = torch.nn.Sequential(L2Norm(), GeM(), Flatten(), lin, L2Norm())
        out["gem_out"] = agg(torch.from_numpy(gx)).numpy()
    np.savez_compressed(os.path.join(GOLD, "heads.npz"), **out)


GENERATORS = {"heads": gen_heads}
