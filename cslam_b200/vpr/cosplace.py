"""CosPlace place-recognition front end — same class API as cslam/vpr/cosplace.py
(`CosPlace(params, node).compute_embedding(keyframe)`); preprocessing and the
L2Norm -> GeM -> Linear -> L2Norm aggregation head are CUDA kernels (csrc/heads.cu), the
ResNet/VGG trunk is a PyTorch/cuDNN module.
"""
import ctypes
from os.path import isfile

import numpy as np

from .. import _lib
from ._common import Preprocessor, as_uint8_cuda, backbone_precision, resolve_share_path

# cslam/vpr/cosplace_utils/network.py:10-16
CHANNELS_NUM_IN_LAST_CONV = {"resnet18": 512, "resnet50": 2048, "resnet101": 2048,
                             "resnet152": 2048, "vgg16": 512}


def _torch():
    import torch
    return torch


def get_backbone(backbone_name):
    """Trunk without average pooling / FC (network.py:38-68), randomly initialised."""
    import torch.nn as nn
    import torchvision
    if backbone_name.startswith("resnet"):
        net = getattr(torchvision.models, backbone_name)(weights=None)
        layers = list(net.children())[:-2]
    elif backbone_name == "vgg16":
        layers = list(torchvision.models.vgg16(weights=None).features.children())[:-2]
    else:
        raise ValueError(f"unknown backbone {backbone_name}")
    return nn.Sequential(*layers), CHANNELS_NUM_IN_LAST_CONV[backbone_name]


class GemHead(object):
    """aggregation = [L2Norm, GeM(p, eps), Flatten, Linear, L2Norm] (network.py:23-29)."""

    def __init__(self, features_dim, fc_output_dim, p=3.0, eps=1e-6, device=0):
        torch = _torch()
        dev = torch.device("cuda", device)
        self.p, self.eps = float(p), float(eps)
        self.C, self.D = features_dim, fc_output_dim
        lin = torch.nn.Linear(features_dim, fc_output_dim)
        self.weight = lin.weight.detach().to(dev).contiguous()
        self.bias = lin.bias.detach().to(dev).contiguous()

    def load_state(self, p, weight, bias):
        torch = _torch()
        self.p = float(torch.as_tensor(p).reshape(-1)[0])
        self.weight.copy_(torch.as_tensor(weight))
        self.bias.copy_(torch.as_tensor(bias))

    def __call__(self, x):
        torch = _torch()
        x = x.contiguous()
        B, C = x.shape[:2]
        S = x.shape[2] * x.shape[3]
        out = torch.empty((B, self.D), dtype=torch.float32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.load().cslam_gem_head_forward(_lib.ptr(x), B, C, S, self.p, self.eps,
                                                      _lib.ptr(self.weight), _lib.ptr(self.bias), self.D,
                                                      _lib.ptr(out), ctypes.c_void_p(stream)))
        return out


class CosPlace(object):
    """CosPlace matcher"""

    def __init__(self, params, node=None, state_dict=None, device=None):
        self.params = params
        self.node = node
        self.enable = self.params['frontend.nn_checkpoint'].lower() != 'disable'
        self.descriptor_dim = self.params.get('frontend.cosplace.descriptor_dim', 64)
        if not self.enable:
            return
        torch = _torch()
        _lib.require_device()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.precision = self.params.get('frontend.backbone_precision', 'fp32')
        backbone, features_dim = get_backbone(self.params.get('frontend.cosplace.backbone', 'resnet18'))
        self.backbone = backbone.to(self.device).eval()
        self.aggregation = GemHead(features_dim, self.descriptor_dim, device=self.device_index)
        if state_dict is None:
            path = resolve_share_path(self.params['frontend.nn_checkpoint'])
            if not isfile(path):
                # reference: logs an error and exit()s (cosplace.py:67-70)
                raise FileNotFoundError(f"CosPlace checkpoint not found: {path}")
            state_dict = torch.load(path, map_location="cpu")
        self.load_state_dict(state_dict)
        self.transform = Preprocessor(self.params["frontend.image_crop_size"], 224, self.device_index)

    def load_state_dict(self, sd):
        bb = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
        self.backbone.load_state_dict(bb)
        self.aggregation.load_state(sd["aggregation.1.p"], sd["aggregation.3.weight"],
                                    sd["aggregation.3.bias"])

    def compute_embeddings_device(self, keyframes):
        torch = _torch()
        with torch.no_grad():
            imgs = as_uint8_cuda(keyframes, self.device)
            x = self.transform(imgs)
            with backbone_precision(self.precision):
                feat = self.backbone(x)
            return self.aggregation(feat.float())

    def compute_embeddings(self, keyframes):
        return self.compute_embeddings_device(keyframes).cpu().numpy()

    def compute_embedding(self, keyframe):
        """Global image descriptor of one keyframe (cosplace.py:81-105)."""
        if self.enable:
            return self.compute_embeddings(keyframe)[0]
        return np.random.rand(self.descriptor_dim)
