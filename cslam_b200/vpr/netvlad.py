"""NetVLAD place-recognition front end — same class API as cslam/vpr/netvlad.py
(`NetVLAD(params, node).compute_embedding(keyframe) -> np.ndarray`), with preprocessing,
the VLAD pooling layer and the PCA projection running as hand-written CUDA kernels
(csrc/heads.cu); the VGG16 convolution stack stays a PyTorch/cuDNN module.
"""
import ctypes
import pickle
from os.path import isfile

import numpy as np

from .. import _lib
from ._common import Preprocessor, as_uint8_cuda, backbone_precision, resolve_share_path


def _torch():
    import torch
    return torch


class NetVLADLayer(object):
    """GPU replacement of the reference's NetVLADLayer (netvlad.py:28-130): holds
    `conv.weight` [K, C, 1, 1] and `centroids` [K, C]; forward() is one fused kernel."""

    def __init__(self, num_clusters=64, dim=512, normalize_input=True, vladv2=False, device=0):
        torch = _torch()
        if vladv2:
            raise NotImplementedError("the reference instantiates vladv2=False (netvlad.py:174-176)")
        assert normalize_input
        self.num_clusters, self.dim = num_clusters, dim
        dev = torch.device("cuda", device)
        self.conv_weight = torch.zeros((num_clusters, dim), dtype=torch.float32, device=dev)
        self.centroids = torch.rand((num_clusters, dim), dtype=torch.float32, device=dev)

    def load_state(self, conv_weight, centroids):
        torch = _torch()
        self.conv_weight.copy_(torch.as_tensor(conv_weight).reshape(self.num_clusters, self.dim))
        self.centroids.copy_(torch.as_tensor(centroids).reshape(self.num_clusters, self.dim))

    def forward(self, x):
        """x: CUDA float32 [N, C, H, W] -> [N, K*C] (netvlad.py:94-130)."""
        torch = _torch()
        assert x.is_cuda and x.dtype == torch.float32
        x = x.contiguous()
        N, C = x.shape[:2]
        S = x.shape[2] * x.shape[3]
        out = torch.empty((N, self.num_clusters * C), dtype=torch.float32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.load().cslam_vlad_forward(_lib.ptr(x), N, C, S, _lib.ptr(self.conv_weight),
                                                  _lib.ptr(self.centroids), self.num_clusters,
                                                  _lib.ptr(out), ctypes.c_void_p(stream)))
        return out

    __call__ = forward


class PCAProjection(object):
    """sklearn `pca.transform` + `preprocessing.normalize` on the GPU (netvlad.py:234-237)."""

    def __init__(self, components, mean, explained_variance=None, whiten=False, device=0):
        torch = _torch()
        dev = torch.device("cuda", device)
        comp = np.asarray(components, dtype=np.float32)
        mean = np.asarray(mean, dtype=np.float32)
        self.dout, self.din = comp.shape
        self.W = torch.from_numpy(comp).to(dev).contiguous()
        # sklearn: X_transformed = X @ components_.T - mean_ @ components_.T
        self.bias = (torch.from_numpy(mean.astype(np.float64)).to(dev) @ self.W.double().T).float().contiguous()
        self.scale = None
        if whiten:
            ev = np.asarray(explained_variance, dtype=np.float64)
            self.scale = torch.from_numpy((1.0 / np.sqrt(ev)).astype(np.float32)).to(dev).contiguous()
        self._work = None

    @classmethod
    def from_sklearn(cls, pca, device=0):
        return cls(pca.components_, pca.mean_, getattr(pca, "explained_variance_", None),
                   bool(getattr(pca, "whiten", False)), device)

    def __call__(self, x):
        torch = _torch()
        x = x.contiguous()
        B = x.shape[0]
        out = torch.empty((B, self.dout), dtype=torch.float32, device=x.device)
        need = int(_lib.load().cslam_pca_workspace_floats(min(B, 64), self.dout))
        if self._work is None or self._work.numel() < need:
            self._work = torch.empty(need, dtype=torch.float32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.load().cslam_pca_project_l2(_lib.ptr(x), B, self.din, _lib.ptr(self.W),
                                                    _lib.ptr(self.bias), _lib.ptr(self.scale), self.dout,
                                                    _lib.ptr(out), _lib.ptr(self._work),
                                                    ctypes.c_void_p(stream)))
        return out


def build_vgg16_encoder():
    """VGG16 conv stack up to conv5_3, last ReLU and max-pool removed (netvlad.py:162-171)."""
    import torch.nn as nn
    import torchvision.models as models
    layers = list(models.vgg16(weights=None).features.children())[:-2]
    return nn.Sequential(*layers)


class NetVLAD(object):
    """NetVLAD matcher"""

    def __init__(self, params, node=None, state_dict=None, pca=None, device=None):
        """
        Args:
            params (dict): the reference's parameter dict.  `frontend.nn_checkpoint`:
                'disable' (random descriptors, reference test mode), a checkpoint path, or —
                when `state_dict`/`pca` are passed explicitly — ignored.
            node: ROS node (only used to read 'frontend.netvlad.pca_checkpoint' like the reference)
            state_dict: optional {'encoder.*', 'pool.conv.weight', 'pool.centroids'} tensors
            pca: optional sklearn PCA object or PCAProjection
        """
        self.params = params
        self.node = node
        self.enable = self.params['frontend.nn_checkpoint'].lower() != 'disable'
        if not self.enable:
            return
        torch = _torch()
        _lib.require_device()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.precision = self.params.get('frontend.backbone_precision', 'fp32')

        self.encoder = build_vgg16_encoder().to(self.device).eval()
        self.pool = NetVLADLayer(num_clusters=64, dim=512, vladv2=False, device=self.device_index)
        if state_dict is None:
            path = resolve_share_path(self.params['frontend.nn_checkpoint'])
            if not isfile(path):
                raise FileNotFoundError(f"NetVLAD checkpoint not found: {path}")
            checkpoint = torch.load(path, map_location="cpu")
            state_dict = checkpoint['state_dict'] if 'state_dict' in checkpoint else checkpoint
        self.load_state_dict(state_dict)

        self.transform = Preprocessor(self.params["frontend.image_crop_size"], 224, self.device_index)
        if pca is None:
            name = self.params.get('frontend.netvlad.pca_checkpoint')
            if name is None and node is not None:
                name = node.get_parameter('frontend.netvlad.pca_checkpoint').value
            with open(resolve_share_path(name), 'rb') as f:
                pca = pickle.load(f)
        self.pca = pca if isinstance(pca, PCAProjection) else PCAProjection.from_sklearn(pca, self.device_index)

    def load_state_dict(self, sd):
        torch = _torch()
        sd = {k.replace("module.", ""): v for k, v in sd.items()}  # DataParallel prefixes
        enc = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
        self.encoder.load_state_dict(enc)
        self.pool.load_state(sd["pool.conv.weight"], sd["pool.centroids"])

    def compute_embeddings_device(self, keyframes):
        """uint8 [B, H, W, 3] (numpy or tensor) -> CUDA float32 [B, D] descriptors."""
        torch = _torch()
        with torch.no_grad():
            imgs = as_uint8_cuda(keyframes, self.device)
            x = self.transform(imgs)
            with backbone_precision(self.precision):
                enc = self.encoder(x)
            vlad = self.pool(enc.float())
            return self.pca(vlad)

    def compute_embeddings(self, keyframes):
        return self.compute_embeddings_device(keyframes).cpu().numpy()

    def compute_embedding(self, keyframe):
        """Global image descriptor of one keyframe (netvlad.py:212-245)."""
        if self.enable:
            return self.compute_embeddings(keyframe)[0]
        # Random descriptor if disabled (reference test mode, netvlad.py:242-245)
        return np.random.rand(128)
