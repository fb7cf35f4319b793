"""Shared pieces of the two place-recognition front ends: GPU preprocessing handle,
checkpoint path resolution, backbone precision context."""
import contextlib
import ctypes
import os

import numpy as np

from .. import _lib

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def resolve_share_path(rel):
    """The reference joins checkpoint names with the ROS package share directory
    (cslam/vpr/netvlad.py:150-157).  Without ROS the path is used as given."""
    if os.path.isabs(rel) or os.path.exists(rel):
        return rel
    try:
        from ament_index_python.packages import get_package_share_directory
        return os.path.join(get_package_share_directory("cslam"), rel)
    except Exception:
        return rel


class Preprocessor(object):
    """CenterCrop(crop) -> Resize(224, bicubic) -> ToTensor -> Normalize on the GPU
    (reference transform: cslam/vpr/netvlad.py:202-208), Pillow-exact."""

    def __init__(self, crop, out_size=224, device=0):
        self.crop, self.out_size, self.device = int(crop), int(out_size), int(device)
        self._handles = {}

    def _handle(self, h, w):
        key = (h, w)
        if key not in self._handles:
            hd = ctypes.c_void_p()
            _lib.check(_lib.load().cslam_preproc_create(h, w, self.crop, self.out_size, self.device,
                                                        ctypes.byref(hd)))
            self._handles[key] = hd
        return self._handles[key]

    def __call__(self, images):
        """images: uint8 CUDA tensor [B, H, W, 3] -> float32 CUDA tensor [B, 3, out, out]"""
        import torch
        assert images.is_cuda and images.dtype == torch.uint8 and images.dim() == 4 and images.shape[3] == 3
        images = images.contiguous()
        B, H, W, _ = images.shape
        out = torch.empty((B, 3, self.out_size, self.out_size), dtype=torch.float32, device=images.device)
        stream = torch.cuda.current_stream(images.device).cuda_stream
        _lib.check(_lib.load().cslam_preproc_run(self._handle(H, W), _lib.ptr(images), B, _lib.ptr(out),
                                                 ctypes.c_void_p(stream)))
        return out

    def __del__(self):
        try:
            for hd in self._handles.values():
                _lib.load().cslam_preproc_destroy(hd)
        except Exception:
            pass


@contextlib.contextmanager
def backbone_precision(mode):
    """fp32 (bit-for-bit comparable with the reference's torch path), tf32 or bf16 for the
    library (cuDNN/cuBLAS) backbone; the hand-written heads always run fp32."""
    import torch
    old_c, old_m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    try:
        tf32 = mode in ("tf32", "bf16")
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        if mode == "bf16":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                yield
        else:
            yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_c, old_m


def as_uint8_cuda(keyframes, device):
    import torch
    if isinstance(keyframes, np.ndarray):
        if keyframes.ndim == 3:
            keyframes = keyframes[None]
        t = torch.from_numpy(np.ascontiguousarray(keyframes))
        return t.to(device, non_blocking=True)
    if keyframes.dim() == 3:
        keyframes = keyframes[None]
    return keyframes.to(device, non_blocking=True)   # async when the source is pinned
