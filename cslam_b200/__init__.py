"""cslam_b200 — B200 (sm_100a) native loop-closure front end with the class API
of lajoiepy/cslam's hot path (descriptor -> cosine NNS -> MAC sparsification).
Arithmetic runs in hand-written CUDA behind the C ABI of include/cslam_b200.h."""
__version__ = "0.1.0"
