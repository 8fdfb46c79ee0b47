"""detsam2_b200 — Blackwell-native SAM 2.1 video-predictor hot path behind the Det-SAM2 API.

Only what the hot path needs lives here: ``csrc/`` (sm_100a kernels + C ABI), the ctypes binding
(``capi``), the kernel-orchestrating engine and the host-side mirror of the reference's
``SAM2VideoPredictor`` / ``VideoProcessor`` interface.
"""
__version__ = "0.1.0"
