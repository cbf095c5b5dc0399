"""B200-native implementation of the TGIS continuous-batching decode hot path.

Layout mirrors the reference's `text_generation_server` package for the modules on the hot path
(models/, utils/) plus `csrc/` (hand-written sm_100a CUDA behind the C ABI in include/b200_tgis.h).
There is no CPU fallback: ops raise if the CUDA library is missing or no GPU is present.
"""
__version__ = "0.1.0"
