"""medtok_b200 -- B200-native (sm_100a) implementation of MedTok's vector-quantization hot path.

Drop-in modules (same names / signatures / state_dict keys as the reference files):
    medtok_b200.norm_ema_quantizer              <- MedTok/norm_ema_quantizer.py
    medtok_b200.vector_quantization_soft_one_new <- MedTok/vector_quantization_soft_one_new.py
Compute lives in libmedtok_vq.so (C ABI: include/medtok_vq.h, built by `python -m medtok_b200.build`).
"""
from . import _lib  # noqa: F401
from .norm_ema_quantizer import EmbeddingEMA, NormEMAVectorQuantizer, kmeans, l2norm  # noqa: F401
from .vector_quantization_soft_one_new import VectorQuantizer  # noqa: F401

__all__ = ["EmbeddingEMA", "NormEMAVectorQuantizer", "VectorQuantizer", "kmeans", "l2norm"]
