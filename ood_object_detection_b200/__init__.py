"""B200-native (sm_100a) dense per-anchor detection path with the effdet python API.

Drop-in modules (same symbols as the reference's effdet package for this path):
``anchors``, ``loss``, ``bench``, ``soft_nms``, ``object_detection``.  All arithmetic runs in
libodk.so (hand-written CUDA, C ABI in include/odk.h); there is no CPU fallback.
"""
__version__ = '0.1.0'
