"""seld_b200 -- B200-native (sm_100a) implementation of the SELD feature-extraction hot path.

Modules
  feature_extractor   drop-in for the reference's feature_extractor.py (same call surface, numpy outputs)
  transforms          drop-in for the reference's mask / simple_mask (+ fused batch masking)
  pipeline            batched, HBM-resident extract -> statistics -> all-reduce -> normalise
  build               nvcc build of libseld_b200.so (the C ABI in include/seld_b200.h)
"""
__version__ = '0.1.0'
