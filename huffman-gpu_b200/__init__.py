"""huffman-gpu_b200: B200-native (sm_100a) Huffman variable-length ENCODE hot path, a drop-in for the
reference's histogram -> codebook -> encode/scan/pack sequence (vlnguyen92/Huffman-GPU).

The product is libhuffb200.so (C ABI, include/huffman_b200.h, sources in csrc/); this package is the
host-side mirror of the reference's call sequence for Python callers, tests and bench.py.
Importing it without a built libhuffb200.so raises on first use -- there is no CPU fallback.
"""
from . import capi, stats, workloads                            # noqa: F401
from .capi import HBError                                        # noqa: F401
from .encoder import (Encoder, PinnedBuffer, bits_from_hist, build_codebook,   # noqa: F401
                      encode_variant, lib, shard_offsets, vlc_encode)

__all__ = ["Encoder", "PinnedBuffer", "HBError", "build_codebook", "bits_from_hist",
           "shard_offsets", "encode_variant", "vlc_encode", "workloads", "capi", "lib", "stats"]
