"""density_b200 — B200-native (sm_100a) Chameleon / Cheetah / Lion encode/decode hot path of g1mv/density.

The product is libdensity_b200.so (hand-written CUDA behind the reference's own C ABI, include/density_b200.h);
this package is the thin host-side mirror of the reference's `Codec` interface plus the torch.distributed plumbing
for the sharded multi-GPU path.
"""
from .codec import (CODECS, Chameleon, Cheetah, DecodeError, EncodeError, Lion, decode_device, encode_device)  # noqa: F401
from ._lib import DensityB200Error, SO_PATH, load  # noqa: F401

__version__ = "0.1.0"
