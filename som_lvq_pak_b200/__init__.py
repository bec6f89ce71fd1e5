"""som_lvq_pak_b200 -- B200-native best-matching-unit engine for SOM_PAK / LVQ_PAK.

The product is the CUDA shared library `libbmu_b200.so` (C ABI: include/bmu.h); `engine`
is the Python mirror of the reference's plugin interface on top of it."""
from . import _lib  # noqa: F401
from .engine import *  # noqa: F401,F403
