"""B200-native backend for XLab-FFTBarotropic's pseudospectral RK4 step (C ABI: include/xfb.h)."""
from .capi import Backend, LoopbackTeam, SlabBackend, XfbError, load, nccl_unique_id, slab_partition  # noqa: F401
