"""B200-native backend for XLab-FFTBarotropic's pseudospectral RK4 step (C ABI: include/xfb.h)."""
from .capi import Backend, XfbError, load  # noqa: F401
