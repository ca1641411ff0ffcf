"""sdrangel_b200 — B200-native (sm_100a) implementation of SDRangel's baseband-to-channel DSP hot path.

The product is the CUDA shared library sdrangel_b200/lib/libb200dsp.so (C ABI: include/b200dsp.h).  This package
is the thin Python host side used by tests and bench.py: a ctypes binding (capi) and mirrors of the reference
classes for the path (dsp.Decimators, ...).  There is no CPU fallback: importing the binding without the built
library raises, and every compute call fails loudly without a CUDA device.
"""
from . import capi  # noqa: F401
from .dsp import Decimators, Decimators8, DecimatorsU, DecimatorsFI, DecimatorsFF, DecimatorsIF, DownChannelizerBank, ShardedBank, SpectrumVis, Interpolator, NCO, IQCorrections, Interpolators, UpChannelizer, Demod, SdriqFile, FftFilt  # noqa: F401

__all__ = ["capi", "Decimators", "Decimators8", "DecimatorsU", "DecimatorsFI", "DecimatorsFF", "DecimatorsIF", "DownChannelizerBank", "ShardedBank", "SpectrumVis", "Interpolator", "NCO", "IQCorrections", "Interpolators", "UpChannelizer", "Demod", "SdriqFile", "FftFilt"]
