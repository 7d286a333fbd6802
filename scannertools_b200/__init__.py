"""scannertools_b200 -- B200-native implementation of scannertools' per-frame analysis hot path
(Histogram, ShotBoundaries scoring, OpticalFlow/Farneback, FlowHistogram, FrameDifference).

See DESIGN.md.  The compute path is hand-written CUDA (csrc/) behind the C ABI in
include/stb.h; this package holds the Python mirror of the reference's op wrappers.
"""
__version__ = '0.1.0'
