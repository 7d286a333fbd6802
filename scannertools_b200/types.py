"""Wire-format readers for the outputs of the ops on this path -- same byte layouts as the
reference's readers (/root/reference/scannertools/scannertools/types.py:23-41 and
scannertools/old/histograms.py:43-46)."""
import numpy as np


def histograms(buf, protobufs=None):
    """192-byte Histogram element -> [r, g, b] int32 arrays of 16 bins (types.py:23-27)."""
    if buf is None:
        return None
    return np.split(np.frombuffer(buf, dtype=np.dtype(np.int32)), 3)


def flow_hist_reader(buf, protobufs=None):
    """512-byte FlowHistogram element -> [magnitude, angle] int32 arrays of 64 bins
    (old/histograms.py:43-46)."""
    if buf is None:
        return None
    return np.split(np.frombuffer(buf, dtype=np.dtype(np.int32)), 2)


def flow(buf, height, width):
    """OpticalFlow frame bytes -> float32 [H, W, 2].  (The reference's reader, types.py:36-41,
    takes the FrameInfo from an undefined `db`; height/width are explicit here.)"""
    if buf is None:
        return None
    return np.frombuffer(buf, dtype=np.dtype(np.float32)).reshape((height, width, 2))


def histogram_bytes(hist_row):
    """int32 [3,16] -> the 192-byte element the Histogram kernel inserts
    (histogram_kernel_cpu.cpp:44)."""
    return np.ascontiguousarray(hist_row, dtype=np.int32).tobytes()
