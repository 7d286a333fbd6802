"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference ops, restated as the exact OpenCV calls their C++ wrappers make
(citations relative to /root/reference/scannertools/).
"""
import numpy as np
import cv2

BINS_RGB = 16     # scannertools_cpp/imgproc/histogram_kernel_cpu.cpp:8
BINS_FLOW = 64    # scannertools/old/cpp_ops/flow_histogram_kernel_cpu.cpp:9
WINDOW_SIZE = 500  # scannertools/shot_detection.py:7

# scannertools_cpp/imgproc/optical_flow_kernel_cpu.cpp:15-16
FARNEBACK_ARGS = dict(numLevels=3, pyrScale=0.5, fastPyramids=False, winSize=15,
                      numIters=3, polyN=5, polySigma=1.2, flags=0)


def histogram(frame):
    """histogram_kernel_cpu.cpp:25-44 -- 3 x calcHist(16 bins, [0,256)) -> int32[3][16]."""
    out = np.empty((3, BINS_RGB), np.int32)
    for j in range(3):
        h = cv2.calcHist([frame], [j], None, [BINS_RGB], [0, 256])
        out[j] = h.reshape(-1).astype(np.int32)   # hist.convertTo(CV_32SC1)
    return out


def gray(frame):
    """optical_flow_kernel_cpu.cpp:38-39 -- COLOR_BGR2GRAY applied to the (RGB) frame."""
    return cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)


_finder = None


def optical_flow(frame0, frame1):
    """optical_flow_kernel_cpu.cpp:27-43 -- flow from stencil[0] to stencil[1], HxWx2 f32."""
    global _finder
    if _finder is None:
        a = FARNEBACK_ARGS
        _finder = cv2.FarnebackOpticalFlow_create(a['numLevels'], a['pyrScale'], a['fastPyramids'],
                                                  a['winSize'], a['numIters'], a['polyN'],
                                                  a['polySigma'], a['flags'])
    return _finder.calc(gray(frame0), gray(frame1), None)


def optical_flow_params(frame0, frame1, num_levels=3, win_size=15, num_iters=3, flags=0, pyr_scale=0.5, poly_n=5,
                        poly_sigma=1.2):
    """The same op with other FarnebackOpticalFlow arguments (flags=256: OPTFLOW_FARNEBACK_GAUSSIAN);
    the reference itself only ever uses FARNEBACK_ARGS."""
    a = FARNEBACK_ARGS
    return cv2.calcOpticalFlowFarneback(gray(frame0), gray(frame1), None, pyr_scale, num_levels, win_size,
                                        num_iters, poly_n, poly_sigma, flags)


def flow_histogram(flow):
    """flow_histogram_kernel_cpu.cpp:27-54 -- int32[2][64]: magnitude [0,64), angle [0,360)."""
    x, y = cv2.split(flow)
    mag, deg = cv2.cartToPolar(x, y, angleInDegrees=True)
    out = np.empty((2, BINS_FLOW), np.int32)
    out[0] = cv2.calcHist([mag], [0], None, [BINS_FLOW], [0, 64.0]).reshape(-1).astype(np.int32)
    out[1] = cv2.calcHist([deg], [0], None, [BINS_FLOW], [0, 360]).reshape(-1).astype(np.int32)
    return out


def frame_difference(prev, cur):
    """Intended semantics of frame_difference_kernel_cpu.cpp:51-61 (dead, uncompilable code):
    out = cur - prev per byte, u8 wrap-around, every x."""
    return ((cur.astype(np.int16) - prev.astype(np.int16)) & 0xFF).astype(np.uint8)


def shot_scores(hists):
    """S_i = 3*diffs[i] = sum_j max_b |h[i-1][j][b] - h[i][j][b]|, S_0 = 0
    (shot_detection.py:14-18; Chebyshev distance, exact integers)."""
    h = np.asarray(hists, dtype=np.int64).reshape(len(hists), 3, -1)
    S = np.zeros(len(h), np.int64)
    if len(h) > 1:
        S[1:] = np.abs(h[1:] - h[:-1]).max(axis=2).sum(axis=1)
    return S.astype(np.int32)


def shot_boundaries(hists):
    """shot_detection.py:12-28, same numpy expressions (scipy's chebyshev replaced by its
    definition max|a-b|; probed identical)."""
    n = len(hists)
    if n == 0:
        return []
    diffs = np.array([
        np.mean([np.max(np.abs(np.asarray(hists[i - 1][j], np.int32) - np.asarray(hists[i][j], np.int32)))
                 for j in range(3)])
        for i in range(1, n)
    ])
    diffs = np.insert(diffs, 0, 0)
    boundaries = []
    for i in range(1, n):
        window = diffs[max(i - WINDOW_SIZE, 0):min(i + WINDOW_SIZE, n)]
        if diffs[i] - np.mean(window) > 2.5 * np.std(window):
            boundaries.append(i)
    return boundaries


def resize(frame, width, height, interpolation='INTER_LINEAR'):
    """Resize op (scannertools_cpp/imgproc/resize_kernel.cpp:31-35,69-71):
    cv::resize(img, out, Size(width, height), 0, 0, <interpolation>), INTER_LINEAR by default."""
    return cv2.resize(frame, (width, height), interpolation=getattr(cv2, interpolation))


def convert_color(frame, conversion):
    """ConvertColor / ConvertToHSVCPP ops: cv::cvtColor(frame, out, <conversion>)
    (scannertools_cpp/imgproc/convert_color_kernel.cpp:270-276, old/cpp_ops/imgproc.cpp:41)."""
    return cv2.cvtColor(frame, getattr(cv2, conversion))
