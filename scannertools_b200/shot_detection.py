"""ShotBoundaries -- host half of the shot-detection path.

Mirrors the reference Python op `shot_boundaries(config, histograms)`
(/root/reference/scannertools/scannertools/shot_detection.py:11-28): same name, arguments
and return layout (row 0 = list of boundary frame indices, rows 1..n-1 = None).

Split of work: the per-frame Chebyshev histogram distances (shot_detection.py:14-18) are exact
integers S[i] = 3*diffs[i], computed on the GPU right after the histograms
(`ops.shot_scores`, stb_shot_scores).  The +-WINDOW_SIZE outlier test (shot_detection.py:22-26)
needs the whole stream and float64 mean/std; it runs here on the host over the n*4-byte score
array after the per-GPU shards are concatenated, evaluated per window with numpy's own
mean/std so boundary decisions are bit-identical with the reference's.
"""
import numpy as np

WINDOW_SIZE = 500           # shot_detection.py:7
BOUNDARY_BATCH = 10000000   # shot_detection.py:8: the op sees the whole stream in one batch
THRESHOLD_SIGMAS = 2.5      # shot_detection.py:25


def boundaries_from_scores(scores, window_size=WINDOW_SIZE, threshold=THRESHOLD_SIGMAS):
    """scores: int array, scores[i] = sum over channels of the Chebyshev distance between the
    histograms of frames i-1 and i (scores[0] = 0).  Returns the list of boundary indices."""
    diffs = np.asarray(scores, dtype=np.float64) / 3.0   # == np.mean of the 3 integer distances
    n = len(diffs)
    found = []
    for i in range(1, n):
        lo = i - window_size if i > window_size else 0
        hi = i + window_size if i + window_size < n else n
        win = diffs[lo:hi]
        if diffs[i] - np.mean(win) > threshold * np.std(win):
            found.append(i)
    return found


def scores_from_histograms(histograms):
    """Host-side integer scores from a sequence of Histogram elements (each: 3 arrays of 16
    int32, as `types.histograms` returns).  Only used when the caller has histograms but no
    device scores (e.g. read back from storage); the GPU path is ops.shot_scores."""
    n = len(histograms)
    h = np.asarray([[np.asarray(c, dtype=np.int64) for c in el] for el in histograms], dtype=np.int64).reshape(n, 3, -1)
    S = np.zeros(n, np.int64)
    if n > 1:
        S[1:] = np.abs(h[1:] - h[:-1]).max(axis=2).sum(axis=1)
    return S


def shot_boundaries(config, histograms=None, scores=None):
    """Drop-in for the reference op body.  Pass either `histograms` (sequence of Histogram
    elements) or the device-computed `scores`."""
    if scores is None:
        scores = scores_from_histograms(histograms)
    n = len(scores)
    if n == 0:
        return []
    return [boundaries_from_scores(scores)] + [None for _ in range(n - 1)]
