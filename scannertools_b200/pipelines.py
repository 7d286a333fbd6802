"""The reference's shipped pipelines over this path, composed from the ops in `ops.py`.

The reference builds these as Scanner op graphs (`Pipeline.make_runner()`); the engine, decode
and storage around them are out of scope here, so the runners below take the decoded
frames directly -- a CUDA uint8 tensor [n, H, W, 3], or a host numpy array / CPU tensor that is
streamed to the device in chunks -- and return what the reference's `parser_fn` would hand back:

  compute_histograms        scannertools/old/histograms.py:6-18    Histogram
  compute_hsv_histograms    old/histograms.py:21-39                ConvertToHSVCPP -> Histogram (one fused pass here)
  compute_flow              old/optical_flow.py:8-27               OpticalFlow (yields batches: "flow fields aren't
                                                                   materialized into memory as they're simply too large")
  compute_flow_histograms   old/histograms.py:49-81                Resize(426x240) -> OpticalFlow -> FlowHistogram
                                                                   (flow never leaves the iteration kernel)
  detect_shots              scannertools/shot_detection.py:11-28   Histogram -> ShotBoundaries

Frame ranges with a one-frame halo are what `sharding.py` hands to each GPU; every runner accepts
any such range, so the multi-GPU form is "run the runner on this rank's range, concatenate".
"""
import numpy as np

from . import ops, shot_detection

FLOW_HIST_WIDTH, FLOW_HIST_HEIGHT = 426, 240    # old/histograms.py:64-68


def _torch():
    import torch
    return torch


def _chunks(frames, size, halo=0):
    """Yields device uint8 tensors [m (+halo), H, W, 3] covering `frames` in order.  Device input
    is sliced (no copy); host input is staged through pinned memory, chunk by chunk."""
    torch = _torch()
    n = int(frames.shape[0])
    if isinstance(frames, torch.Tensor) and frames.is_cuda:
        for a in range(0, n - halo, size):
            yield frames[a:min(a + size + halo, n)]
        return
    host = frames if isinstance(frames, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frames))
    if host.dtype != torch.uint8:
        raise TypeError('frames must be uint8, got %s' % (host.dtype,))
    stage = None
    for a in range(0, n - halo, size):
        part = host[a:min(a + size + halo, n)]
        if stage is None:
            stage = torch.empty((size + halo,) + tuple(part.shape[1:]), dtype=torch.uint8).pin_memory()
        stage[:part.shape[0]].copy_(part)
        yield stage[:part.shape[0]].cuda(non_blocking=False)


def _check_frames(frames):
    if frames.ndim != 4 or frames.shape[3] != 3:
        raise ValueError('frames: expected [n, H, W, 3] uint8, got shape %s' % (tuple(frames.shape),))


def compute_histograms(frames, batch=64, hsv=False):
    """[n, 3, 16] int32 numpy array; element i is what `readers.histograms` parses for frame i."""
    torch = _torch()
    _check_frames(frames)
    out = [ops.histogram(c, hsv='COLOR_RGB2HSV' if hsv else None) for c in _chunks(frames, batch)]
    if not out:
        return np.zeros((0, 3, ops.HIST_BINS), np.int32)
    return torch.cat(out).cpu().numpy()


def compute_hsv_histograms(frames, batch=64):
    return compute_histograms(frames, batch=batch, hsv=True)


def detect_shots(frames, batch=64):
    """Histogram -> ShotBoundaries over one stream: returns the list of boundary frame indices
    (row 0 of the reference op's output).  The integer scores are computed on the device, chunk by
    chunk, each chunk seeded with the previous chunk's last histogram."""
    torch = _torch()
    _check_frames(frames)
    scores, prev = [], None
    for c in _chunks(frames, batch):
        h = ops.histogram(c)
        scores.append(ops.shot_scores(h, prev_hist=prev))
        prev = h[-1:].clone()
    if not scores:
        return []
    S = torch.cat(scores).cpu().numpy()
    return shot_detection.shot_boundaries(None, scores=S)[0]


def compute_flow(frames, batch=16, **farneback_args):
    """Generator of (first_pair_index, flow) with flow a CUDA float32 tensor [m, H, W, 2]; pair i
    maps frame i -> frame i+1.  The tensor is reused by the next iteration."""
    _check_frames(frames)
    n, H, W = int(frames.shape[0]), int(frames.shape[1]), int(frames.shape[2])
    if n < 2:
        return
    of = ops.OpticalFlow(W, H, max_batch=batch, **farneback_args)
    try:
        a = 0
        for c in _chunks(frames, batch, halo=1):
            yield a, of.execute(c)
            a += c.shape[0] - 1
    finally:
        of.close()


def compute_flow_histograms(frames, width=FLOW_HIST_WIDTH, height=FLOW_HIST_HEIGHT, batch=64, **farneback_args):
    """[n-1, 2, 64] int32 numpy array (magnitude bins then angle bins, the layout
    `flow_hist_reader` splits): Resize -> OpticalFlow -> FlowHistogram, with the resized frames
    and the flow fields staying on the device and the flow binned inside the last iteration."""
    torch = _torch()
    _check_frames(frames)
    n, H, W = int(frames.shape[0]), int(frames.shape[1]), int(frames.shape[2])
    if n < 2:
        return np.zeros((0, 2, ops.FLOW_HIST_BINS), np.int32)
    resize = (width, height) != (W, H) and width > 0 and height > 0
    tw, th = (width, height) if resize else (W, H)
    of = ops.OpticalFlow(tw, th, max_batch=batch, **farneback_args)
    out = []
    try:
        for c in _chunks(frames, batch, halo=1):
            small = ops.resize(c, width=tw, height=th) if resize else c
            out.append(of.execute_with_histogram(small, want_flow=False)[1])
    finally:
        of.close()
    return torch.cat(out).cpu().numpy()
