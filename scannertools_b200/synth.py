"""Seeded synthetic clips for parity tests and benchmarks.

Pure numpy (+ cv2 only for resizing/blurring the *inputs*; never for the ops under test).
Nothing here is on the measured path: frames are generated before any timed region.

* textured_clip  -- band-limited noise with global translation and independently moving
                    blobs: well-conditioned content for Farneback parity.
* cut_clip       -- shots of a fixed up-scaled random image x gain + per-frame noise with
                    hard cuts at seeded positions: mirrors the shape of the reference's
                    test_shot_detection (scannertools/tests/test_all.py:222-233).
* noise_clip     -- i.i.d. uniform bytes (histogram stress).
* const_clip     -- constant frames (worst-case atomic contention).
"""
import numpy as np


def _smooth_noise(rng, h, w, sigma):
    import cv2
    img = rng.random((h, w), dtype=np.float32)
    img = cv2.GaussianBlur(img, (0, 0), sigma)
    lo, hi = float(img.min()), float(img.max())
    return (img - lo) / (hi - lo)


def texture_rgb(seed, h, w, sigma=2.5):
    """3-channel band-limited noise, float32 in [0,1], independent seeds per channel."""
    chans = [_smooth_noise(np.random.default_rng(seed * 101 + c), h, w, sigma) for c in range(3)]
    return np.stack(chans, axis=-1)


def textured_clip(seed, n_frames, h, w, sigma=2.5, max_shift=4.0, n_blobs=3):
    """n_frames x h x w x 3 uint8.  Background translates sub-pixel (bilinear warp) by a
    seeded velocity; n_blobs textured discs move independently on top."""
    import cv2
    rng = np.random.default_rng(seed)
    pad = int(np.ceil(max_shift * n_frames)) + 8
    pad = min(pad, 64 + 8)
    big = texture_rgb(seed, h + 2 * pad, w + 2 * pad, sigma)
    vel = rng.uniform(-max_shift, max_shift, size=2).astype(np.float64)
    blobs = []
    for b in range(n_blobs):
        r = int(rng.integers(min(h, w) // 12 + 2, min(h, w) // 6 + 3))
        tex = texture_rgb(seed * 7 + 13 + b, 2 * r + 1, 2 * r + 1, max(1.0, sigma * 0.6))
        yy, xx = np.mgrid[-r:r + 1, -r:r + 1]
        d = np.sqrt((xx * xx + yy * yy).astype(np.float32)) / r
        alpha = np.clip((1.0 - d) * 4.0, 0.0, 1.0).astype(np.float32)[..., None]
        pos = np.array([rng.uniform(r, w - r), rng.uniform(r, h - r)])
        bv = rng.uniform(-max_shift, max_shift, size=2)
        blobs.append((r, tex, alpha, pos, bv))
    out = np.empty((n_frames, h, w, 3), np.uint8)
    for t in range(n_frames):
        # wrap the accumulated shift into the padding so long clips stay in range
        sx = ((vel[0] * t + (pad - 8)) % (2 * (pad - 8))) - (pad - 8)
        sy = ((vel[1] * t + (pad - 8)) % (2 * (pad - 8))) - (pad - 8)
        Mw = np.array([[1, 0, pad - sx], [0, 1, pad - sy]], np.float64)
        frame = cv2.warpAffine(big, Mw, (w, h), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP)
        for (r, tex, alpha, pos, bv) in blobs:
            p = pos + bv * t
            cx = int(round(p[0])) % w
            cy = int(round(p[1])) % h
            x0, x1 = max(cx - r, 0), min(cx + r + 1, w)
            y0, y1 = max(cy - r, 0), min(cy + r + 1, h)
            if x1 <= x0 or y1 <= y0:
                continue
            ts = tex[y0 - (cy - r):y1 - (cy - r), x0 - (cx - r):x1 - (cx - r)]
            al = alpha[y0 - (cy - r):y1 - (cy - r), x0 - (cx - r):x1 - (cx - r)]
            frame[y0:y1, x0:x1] = frame[y0:y1, x0:x1] * (1 - al) + ts * al
        out[t] = np.clip(frame * 255.0 + 0.5, 0, 255).astype(np.uint8)
    return out


def cut_positions(seed, n_frames, n_cuts):
    rng = np.random.default_rng(seed + 9001)
    lo = max(n_frames // (4 * (n_cuts + 1)), 2)
    while True:
        cuts = np.sort(rng.choice(np.arange(lo, n_frames - lo), size=n_cuts, replace=False))
        if n_cuts < 2 or np.min(np.diff(cuts)) >= lo:
            return [int(c) for c in cuts]


def cut_clip(seed, n_frames, h, w, n_cuts=7, noise=3):
    """Returns (frames uint8 [n,h,w,3], cuts).  Frame `c` for c in cuts is the first frame of
    a new shot."""
    import cv2
    rng = np.random.default_rng(seed)
    cuts = cut_positions(seed, n_frames, n_cuts)
    bounds = [0] + cuts + [n_frames]
    out = np.empty((n_frames, h, w, 3), np.uint8)
    for s in range(len(bounds) - 1):
        small = rng.integers(0, 256, size=(9, 16, 3), dtype=np.uint8)
        base = cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC).astype(np.float32)
        base *= float(rng.uniform(0.4, 1.0))
        for t in range(bounds[s], bounds[s + 1]):
            nz = rng.integers(-noise, noise + 1, size=(h, w, 3)).astype(np.float32)
            out[t] = np.clip(base + nz, 0, 255).astype(np.uint8)
    return out, cuts


def noise_clip(seed, n_frames, h, w):
    return np.random.default_rng(seed).integers(0, 256, size=(n_frames, h, w, 3), dtype=np.uint8)


def const_clip(value, n_frames, h, w):
    return np.full((n_frames, h, w, 3), value, np.uint8)


def textured_flow_field(seed, h, w, scale=6.0):
    """A float32 h x w x 2 field with magnitudes spread over several bins, including exact
    zeros, negative-zero-ish tiny components and values beyond the 64 px range, to stress
    the FlowHistogram edge cases."""
    rng = np.random.default_rng(seed)
    f = rng.normal(0.0, scale, size=(h, w, 2)).astype(np.float32)
    f[::17, ::13] = 0.0
    f[1::29, 2::31, 1] = -1e-8
    f[1::29, 2::31, 0] = 1.0
    f[3::37, 5::41] *= 20.0
    f[5::43, ::47, 0] = 0.0
    f[::53, 7::59, 1] = 0.0
    return f
