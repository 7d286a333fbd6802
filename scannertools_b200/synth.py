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


def textured_clip(seed, n_frames, h, w, sigma=2.5, max_shift=4.0, n_blobs=3, t0=0, total=None):
    """n_frames x h x w x 3 uint8.  Background translates sub-pixel (bilinear warp) by a
    seeded velocity; n_blobs textured discs move independently on top.

    t0 / total: frames [t0, t0 + n_frames) of the clip of `total` frames with this seed -- every
    frame depends only on (seed, total, its own index), so a frame-range shard of a long clip is
    generated without generating the rest (textured_clip(s, n, h, w, t0=a, total=T) ==
    textured_clip(s, T, h, w)[a:a + n])."""
    import cv2
    rng = np.random.default_rng(seed)
    if total is None:
        total = t0 + n_frames
    pad = int(np.ceil(max_shift * total)) + 8
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
    for ti in range(n_frames):
        t = t0 + ti
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
        out[ti] = np.clip(frame * 255.0 + 0.5, 0, 255).astype(np.uint8)
    return out


def warped_clip(seed, n_frames, h, w, sigma=2.5):
    """n_frames x h x w x 3 uint8: one band-limited texture under a smooth global motion -- sub-pixel
    translation plus a small rotation and zoom about a seeded centre -- so the flow field covers a
    continuous range of magnitudes and directions and nothing moves exactly along an axis or by an
    integer vector.  (textured_clip's blobs move by integer vectors, i.e. exactly ON the 0/45/90 degree
    bin edges of the FlowHistogram angle axis: fine for flow parity, degenerate for per-bin counts.)"""
    import cv2
    rng = np.random.default_rng([seed, 77])
    pad = 48
    big = texture_rgb(seed + 4000, h + 2 * pad, w + 2 * pad, sigma)
    vel = rng.uniform(1.1, 2.9, size=2) * rng.choice([-1.0, 1.0], size=2)
    theta = float(rng.uniform(0.6, 1.2)) * 3.0 / max(h, w) * float(rng.choice([-1.0, 1.0]))   # ~ +-3 px at the far corner
    zoom = 1.0 + float(rng.uniform(0.3, 0.9)) * 2.0 / max(h, w)
    cx, cy = w * float(rng.uniform(0.35, 0.65)), h * float(rng.uniform(0.35, 0.65))
    out = np.empty((n_frames, h, w, 3), np.uint8)
    for t in range(n_frames):
        a, s = theta * t, zoom ** t
        c, sn = np.cos(a) * s, np.sin(a) * s
        # destination (x, y) -> source: rotate/zoom about (cx, cy), translate, shift into the padded texture
        Mw = np.array([[c, -sn, cx - c * cx + sn * cy + pad - vel[0] * t],
                       [sn, c, cy - sn * cx - c * cy + pad - vel[1] * t]], np.float64)
        frame = cv2.warpAffine(big, Mw, (w, h), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP)
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


def cut_clip_range(seed, n_total, h, w, t0, t1, cuts, noise=3):
    """Frames [t0, t1) of an n_total-frame clip with hard cuts at the given frame indices (frame c
    is the first frame of a new shot), generated without the rest of the clip: shot s has a base
    image seeded by (seed, s) and every frame its own noise seeded by (seed, t).  Used for
    frame-range-sharded shot detection, where `cuts` may be planted exactly on shard seams."""
    import cv2
    bounds = [0] + sorted(int(c) for c in cuts) + [n_total]
    out = np.empty((max(t1 - t0, 0), h, w, 3), np.uint8)
    bases = {}
    for t in range(t0, t1):
        s = max(i for i in range(len(bounds) - 1) if bounds[i] <= t)
        if s not in bases:
            rs = np.random.default_rng([seed, 1, s])
            small = rs.integers(0, 256, size=(9, 16, 3), dtype=np.uint8)
            base = cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC).astype(np.float32)
            # alternate dark / bright shots so every planted cut is a clear histogram change
            bases[s] = base * float((0.45 if s % 2 == 0 else 0.9) * rs.uniform(0.92, 1.08))
        nz = np.random.default_rng([seed, 2, t]).integers(-noise, noise + 1, size=(h, w, 3), dtype=np.int8)
        out[t - t0] = np.clip(bases[s] + nz, 0, 255).astype(np.uint8)
    return out


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


def edge_flow_field(seed, h, w):
    """A float32 h x w x 2 flow field for FlowHistogram binning tests: generic vectors of several scales, magnitudes
    on and next to the integer bin edges, directions on and next to the 64 angle-bin edges (5.625 degrees apart),
    and 16 tiny / huge / infinite / axis-aligned specials (no NaN: OpenCV bins cvFloor(NaN) garbage)."""
    rng = np.random.default_rng(seed)
    n = h * w
    f = (rng.standard_normal((n, 2)) * rng.choice([0.3, 3.0, 12.0, 40.0], size=(n, 1))).astype(np.float32)
    k = n // 8
    ang = np.deg2rad(rng.integers(0, 64, k) * 5.625 + rng.choice([0.0, 1e-4, -1e-4, 3e-3, -3e-3], k))
    mag = rng.uniform(0.1, 70.0, k)
    f[:k] = np.stack([mag * np.cos(ang), mag * np.sin(ang)], 1).astype(np.float32)
    m2 = rng.integers(0, 66, k).astype(np.float64) + rng.choice([0.0, 1e-6, -1e-6, 1e-4, -1e-4], k)
    a2 = rng.uniform(0, 2 * np.pi, k)
    f[k:2 * k] = np.stack([m2 * np.cos(a2), m2 * np.sin(a2)], 1).astype(np.float32)
    f[2 * k:2 * k + 16] = np.array([[0, 0], [1e-20, 0], [0, -1e-20], [1e-8, 1e-8], [1e20, 1], [-1e20, -1e20], [np.inf, 1], [1, -np.inf],
                                    [2, -2], [3, 4], [-3, 4], [0, 5], [0, -5], [64, 0], [-64, 0], [1e-7, -1e-30]], np.float32)
    return np.ascontiguousarray(f.reshape(h, w, 2))
