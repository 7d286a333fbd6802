"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): ctypes loader for oracle/restate.c."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'liboracle_restate.so')
_lib = None


class _Dump(C.Structure):
    _fields_ = [('level', C.c_int), ('I0', C.c_void_p), ('I1', C.c_void_p), ('R0', C.c_void_p),
                ('R1', C.c_void_p), ('M0', C.c_void_p), ('flow_out', C.c_void_p)]


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE])


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, 'restate.c')
        if not os.path.isfile(_SO) or (os.path.isfile(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def gray(rgb):
    rgb = np.ascontiguousarray(rgb, np.uint8)
    out = np.empty(rgb.shape[:-1], np.uint8)
    lib().orc_gray(_p(rgb), C.c_int(out.size), _p(out))
    return out


def histogram(frame):
    frame = np.ascontiguousarray(frame, np.uint8)
    out = np.empty((3, 16), np.int32)
    lib().orc_hist_rgb16(_p(frame), C.c_int(frame.shape[1]), C.c_int(frame.shape[0]), _p(out))
    return out


def shot_scores(hists):
    h = np.ascontiguousarray(np.asarray(hists, np.int32).reshape(len(hists), 48))
    out = np.zeros(len(h), np.int32)
    lib().orc_shot_scores(_p(h), C.c_int(len(h)), _p(out))
    return out


def frame_difference(prev, cur):
    prev = np.ascontiguousarray(prev, np.uint8)
    cur = np.ascontiguousarray(cur, np.uint8)
    out = np.empty_like(cur)
    lib().orc_frame_diff(_p(prev), _p(cur), _p(out), C.c_size_t(cur.size))
    return out


def polar(flow):
    flow = np.ascontiguousarray(flow, np.float32)
    n = flow.shape[0] * flow.shape[1]
    mag = np.empty(flow.shape[:2], np.float32)
    deg = np.empty(flow.shape[:2], np.float32)
    lib().orc_polar(_p(flow), C.c_int(n), _p(mag), _p(deg))
    return mag, deg


def flow_histogram(flow):
    flow = np.ascontiguousarray(flow, np.float32)
    out = np.empty((2, 64), np.int32)
    lib().orc_flow_hist(_p(flow), C.c_int(flow.shape[1]), C.c_int(flow.shape[0]), _p(out))
    return out


def pyramid_info(W, H, num_levels=3, pyr_scale=0.5):
    ws = (C.c_int * 8)()
    hs = (C.c_int * 8)()
    l = lib().orc_pyramid_info(C.c_int(W), C.c_int(H), C.c_int(num_levels), C.c_double(pyr_scale), ws, hs)
    return [(ws[k], hs[k]) for k in range(l + 1)]


def poly_consts(n=5, sigma=1.2):
    g = np.zeros(n + 1, np.float32)
    xg = np.zeros(n + 1, np.float32)
    xxg = np.zeros(n + 1, np.float32)
    ig = np.zeros(4, np.float64)
    lib().orc_poly_consts(C.c_int(n), C.c_double(sigma), _p(g), _p(xg), _p(xxg), _p(ig))
    return g, xg, xxg, ig


def farneback(gray0, gray1, dump_level=None, winsize=15, iters=3, levels=3, flags=0, pyr_scale=0.5, poly_n=5, poly_sigma=1.2):
    """Returns flow (HxWx2 f32).  With dump_level=k also returns a dict of level-k
    intermediates: I0, I1 (h x w), R0, R1, M0 (h x w x 5), flow (h x w x 2)."""
    g0 = np.ascontiguousarray(gray0, np.uint8)
    g1 = np.ascontiguousarray(gray1, np.uint8)
    H, W = g0.shape
    out = np.empty((H, W, 2), np.float32)
    if dump_level is None:
        lib().orc_farneback_flags(_p(g0), _p(g1), C.c_int(W), C.c_int(H), _p(out), C.c_int(levels), C.c_double(pyr_scale),
                                  C.c_int(winsize), C.c_int(iters), C.c_int(poly_n), C.c_double(poly_sigma), C.c_int(flags), None)
        return out
    w, h = pyramid_info(W, H)[dump_level]
    d = dict(I0=np.empty((h, w), np.float32), I1=np.empty((h, w), np.float32),
             R0=np.empty((h, w, 5), np.float32), R1=np.empty((h, w, 5), np.float32),
             M0=np.empty((h, w, 5), np.float32), flow=np.empty((h, w, 2), np.float32))
    dump = _Dump(dump_level, _p(d['I0']), _p(d['I1']), _p(d['R0']), _p(d['R1']), _p(d['M0']), _p(d['flow']))
    lib().orc_farneback_ex(_p(g0), _p(g1), C.c_int(W), C.c_int(H), _p(out), C.c_int(3), C.c_double(0.5),
                           C.c_int(winsize), C.c_int(iters), C.c_int(5), C.c_double(1.2), C.byref(dump))
    return out, d


def optical_flow(frame0, frame1):
    f0 = np.ascontiguousarray(frame0, np.uint8)
    f1 = np.ascontiguousarray(frame1, np.uint8)
    H, W = f0.shape[:2]
    out = np.empty((H, W, 2), np.float32)
    lib().orc_optical_flow_rgb(_p(f0), _p(f1), C.c_int(W), C.c_int(H), _p(out))
    return out


RESIZE_INTERP = {'INTER_LINEAR': 0, 'INTER_NEAREST': 1, 'INTER_AREA': 2, 'INTER_CUBIC': 3, 'INTER_LANCZOS4': 4}


def resize(frame, width, height, interpolation='INTER_LINEAR'):
    f = np.ascontiguousarray(frame, np.uint8)
    sh, sw = f.shape[:2]
    cn = f.shape[2] if f.ndim == 3 else 1
    out = np.empty((height, width) + ((cn,) if f.ndim == 3 else ()), np.uint8)
    rc = lib().orc_resize_u8(_p(f), C.c_int(sw), C.c_int(sh), C.c_int(cn), _p(out), C.c_int(width), C.c_int(height),
                             C.c_int(RESIZE_INTERP[interpolation]))
    assert rc == 0
    return out


def rgb2hsv(frame):
    f = np.ascontiguousarray(frame, np.uint8)
    out = np.empty_like(f)
    lib().orc_rgb2hsv_u8(_p(f), C.c_size_t(f.shape[0] * f.shape[1]), _p(out))
    return out
