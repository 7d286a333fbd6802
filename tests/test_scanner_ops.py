"""The C++ Scanner kernel classes (scannertools_b200/csrc/scanner_ops), built against the
Scanner-API compat shim into libscannertools_imgproc.so, driven the way Scanner's evaluator
drives a kernel: registry lookup by op name, Elements in, execute(), Elements out."""
import ctypes as C

import numpy as np
import pytest

from oracle import restate
from scannertools_b200 import synth


@pytest.fixture(scope='module')
def shim():
    from scannertools_b200 import build
    lib = C.CDLL(build.build_scanner_ops())
    lib.stb_shim_registered.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return lib


def P(a):
    return C.c_void_p(a.ctypes.data)


def test_ops_and_kernels_registered_under_reference_names(shim):
    """REGISTER_OP / REGISTER_KERNEL lines keep the reference's names, batching and stencils
    (histogram_kernel_gpu.cpp:79-82, optical_flow_kernel_cpu.cpp:51-54, optical_flow_kernel_gpu.cpp:109-112,
    flow_histogram_kernel_cpu.cpp:62-67, frame_difference_kernel_cpu.cpp:74-80)."""
    expect = {b'Histogram': (1, 0, 0), b'OpticalFlow': (1, 0, 1), b'FlowHistogram': (1, 0, 0), b'FrameDifference': (0, -1, 0),
              b'Resize': (1, 0, 0), b'ConvertColor': (1, 0, 0), b'ConvertToHSVCPP': (1, 0, 0)}
    for op, (batched, lo, hi) in expect.items():
        b, l, h = C.c_int(-1), C.c_int(99), C.c_int(99)
        assert shim.stb_shim_registered(op, C.byref(b), C.byref(l), C.byref(h)) == 1, op
        assert (b.value, l.value, h.value) == (batched, lo, hi), op
    assert shim.stb_shim_registered(b'NoSuchOp', None, None, None) == 0


@pytest.mark.gpu
def test_histogram_kernel_class(shim):
    clip = synth.noise_clip(3, 9, 90, 160)
    out = np.zeros((9, 48), np.int32)
    assert shim.stb_shim_histogram(P(clip), 9, 160, 90, P(out), 0) == 0
    assert np.array_equal(out, np.stack([restate.histogram(f).reshape(-1) for f in clip]))


@pytest.mark.gpu
def test_optical_flow_and_flow_histogram_kernel_classes(shim):
    h, w = 120, 160
    clip = synth.textured_clip(2, 4, h, w)
    flow = np.zeros((3, h, w, 2), np.float32)
    assert shim.stb_shim_optical_flow(P(clip), 3, w, h, P(flow), 0) == 0
    for i in range(3):
        d = flow[i].astype(np.float64) - restate.optical_flow(clip[i], clip[i + 1])
        e = np.sqrt((d * d).sum(-1))
        assert e.mean() <= 1e-3 and e.max() <= 1e-2, (i, e.max())
    fh = np.zeros((3, 128), np.int32)
    assert shim.stb_shim_flow_histogram(P(flow), 3, w, h, P(fh), 0) == 0
    for i in range(3):
        assert np.array_equal(fh[i].reshape(2, 64), restate.flow_histogram(flow[i]))


@pytest.mark.gpu
def test_frame_difference_kernel_class(shim):
    a = synth.noise_clip(5, 2, 37, 53)
    out = np.zeros_like(a[0])
    assert shim.stb_shim_frame_difference(P(a[0]), P(a[1]), 53, 37, 3, P(out), 0) == 0
    assert np.array_equal(out, restate.frame_difference(a[0], a[1]))


def _resize_args(width=0, height=0, minflag=False, preserve_aspect=False, interpolation=b''):
    """protobuf wire encoding of ResizeArgs (scannertools_imgproc.proto:33-39)."""
    def varint(v):
        out = bytearray()
        while True:
            b = v & 0x7f
            v >>= 7
            out.append(b | (0x80 if v else 0))
            if not v:
                return bytes(out)
    buf = b''
    if width: buf += b'\x08' + varint(width)
    if height: buf += b'\x10' + varint(height)
    if minflag: buf += b'\x18\x01'
    if preserve_aspect: buf += b'\x20\x01'
    if interpolation: buf += b'\x2a' + varint(len(interpolation)) + interpolation
    return buf


@pytest.mark.gpu
def test_resize_kernel_class_with_serialized_args(shim):
    fr = synth.noise_clip(6, 3, 270, 480)
    args = _resize_args(width=213, preserve_aspect=True, interpolation=b'INTER_LINEAR')
    out = np.zeros((3, 200, 300, 3), np.uint8).reshape(-1)
    ow, oh = C.c_int(), C.c_int()
    rc = shim.stb_shim_resize(P(fr), 3, 480, 270, 3, args, len(args), P(out), out.size, C.byref(ow), C.byref(oh), 0)
    assert rc == 0 and (ow.value, oh.value) == (213, 270 * 213 // 480)
    got = out[:3 * oh.value * ow.value * 3].reshape(3, oh.value, ow.value, 3)
    for i in range(3):
        assert np.array_equal(got[i], restate.resize(fr[i], ow.value, oh.value))
    # interpolation names: implemented ones are honoured, names outside the reference's table mean
    # INTER_LINEAR (resize_kernel.cpp:31-35), table names that are not implemented fail validate()
    for name, oracle_name in ((b'INTER_AREA', 'INTER_AREA'), (b'INTER_NEAREST', 'INTER_NEAREST'), (b'INTER_CUBIC', 'INTER_CUBIC'),
                              (b'INTER_LANCZOS4', 'INTER_LANCZOS4'), (b'bogus', 'INTER_LINEAR')):
        args = _resize_args(width=107, height=60, interpolation=name)
        rc = shim.stb_shim_resize(P(fr), 2, 480, 270, 3, args, len(args), P(out), out.size, C.byref(ow), C.byref(oh), 0)
        assert rc == 0 and (ow.value, oh.value) == (107, 60)
        got = out[:2 * 60 * 107 * 3].reshape(2, 60, 107, 3)
        for i in range(2):
            assert np.array_equal(got[i], restate.resize(fr[i], 107, 60, oracle_name)), name
    bad = _resize_args(width=10, height=10, interpolation=b'INTER_MAX')
    assert shim.stb_shim_resize(P(fr), 1, 480, 270, 3, bad, len(bad), P(out), out.size, C.byref(ow), C.byref(oh), 0) == -4


@pytest.mark.gpu
def test_convert_color_kernel_classes(shim):
    fr = synth.noise_clip(8, 2, 45, 64)
    out = np.zeros((2, 45, 64, 3), np.uint8)
    oc = C.c_int()
    name = b'COLOR_RGB2HSV'
    args = b'\x0a' + bytes([len(name)]) + name               # ConvertColorArgs{conversion = 1}
    assert shim.stb_shim_convert_color(P(fr), 2, 64, 45, args, len(args), P(out), C.byref(oc), 0) == 0 and oc.value == 3
    for i in range(2):
        assert np.array_equal(out[i], restate.rgb2hsv(fr[i]))
    out2 = np.zeros_like(out)
    assert shim.stb_shim_convert_color(P(fr), 2, 64, 45, None, -1, P(out2), C.byref(oc), 0) == 0     # ConvertToHSVCPP
    assert np.array_equal(out2, out)
    bad = b'\x0a' + bytes([len(b'COLOR_BGR2XYZ')]) + b'COLOR_BGR2XYZ'
    assert shim.stb_shim_convert_color(P(fr), 1, 64, 45, bad, len(bad), P(out), C.byref(oc), 0) == -4
