"""Host-side logic of the product that needs no GPU: ShotBoundaries windowed test, wire-format
readers, frame-range sharding."""
import numpy as np

from scannertools_b200 import sharding, shot_detection, types


def test_shot_boundaries_matches_reference_logic(golden):
    g = golden('shot_c1.npz')
    hists = g['hists'].reshape(-1, 3, 16)
    elements = [types.histograms(types.histogram_bytes(h)) for h in hists]
    rows = shot_detection.shot_boundaries(None, elements)
    assert len(rows) == len(hists)
    assert rows[0] == list(g['boundaries']) and len(rows[0]) == 7      # tests/test_all.py:233
    assert all(r is None for r in rows[1:])                             # shot_detection.py:28
    # device-score entry point gives the same answer
    assert shot_detection.shot_boundaries(None, scores=g['scores'])[0] == list(g['boundaries'])
    assert np.array_equal(shot_detection.scores_from_histograms(elements), g['scores'])


def test_shot_boundaries_edge_cases():
    assert shot_detection.shot_boundaries(None, scores=np.zeros(0, np.int32)) == []
    assert shot_detection.shot_boundaries(None, scores=np.zeros(1, np.int32)) == [[]]
    # constant clip: std = 0 and diff - mean = 0 -> strict '>' finds nothing
    assert shot_detection.shot_boundaries(None, scores=np.zeros(50, np.int32))[0] == []
    s = np.zeros(1200, np.int32)
    s[[100, 700, 1199]] = 9000
    assert shot_detection.shot_boundaries(None, scores=s)[0] == [100, 700, 1199]


def test_readers_roundtrip():
    h = np.arange(48, dtype=np.int32).reshape(3, 16)
    r = types.histograms(types.histogram_bytes(h))
    assert len(r) == 3 and np.array_equal(np.stack(r), h)
    assert types.histograms(None) is None and types.flow_hist_reader(None) is None
    fh = np.arange(128, dtype=np.int32)
    m, a = types.flow_hist_reader(fh.tobytes())
    assert np.array_equal(m, fh[:64]) and np.array_equal(a, fh[64:])
    f = np.arange(2 * 3 * 2, dtype=np.float32)
    assert types.flow(f.tobytes(), 2, 3).shape == (2, 3, 2)


def test_frame_ranges_cover_and_halo():
    for n in (0, 1, 7, 1000, 10001):
        for world in (1, 2, 3, 8):
            rs = [sharding.frame_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
            pairs = [sharding.pair_range(n, r, world) for r in range(world)]
            covered = sum(p1 - p0 for (p0, p1), _ in pairs)
            assert covered == max(n - 1, 0)
            for (p0, p1), (f0, f1) in pairs:
                if p1 > p0:
                    assert (f0, f1) == (p0, p1 + 1)   # one halo frame
    assert sharding.stream_assignment(64, 8)[3] == [3, 11, 19, 27, 35, 43, 51, 59]


def test_numa_binding_helpers(tmp_path):
    from scannertools_b200 import sharding
    assert sharding.parse_cpulist('0-3,8,10-11\n') == [0, 1, 2, 3, 8, 10, 11]
    assert sharding.parse_cpulist('') == []
    # no CUDA device / no topology: nothing changes and None comes back
    import os
    before = os.sched_getaffinity(0)
    assert sharding.bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) is None
    assert os.sched_getaffinity(0) == before


def test_scanner_registration_shim_registers_shot_boundaries(golden):
    """scannertools_b200.scanner_register mirrors `@scannerpy.register_python_op(name='ShotBoundaries',
    batch=BOUNDARY_BATCH)` (shot_detection.py:11) and the imgproc library registration
    (imgproc/__init__.py:1-3).  scannerpy is not in this image: a recording stub stands in."""
    import importlib
    import sys
    import types as pytypes
    calls = {}
    sp = pytypes.ModuleType('scannerpy')

    def register_python_op(**kwargs):
        def deco(fn):
            calls['op'] = (kwargs, fn)
            return fn
        return deco
    sp.register_python_op = register_python_op
    st = pytypes.ModuleType('scannerpy.types')
    st.Histogram = object
    op = pytypes.ModuleType('scannerpy.op')
    op.register_module = lambda so, proto=None: calls.setdefault('module', (so, proto))
    sp.op = op
    saved = {k: sys.modules.get(k) for k in ('scannerpy', 'scannerpy.types', 'scannerpy.op')}
    sys.modules.update({'scannerpy': sp, 'scannerpy.types': st, 'scannerpy.op': op})
    try:
        sys.modules.pop('scannertools_b200.scanner_register', None)
        mod = importlib.import_module('scannertools_b200.scanner_register')
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        sys.modules.pop('scannertools_b200.scanner_register', None)
    kwargs, fn = calls['op']
    assert kwargs == {'name': 'ShotBoundaries', 'batch': 10000000}
    assert calls['module'][0].endswith('libscannertools_imgproc.so')
    assert mod.WINDOW_SIZE == 500
    g = golden('shot_c1.npz')
    elements = [types.histograms(types.histogram_bytes(h)) for h in g['hists'].reshape(-1, 3, 16)]
    rows = fn(None, elements)
    assert rows[0] == list(g['boundaries']) and all(r is None for r in rows[1:])


def test_scanner_registration_needs_scannerpy():
    import importlib
    import sys
    import pytest
    if 'scannerpy' in sys.modules:
        pytest.skip('a scannerpy (stub) is loaded')
    sys.modules.pop('scannertools_b200.scanner_register', None)
    with pytest.raises(ImportError):
        importlib.import_module('scannertools_b200.scanner_register')


def test_synthetic_clips_are_shardable_by_frame_range():
    from scannertools_b200 import synth
    whole = synth.textured_clip(5, 40, 45, 80)
    part = synth.textured_clip(5, 7, 45, 80, t0=20, total=40)
    assert np.array_equal(whole[20:27], part)
    cuts = [9, 16, 20]
    a = synth.cut_clip_range(3, 32, 36, 64, 0, 32, cuts)
    b = synth.cut_clip_range(3, 32, 36, 64, 15, 21, cuts)
    assert np.array_equal(a[15:21], b)
    # a cut changes the shot's base image: large frame difference exactly at the cut
    d = np.abs(a[1:].astype(np.int32) - a[:-1]).mean(axis=(1, 2, 3))
    assert set(np.nonzero(d > 8)[0] + 1) == set(cuts)


def test_sharded_shot_detection_requires_the_halo():
    import pytest
    fr = np.zeros((4, 2, 2, 3), np.uint8)
    hist = lambda f: np.zeros((len(f), 3, 16), np.int32)
    sc = lambda h, p: np.zeros(len(h), np.int32)
    with pytest.raises(ValueError):      # rank 1 of 2 without halo frame / prev_hist: a seam cut would be lost silently
        sharding.sharded_shot_detection(fr, 8, 1, 2, hist, sc)
    with pytest.raises(ValueError):      # wrong number of frames for the shard
        sharding.sharded_shot_detection(fr[:3], 8, 0, 2, hist, sc)
