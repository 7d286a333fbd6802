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
