"""N > 1 host-side logic on CPU: two gloo ranks shard a clip by frame range (with the one-frame /
one-histogram halo), compute their per-frame outputs, concatenate on the host and must
reproduce the single-rank result exactly.  The per-frame compute is injected (numpy stand-ins
here, the CUDA ops on the GPU box: same call shapes), so what is tested is the sharding,
halo and concatenation logic the multi-GPU path uses."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from scannertools_b200 import sharding, shot_detection, synth


def _np_hist(frames):
    return np.stack([np.stack([np.bincount(f[..., c].reshape(-1) >> 4, minlength=16) for c in range(3)])
                     for f in frames]).astype(np.int32) if len(frames) else np.zeros((0, 3, 16), np.int32)


def _np_scores(hist, prev):
    h = hist.astype(np.int64)
    S = np.zeros(len(h), np.int64)
    if len(h) > 1:
        S[1:] = np.abs(h[1:] - h[:-1]).max(axis=2).sum(axis=1)
    if prev is not None and len(h):
        S[0] = np.abs(h[0] - prev.astype(np.int64)).max(axis=1).sum()
    return S.astype(np.int32)


def _pair_feature(f0, f1):
    # stand-in for a per-pair op (OpticalFlow): depends on BOTH frames of the pair
    return np.array([int(f0.astype(np.int64).sum()) - int(f1.astype(np.int64).sum())], np.int64)


def _cuts(n_frames, world):
    # planted cuts, one of them EXACTLY on the shard seam (first frame of rank 1's range)
    seam = sharding.frame_range(n_frames, 1, world)[0]
    return sorted({40, 97, seam, n_frames - 30})


def _worker(rank, world, port, n_frames, out_dir):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    f0, f1 = sharding.frame_range(n_frames, rank, world)
    cuts = _cuts(n_frames, world)
    # each rank generates ONLY its own frames (+ the one halo frame before its range) from the seed
    a0 = max(f0 - 1, 0)
    mine_fr = synth.cut_clip_range(17, n_frames, 18, 32, a0, f1, cuts)
    halo = mine_fr[0:1] if f0 > 0 else None
    bounds, scores = sharding.sharded_shot_detection(mine_fr[f0 - a0:], n_frames, rank, world, _np_hist, _np_scores,
                                                     halo_frame=halo)
    clip = synth.cut_clip_range(17, n_frames, 18, 32, 0, n_frames, cuts)
    (p0, p1), (a, b) = sharding.pair_range(n_frames, rank, world)
    mine = np.stack([_pair_feature(clip[i], clip[i + 1]) for i in range(p0, p1)]) if p1 > p0 else np.zeros((0, 1), np.int64)
    assert (a, b) == ((p0, p1 + 1) if p1 > p0 else (p0, p0))     # one halo frame
    pairs = sharding.gather_frame_outputs(mine, n_frames - 1, rank, world)
    np.savez(os.path.join(out_dir, 'r%d.npz' % rank), bounds=np.array(bounds), scores=scores, pairs=pairs, cuts=np.array(cuts))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_frame_range_sharding_matches_single_rank(tmp_path):
    n = 301   # odd: uneven shards
    mp.spawn(_worker, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
    cuts = _cuts(n, 2)
    clip = synth.cut_clip_range(17, n, 18, 32, 0, n, cuts)
    assert sharding.frame_range(n, 1, 2)[0] in cuts
    ref_scores = _np_scores(_np_hist(clip), None)
    ref_bounds = shot_detection.boundaries_from_scores(ref_scores)
    ref_pairs = np.stack([_pair_feature(clip[i], clip[i + 1]) for i in range(n - 1)])
    for r in range(2):
        d = np.load(os.path.join(str(tmp_path), 'r%d.npz' % r))
        assert np.array_equal(d['scores'], ref_scores)
        assert list(d['bounds']) == ref_bounds == cuts
        assert np.array_equal(d['pairs'], ref_pairs)
