"""The frame-range-sharded path on real CUDA devices: two ranks (one process each; on a one-GPU box both
use cuda:0) each generate ONLY their own frames of one seeded clip (+ the one halo frame), run the
CUDA ops on them, and the per-frame outputs concatenated on the host must equal a single process
running the whole clip -- bit for bit, including a hard cut planted exactly on the shard seam and the
flow pair that straddles it (optical_flow_kernel_gpu.cpp:52-57 halo; shot_detection.py:22-26 window
over the whole stream)."""
import os
import socket

import numpy as np
import pytest

from scannertools_b200 import sharding, shot_detection, synth

pytestmark = pytest.mark.gpu

N_FRAMES, H, W = 61, 120, 160      # odd: uneven shards (31 + 30 frames; 30 + 30 pairs)
SEED_CUT, SEED_FLOW = 23, 31


def _cuts(world):
    seam = sharding.frame_range(N_FRAMES, 1, world)[0]
    return sorted({9, seam, 47})


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from scannertools_b200 import ops
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.cuda.set_device(rank % torch.cuda.device_count())
    cuts = _cuts(world)
    # ---- shot detection: frames [f0, f1) + halo frame f0-1
    f0, f1 = sharding.frame_range(N_FRAMES, rank, world)
    a0 = max(f0 - 1, 0)
    fr = torch.from_numpy(synth.cut_clip_range(SEED_CUT, N_FRAMES, H, W, a0, f1, cuts)).cuda()
    bounds, scores = sharding.sharded_shot_detection(
        fr[f0 - a0:], N_FRAMES, rank, world, ops.histogram,
        lambda h, p: ops.shot_scores(h, prev_hist=p).cpu().numpy(), halo_frame=fr[0:1] if f0 > 0 else None)
    # ---- optical flow + flow histogram: pairs [p0, p1) read frames [p0, p1]
    (p0, p1), (fa, fb) = sharding.pair_range(N_FRAMES, rank, world)
    clip = torch.from_numpy(synth.textured_clip(SEED_FLOW, fb - fa, H, W, t0=fa, total=N_FRAMES)).cuda()
    of = ops.OpticalFlow(W, H, max_batch=p1 - p0)
    flow, fh = of.execute_with_histogram(clip)
    of.close()
    fh_all = sharding.gather_frame_outputs(fh.cpu().numpy(), N_FRAMES - 1, rank, world)
    fl_all = sharding.gather_frame_outputs(flow.cpu().numpy(), N_FRAMES - 1, rank, world)
    if rank == 0:
        np.savez(os.path.join(out_dir, 'sharded.npz'), bounds=np.array(bounds), scores=scores, fh=fh_all, flow=fl_all)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_sharded_clip_equals_single_gpu(tmp_path):
    import torch
    import torch.multiprocessing as mp
    from scannertools_b200 import ops
    assert torch.cuda.is_available()
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    d = np.load(os.path.join(str(tmp_path), 'sharded.npz'))
    cuts = _cuts(world)
    # single process, whole clip
    whole = torch.from_numpy(synth.cut_clip_range(SEED_CUT, N_FRAMES, H, W, 0, N_FRAMES, cuts)).cuda()
    S = ops.shot_scores(ops.histogram(whole)).cpu().numpy()
    assert np.array_equal(d['scores'], S)
    assert list(d['bounds']) == shot_detection.boundaries_from_scores(S)
    assert set(d['bounds']) <= set(cuts)
    assert sharding.frame_range(N_FRAMES, 1, world)[0] in d['bounds']    # the cut ON the seam was found
    clip = torch.from_numpy(synth.textured_clip(SEED_FLOW, N_FRAMES, H, W)).cuda()
    of = ops.OpticalFlow(W, H, max_batch=N_FRAMES - 1)
    flow, fh = of.execute_with_histogram(clip)
    of.close()
    assert np.array_equal(d['fh'], fh.cpu().numpy())
    assert np.array_equal(d['flow'], flow.cpu().numpy())                  # incl. the pair that straddles the seam
