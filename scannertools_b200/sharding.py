"""Frame-range sharding across GPUs: frames are independent, so each rank owns
a contiguous range; flow shards read one halo frame past their end; nothing is exchanged
between GPUs.  Per-frame outputs are concatenated on the host in rank order."""


def frame_range(n_frames, rank, world):
    """[start, end) of the frames owned by `rank` (balanced to within one frame)."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pair_range(n_frames, rank, world):
    """Flow pairs (i -> i+1, i in [0, n_frames-1)) owned by `rank`, and the frame range it must
    read: pairs [p0, p1) need frames [p0, p1] (one halo frame)."""
    p0, p1 = frame_range(max(n_frames - 1, 0), rank, world)
    return (p0, p1), (p0, p1 + 1 if p1 > p0 else p0)


def stream_assignment(n_streams, world):
    """C5: whole streams are pinned to GPUs so per-stream previous-frame state never crosses
    devices.  Returns rank -> list of stream ids (round-robin)."""
    return [[s for s in range(n_streams) if s % world == r] for r in range(world)]


def gather_frame_outputs(local, n_frames, rank, world, group=None):
    """Concatenates per-frame outputs (numpy, first axis = this rank's frames, in order) from all
    ranks on every rank, in frame order.  Host-side plumbing over torch.distributed (gloo or
    nccl object collectives); the data path itself has no collective -- this is the 'per-frame
    outputs are concatenated on the host' step."""
    import numpy as np
    import torch.distributed as dist
    if world == 1:
        return np.asarray(local)
    parts = [None] * world
    dist.all_gather_object(parts, np.asarray(local), group=group)
    out = np.concatenate([p for p in parts if len(p)], axis=0) if any(len(p) for p in parts) else np.asarray(local)
    assert out.shape[0] == n_frames, (out.shape, n_frames)
    return out


def sharded_shot_detection(frames_of_rank, n_frames, rank, world, histogram_fn, scores_fn, prev_hist=None, halo_frame=None):
    """Shot detection over a frame-range-sharded clip.

    frames_of_rank : this rank's frames [f0, f1)
    histogram_fn   : frames -> int32 [m, 3, 16]     (ops.histogram on the GPU path)
    scores_fn      : (hist, prev_hist or None) -> int32 [m]   (ops.shot_scores)
    halo_frame     : frame f0-1 as a one-frame batch [1, H, W, 3] (the previous shard's last frame, which
                     this rank reads itself -- nothing is exchanged between GPUs); its histogram is
                     computed here.  Alternatively pass that histogram as prev_hist.
    A shard that does not start the stream (f0 > 0) MUST get one of the two: without it the first
    score of the shard would silently be 0 and a cut on the shard seam would be lost.
    Returns (boundaries, scores[n_frames]) on every rank: the +-500-frame window test spans shard
    boundaries, so it runs once over the concatenated scores (shot_detection.py:22-26)."""
    from . import shot_detection
    f0, f1 = frame_range(n_frames, rank, world)
    if f1 - f0 != len(frames_of_rank):
        raise ValueError('rank %d owns frames [%d, %d) but got %d frames' % (rank, f0, f1, len(frames_of_rank)))
    if f0 > 0 and f1 > f0 and prev_hist is None:
        if halo_frame is None:
            raise ValueError('shard [%d, %d) does not start the stream: pass halo_frame (frame %d) or prev_hist' % (f0, f1, f0 - 1))
        prev_hist = histogram_fn(halo_frame)[0]
    if f0 == 0:
        prev_hist = None         # stream start: diffs[0] = 0 (shot_detection.py:18)
    hist = histogram_fn(frames_of_rank)
    scores = scores_fn(hist, prev_hist)
    all_scores = gather_frame_outputs(scores, n_frames, rank, world)
    return shot_detection.boundaries_from_scores(all_scores), all_scores


def parse_cpulist(text):
    """'0-3,8,10-11' (the sysfs cpulist format) -> [0, 1, 2, 3, 8, 10, 11]."""
    cpus = []
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index, sysfs='/sys'):
    """Pins the calling process to the CPUs of the NUMA node the GPU hangs off, BEFORE it
    allocates its pinned staging buffers, so those pages and the feeder threads sit on the
    socket whose PCIe root complex serves that GPU (the >= 7x target at 8 GPUs 'is
    threatened only by host-side feed: pinned-memory bandwidth, PCIe root-complex sharing, NUMA
    placement of the feeder threads').  Returns the node number, or None when the topology is
    not exposed (virtualised boxes report -1) -- in which case nothing is changed."""
    import os
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open(os.path.join(sysfs, 'bus/pci/devices', bus, 'numa_node')) as f:
            node = int(f.read())
        if node < 0:
            return None
        with open(os.path.join(sysfs, 'devices/system/node/node%d/cpulist' % node)) as f:
            cpus = parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError, AssertionError):
        return None
