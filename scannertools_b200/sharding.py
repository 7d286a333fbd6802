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
