"""Profiling target for `ncu --profile-from-start off`: one warm pass, then inside the
cudaProfilerStart/Stop window one 1080p Farneback pair (+ flow histogram), one 4K histogram
batch (RGB and fused HSV) and one 1080p flow histogram."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from scannertools_b200 import ops, synth  # noqa: E402


def main():
    what = sys.argv[1:] or ['flow', 'hist', 'flowhist']
    npairs = int(os.environ.get('PROF_PAIRS', '2'))
    W, H = int(os.environ.get('PROF_W', '1920')), int(os.environ.get('PROF_H', '1080'))
    base = synth.textured_clip(1, 5, H, W)
    clip = np.concatenate([base] * ((npairs + 5) // 5 + 1))[:npairs + 1]
    fr = torch.from_numpy(clip).cuda()
    of = ops.OpticalFlow(W, H, max_batch=npairs)
    f4k = torch.randint(0, 256, (8, 2160, 3840, 3), dtype=torch.uint8, device='cuda')
    flow = torch.randn((4, 1080, 1920, 2), device='cuda') * 5

    def work():
        if 'flow' in what:
            of.execute_with_histogram(fr)
        if 'hist' in what:
            ops.shot_scores(ops.histogram(f4k))
            ops.histogram(f4k, hsv='COLOR_RGB2HSV')
        if 'flowhist' in what:
            ops.flow_histogram(flow)
    for _ in range(2):
        work()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    work()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == '__main__':
    main()
