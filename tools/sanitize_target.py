"""Target for compute-sanitizer (memcheck / racecheck / synccheck): one small invocation of every kernel family --
RGB histogram (joint tables, ragged and unaligned frames), shot scores, FlowHistogram, FrameDifference, fused HSV
histogram, Resize (all modes), OpticalFlow with the fused histogram at a size where the TMA + window kernels run
on the fine levels and the fall-backs on the coarse ones."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from scannertools_b200 import ops, synth  # noqa: E402


def main():
    fr = torch.from_numpy(synth.noise_clip(1, 5, 97, 131)).cuda()
    h = ops.histogram(fr)
    ops.shot_scores(h)
    ops.histogram(fr, hsv='COLOR_RGB2HSV')
    buf = torch.zeros(97 * 131 * 3 + 7, dtype=torch.uint8, device='cuda')
    buf[3:3 + 97 * 131 * 3] = fr[0].reshape(-1)
    ops.histogram(buf[3:3 + 97 * 131 * 3].reshape(1, 97, 131, 3))
    ops.frame_difference(fr[0], fr[1]) if hasattr(ops, 'frame_difference') else None
    ops.flow_histogram(torch.randn((3, 97, 131, 2), device='cuda') * 7)
    for name in ('INTER_LINEAR', 'INTER_NEAREST', 'INTER_AREA', 'INTER_CUBIC', 'INTER_LANCZOS4'):
        ops.resize(fr[:2], width=61, height=40, interpolation=name)
        ops.resize(fr[:1], width=200, height=150, interpolation=name)
    for (hh, ww, gen) in [(272, 480, synth.textured_clip), (120, 164, synth.warped_clip), (76, 112, synth.noise_clip)]:
        clip = torch.from_numpy(gen(3, 4, hh, ww)).cuda()
        of = ops.OpticalFlow(ww, hh, max_batch=3)
        of.execute_with_histogram(clip)
        of.execute_with_histogram(clip, want_flow=False)
        of.close()
    torch.cuda.synchronize()
    print('sanitize target ok')


if __name__ == '__main__':
    main()
