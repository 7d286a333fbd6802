"""Developer timing probe (not the contract bench): whole-call CUDA-event timings of each op
on device-resident synthetic frames."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from scannertools_b200 import ops, synth  # noqa: E402

PEAK = 6514.2e9


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def main():
    which = sys.argv[1:] or ['hist', 'flow', 'flowhist']
    print(torch.cuda.get_device_name(0))
    if 'hist' in which:
        for (h, w, n, kind) in [(2160, 3840, 32, 'noise'), (2160, 3840, 32, 'const'), (360, 640, 1000, 'noise'), (1080, 1920, 64, 'noise')]:
            if kind == 'noise':
                fr = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device='cuda')
            else:
                fr = torch.full((n, h, w, 3), 77, dtype=torch.uint8, device='cuda')
            t = timeit(lambda: ops.histogram(fr))
            by = n * (3 * h * w + 192)
            print('hist %dx%d n=%d %s: %.3f ms  %.0f fps  %.1f GB/s  %.1f%% of measured HBM' % (w, h, n, kind, t * 1e3, n / t, by / t / 1e9, 100 * by / t / PEAK))
    if 'framediff' in which:
        # FrameDifference over a whole batch in one call: frames [1..n) against [0..n-1) (9WH bytes per frame: 2 reads + 1 write of 3WH)
        for (h, w, n) in [(2160, 3840, 16), (1080, 1920, 64)]:
            fr = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device='cuda')
            t = timeit(lambda: ops.frame_difference(fr[:-1], fr[1:]))
            by = (n - 1) * 9 * h * w
            print('framediff %dx%d n=%d: %.3f ms %.0f fps %.1f GB/s %.1f%%' % (w, h, n - 1, t * 1e3, (n - 1) / t, by / t / 1e9, 100 * by / t / PEAK))
    if 'flowhist' in which:
        f = torch.randn((16, 1080, 1920, 2), device='cuda') * 5
        t = timeit(lambda: ops.flow_histogram(f))
        by = 16 * (8 * 1080 * 1920 + 512)
        print('flowhist 1080p n=16: %.3f ms %.0f fps %.1f GB/s %.1f%%' % (t * 1e3, 16 / t, by / t / 1e9, 100 * by / t / PEAK))
    if 'flow' in which:
        for (h, w, n, bytes_per) in [(480, 640, 16, 114336000), (480, 640, 64, 114336000), (720, 1280, 8, 343008000), (720, 1280, 32, 343008000), (1080, 1920, 8, 771768000), (1080, 1920, 16, 771768000)]:
            base = synth.textured_clip(1, 4, h, w)
            clip = np.concatenate([base] * ((n + 1 + 3) // 4))[:n + 1]
            fr = torch.from_numpy(clip).cuda()
            of = ops.OpticalFlow(w, h, max_batch=n)
            out = torch.empty((n, h, w, 2), dtype=torch.float32, device='cuda')
            t = timeit(lambda: of.execute(fr, out=out), iters=5, warm=2)
            print('flow %dx%d n=%d: %.3f ms/batch  %.1f us/frame  %.0f fps  %.1f GB/s canonical  %.1f%% of measured HBM' % (
                w, h, n, t * 1e3, t / n * 1e6, n / t, n * bytes_per / t / 1e9, 100 * n * bytes_per / t / PEAK))
            of.close()


if __name__ == '__main__':
    main()
