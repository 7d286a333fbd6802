import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from scannertools_b200 import ops
from oracle import cv2_ops, restate
h, w = 240, 320
rng = np.random.default_rng(0)
noise = rng.integers(0, 256, size=(2, h, w, 3), dtype=np.uint8)
same = np.stack([noise[0], noise[0]])
of = ops.OpticalFlow(w, h, max_batch=1)
def epe(a, b): return np.sqrt(((a - b) ** 2).sum(-1))
for tag, clip in (('noise', noise), ('same', same)):
    outs = [of.execute(torch.from_numpy(clip).cuda()).cpu().numpy()[0] for _ in range(3)]
    cv = cv2_ops.optical_flow(clip[0], clip[1]); rs = restate.optical_flow(clip[0], clip[1])
    e = epe(outs[0], cv)
    y, x = np.unravel_index(e.argmax(), e.shape)
    print(tag, 'gpu-vs-cv2 max %.3e mean %.3e at' % (e.max(), e.mean()), (y, x), 'gpu', outs[0][y, x], 'cv', cv[y, x], 'rs', rs[y, x],
          'restate-vs-cv2 %.3e' % epe(rs, cv).max(), 'repeatable', all(np.array_equal(outs[0], o) for o in outs[1:]), 'n>1e-3:', int((e > 1e-3).sum()))
    for k in range(len(of.levels()) - 1, -1, -1):
        _, d = of.debug_level(torch.from_numpy(clip).cuda(), k)
        _, r = restate.farneback(restate.gray(clip[0]), restate.gray(clip[1]), dump_level=k)
        ef = epe(d['flow'].cpu().numpy(), r['flow'])
        print('   level', k, 'I %.2e R0 %.2e R1 %.2e M0 %.2e flow epe max %.2e' % (
            np.abs(d['I0'].cpu().numpy() - r['I0']).max(), np.abs(d['R0'].cpu().numpy() - r['R0'].transpose(2, 0, 1)).max(),
            np.abs(d['R1'].cpu().numpy() - r['R1'].transpose(2, 0, 1)).max(), np.abs(d['M0'].cpu().numpy() - r['M0'].transpose(2, 0, 1)).max(), ef.max()),
            'at', np.unravel_index(ef.argmax(), ef.shape))
