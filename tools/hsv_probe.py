import sys, torch
sys.path.insert(0, '.')
from scannertools_b200 import ops
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3
for kind in ('noise', 'smooth'):
    if kind == 'noise':
        fr = torch.randint(0, 256, (32, 2160, 3840, 3), dtype=torch.uint8, device='cuda')
    else:
        base = torch.linspace(0, 255, 3840, device='cuda').view(1, 1, 3840, 1).expand(32, 2160, 3840, 3)
        fr = (base + torch.randint(-3, 4, (32, 2160, 3840, 3), device='cuda')).clamp(0, 255).to(torch.uint8).contiguous()
    t = timeit(lambda: ops.histogram(fr, hsv='COLOR_RGB2HSV'))
    by = 32 * (3 * 2160 * 3840 + 192)
    print('hsv hist 4K n=32 %s: %.3f ms %.0f fps %.1f%% of measured HBM' % (kind, t * 1e3, 32 / t, 100 * by / t / 6514.2e9))
