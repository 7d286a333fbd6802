set -x
python -m pytest tests -m gpu -x -q -k "resize or scanner or kernel_class or histogram_goldens" 2>&1 | tail -5
cat > /tmp/rs.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
from scannertools_b200 import ops
fr = torch.randint(0, 256, (2, 270, 480, 3), dtype=torch.uint8, device='cuda')
for name in ('INTER_CUBIC', 'INTER_LANCZOS4'):
    for (tw, th) in [(213, 120), (700, 301), (1, 1)]:
        ops.resize(fr, width=tw, height=th, interpolation=name)
torch.cuda.synchronize()
print('ok')
PY
compute-sanitizer --tool memcheck python /tmp/rs.py 2>&1 | tail -4
