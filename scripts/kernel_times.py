"""Per-kernel CUDA-event times of one detect call (batch 64 at 480x640 by default), a few lines: for A/B runs under
environment switches.  Usage: python scripts/kernel_times.py [name-filter] [batch height width]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'feature-point-cnn_b200'))
import torch
import spb200
CKPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'super_point.pt')
flt = sys.argv[1] if len(sys.argv) > 1 else ''
b, h, w = (int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 5 else (64, 480, 640)
g = torch.Generator().manual_seed(1)
imgs = [torch.rand((b, 1, h, w), generator=g).cuda() for _ in range(3)]
e = spb200.Engine(0); e.load_checkpoint(CKPT); e.finalize(os.environ.get('PREC', 'fp16')); e.set_params()
cap = e.max_keypoints(h, w)
out = e.alloc_outputs(b, cap, imgs[0].device)
for i in range(3): e.detect(imgs[i % 3], cap, out=out)
torch.cuda.synchronize()
e.profile_begin()
n = 8
for i in range(n): e.detect(imgs[i % 3], cap, out=out)
ent = e.profile_end()
agg = {}
for name, ms, fl, by in ent: agg[name] = agg.get(name, 0.0) + ms / n
print('total %.4f ms' % sum(agg.values()), ' '.join('%s=%.4f' % (k, v) for k, v in agg.items() if flt in k))
