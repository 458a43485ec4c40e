"""Small fixed workload for ncu: N detect steps on the bench workload (batch 64, 480x640, fp16)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'feature-point-cnn_b200'))
import torch
import spb200
from spb200 import synth
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, H, W = 64, 480, 640
e = spb200.Engine(0)
e.load_checkpoint(os.path.join(REPO, 'tests/golden/super_point.pt'))
e.finalize(sys.argv[2] if len(sys.argv) > 2 else 'fp16')
e.set_params()
base = torch.stack([synth.shapes_image(i, H, W) for i in range(16)])
img = base.repeat(4, 1, 1)[:, None].contiguous().cuda()
cap = e.max_keypoints(H, W)
out = e.alloc_outputs(B, cap, img.device)
for _ in range(steps):
    e.detect(img, cap, out=out)
torch.cuda.synchronize()
print('ok', int(out[0].sum()))
