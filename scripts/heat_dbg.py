import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'feature-point-cnn_b200'))
os.environ['SPB200_NO_GRAPH'] = '1'
import torch, spb200
from spb200 import synth
B, H, W = 1, 240, 320
e = spb200.Engine(0); e.load_checkpoint(os.path.join(REPO, 'tests/golden/super_point.pt')); e.finalize('fp16'); e.set_params()
img = torch.stack([synth.shapes_image(i, H, W) for i in range(B)])[:, None].contiguous().cuda()
cap = e.max_keypoints(H, W)
count, xy, conf, desc, prob = e.detect(img, cap, want_prob=True)
torch.cuda.synchronize()
print('ok', count)
