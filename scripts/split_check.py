"""Stage-by-stage check of the split-precision levels against the fp32 CUDA-core path on the device (run on the GPU box):
    timeout -s KILL 300 python scripts/split_check.py [height width]
"""
import os, sys
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(REPO, 'feature-point-cnn_b200'))
sys.path.insert(0, REPO)
import torch
import spb200
from spb200 import synth
CKPT = os.path.join(REPO, 'tests', 'golden', 'super_point.pt')
h, w = (int(a) for a in sys.argv[1:3]) if len(sys.argv) > 2 else (240, 320)
img = torch.stack([synth.rand_image(1, h, w), synth.shapes_image(0, h, w)])[:, None].contiguous().cuda()
names = ['pool', 'l1a', 'l1b', 'l2a', 'feat', 'd0', 'logits', 'i0', 'desc']
ref = {}
e = spb200.Engine(0); e.load_checkpoint(CKPT); e.finalize('fp32'); e.set_params()
prob32 = e.forward(img)[0].cpu()
for n in names:
    ref[n] = e.export_activation(n, 2).cpu()
e.close()
for mode in ('fp16', 'fp16+layer1', 'fp16+encoder', 'fp16+all'):
    e = spb200.Engine(0); e.load_checkpoint(CKPT); e.finalize(mode); e.set_params()
    prob = e.forward(img)[0]
    torch.cuda.synchronize()
    line = []
    for n in names:
        a = e.export_activation(n, 2).cpu()
        c = min(a.shape[1], ref[n].shape[1])
        scale = float(ref[n].abs().max()) + 1e-9
        line.append('%s %.1e' % (n, float((a[:, :c] - ref[n][:, :c]).abs().max()) / scale))
    print('%-13s heat %.2e | rel: %s' % (mode, float((prob.cpu() - prob32).abs().max()), '  '.join(line)), flush=True)
    e.close()
