"""Small workload for compute-sanitizer: default path, split path, host pipeline (two batches in flight), loaders, matcher."""
import os, sys
REPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(REPO, 'feature-point-cnn_b200'))
import numpy as np, torch, spb200
from spb200 import synth
CKPT = os.path.join(REPO, 'tests', 'golden', 'super_point.pt')
h, w = 112, 144
img = torch.stack([synth.rand_image(1, h, w), synth.shapes_image(0, h, w), synth.shapes_image(1, h, w)])[:, None].contiguous()
for mode in ('fp16', 'fp16+all', 'bf16+layer1'):
    e = spb200.Engine(0); e.load_checkpoint(CKPT); e.finalize(mode); e.set_params()
    cap = e.max_keypoints(h, w)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):                       # third call replays the captured graph
            out = e.detect(img.cuda(), cap)
    torch.cuda.synchronize()
    n = int(out[0].sum())
    e.forward(img.cuda())
    outs = [e.host_outputs(3, cap, True, pinned=True) for _ in range(2)]
    t0 = e.detect_host_submit(img.numpy(), cap, out=outs[0])
    t1 = e.detect_host_submit(img.numpy(), cap, out=outs[1])
    e.detect_host_wait(t0); e.detect_host_wait(t1)
    m = e.match(out[3][:1], out[0][:1], out[3][1:2], out[0][1:2], 0.7)
    g = e.preprocess_u8(torch.randint(0, 256, (2, 150, 200, 3), dtype=torch.uint8, device='cuda'), h, w)
    e.detect_u8(g, cap)
    torch.cuda.synchronize()
    print(mode, 'keypoints', n, 'host', int(outs[0][0].sum()), int(outs[1][0].sum()), flush=True)
    e.close()
print('done')
