"""Host enqueue time per detect call against the device time of the call (is a small batch host-bound?)."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'feature-point-cnn_b200'))
import torch
import spb200
from spb200 import synth
for (B, H, W, desc) in [(32, 240, 320, False), (32, 240, 320, True), (64, 480, 640, True), (1, 480, 640, True)]:
    e = spb200.Engine(0); e.load_checkpoint(os.path.join(REPO, 'tests/golden/super_point.pt')); e.finalize('fp16')
    e.set_params(descriptor_enabled=desc)
    img = torch.stack([synth.shapes_image(i % 16, H, W) for i in range(B)])[:, None].contiguous().cuda()
    cap = e.max_keypoints(H, W); out = e.alloc_outputs(B, cap, img.device)
    for _ in range(5): e.detect(img, cap, out=out)
    torch.cuda.synchronize()
    n = 100
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); ev0.record()
    for _ in range(n): e.detect(img, cap, out=out)
    ev1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    print('B=%d %dx%d desc=%s: host enqueue %.1f us/call, device %.1f us/call' % (B, H, W, desc, (t1 - t0) / n * 1e6, ev0.elapsed_time(ev1) / n * 1e3))
    e.close()
