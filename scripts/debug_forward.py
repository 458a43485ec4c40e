"""Debug helper (GPU box): network-only parity of every precision against the oracle, no NMS."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'feature-point-cnn_b200'))
import numpy as np, torch
import spb200
from oracle import model, weights
ckpt = os.path.join(REPO, 'tests/golden/super_point.pt')
sd = weights.load_state_dict(ckpt)
for name, img in (('shapes', weights.shapes_image(0, 240, 320)), ('rand', weights.rand_image(0, 240, 320))):
    x = img[None, None].contiguous()
    po, do, lo = model.forward(x, sd)
    for prec in sys.argv[1:] or ['fp32', 'fp16', 'bf16']:
        e = spb200.Engine(0); e.load_checkpoint(ckpt); e.finalize(prec); e.set_params()
        p, d, l = e.forward(x.cuda())
        torch.cuda.synchronize()
        print('%s %s: heat %.3e logits %.3e (max %.1f) desc %.3e (max %.1f)' % (name, prec, float((p.cpu()-po).abs().max()),
              float((l.cpu()-lo).abs().max()), float(lo.max()), float((d.cpu()-do).abs().max()), float(do.max())), flush=True)
        e.close()
