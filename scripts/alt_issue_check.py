"""Stress check of the alternating MMA issue (halo_tc.cu): N repetitions of detect + forward on the bench batch against one run
with SPB200_NO_ALT_ISSUE=1 - every output of every repetition must be bit-identical (the order of the MMAs on an accumulator is
the step-list order in both forms).  Usage: python scripts/alt_issue_check.py [reps] [precision]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'feature-point-cnn_b200'))
import torch
import spb200
from spb200 import synth
CKPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'super_point.pt')
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
prec = sys.argv[2] if len(sys.argv) > 2 else 'fp16'
B, H, W = 64, 480, 640
imgs = [torch.stack([synth.shapes_image(100 * k + i, H, W) for i in range(16)]).repeat(4, 1, 1)[:, None].contiguous().cuda() for k in range(2)]

def run(no_alt, n):
    os.environ['SPB200_NO_ALT_ISSUE'] = '1' if no_alt else '0'
    e = spb200.Engine(0); e.load_checkpoint(CKPT); e.finalize(prec); e.set_params()
    cap = e.max_keypoints(H, W)
    res = []
    for r in range(n):
        img = imgs[r % 2]
        count, xy, conf, desc = e.detect(img, cap)[:4]
        torch.cuda.synchronize()
        c = count.cpu()
        # compact checksums of the valid rows (exact: integer views)
        m = (torch.arange(cap, device='cuda')[None, :] < count[:, None])
        res.append((c.clone(), int((xy.long() * m[..., None]).sum()), int((conf.view(torch.int32).long() * m).sum()),
                    int((desc.view(torch.int32).long() * m[..., None]).sum())))
    prob, d, logits = e.forward(imgs[0][:8])
    res.append((int(prob.view(torch.int32).long().sum()), int(d.view(torch.int32).long().sum()), int(logits.view(torch.int32).long().sum())))
    e.close()
    return res

ref = run(True, 2)
got = run(False, reps)
bad = 0
for r in range(reps):
    a, b = ref[r % 2], got[r]
    if not (torch.equal(a[0], b[0]) and a[1:] == b[1:]): bad += 1
if ref[-1] != got[-1]: bad += 1
print('alternating issue, %s: %d repetitions of batch %d at %dx%d, %d keypoints per batch: %d mismatches' % (prec, reps, B, H, W, int(ref[0][0].sum()), bad))
sys.exit(1 if bad else 0)
