"""Debug helper (GPU box): NMS + sort alone on random heatmaps of several sizes."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'feature-point-cnn_b200'))
import numpy as np, torch
import spb200
from oracle import postproc
e = spb200.Engine(0)
e.set_params()
for (h, w) in [(16, 32), (64, 96), (48, 320), (240, 320), (480, 640)]:
    g = torch.Generator().manual_seed(1)
    heat = torch.rand((2, h, w), generator=g) ** 4
    try:
        count, xy, conf = e.nms(heat.cuda(), e.max_keypoints(h, w))
        torch.cuda.synchronize()
    except Exception as ex:
        print(h, w, 'FAILED:', str(ex)[:200], flush=True)
        sys.exit(1)
    for i in range(2):
        n = int(count[i]); want = postproc.get_points(heat[i].numpy())
        got = np.zeros((3, n)); got[:2] = xy[i, :n].t().cpu().numpy(); got[2] = conf[i, :n].cpu().numpy()
        sg = {(int(a), int(b)) for a, b in got[:2].T}; sw = {(int(a), int(b)) for a, b in want[:2].T}
        same = got.shape == want.shape and bool((got == want).all())
        print(h, w, 'image', i, 'exact', same, 'n', n, want.shape[1], 'set equal', sg == sw, 'missing', len(sw - sg), 'extra', len(sg - sw),
              'sorted', bool((np.diff(got[2]) <= 0).all()), flush=True)
        if not same:
            print('  got ', got[:, :6].T.tolist()); print('  want', want[:, :6].T.tolist())
            hv = heat[i].numpy()
            bad = [(x, y) for x, y in list(sg - sw)[:4]]
            print('  extras', bad, 'missing', list(sw - sg)[:4])
            print('  conf check', [float(hv[int(y), int(x)]) for x, y in got[:2, :4].T], flush=True)
