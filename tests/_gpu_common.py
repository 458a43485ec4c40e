"""Shared helpers of the GPU parity tests."""
import os

import numpy as np
import torch

from oracle import model, postproc, weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CKPT = os.path.join(GOLDEN, 'super_point.pt')
GOLDEN_CASES = ['shapes240_0', 'shapes240_1', 'shapes240_2', 'rand240_0', 'rand240_1', 'shapes480_0']


def load_spb():
    import spb200
    return spb200


class LazyEngines(dict):
    """precision -> engine holding the golden checkpoint, created on first use."""

    def __missing__(self, prec):
        e = load_spb().Engine(0)
        e.load_checkpoint(CKPT)
        e.finalize(prec)
        e.set_params()
        self[prec] = e
        return e

    def close(self):
        for e in self.values():
            e.close()


def golden_image(name):
    imgs = np.load(os.path.join(GOLDEN, 'images.npz'))
    fam, idx = name.rsplit('_', 1)
    if fam.startswith('shapes'):
        return torch.from_numpy(imgs['s%s_%s' % (fam[len('shapes'):], idx)].astype(np.float32) / 255.)
    h = int(fam[len('rand'):])
    return weights.rand_image(int(idx), h, h * 4 // 3)


def pset(pts):
    return {(int(x), int(y)) for x, y in zip(pts[0], pts[1])}


def points_from(count, xy, conf, i=0):
    n = int(count[i])
    p = np.zeros((3, n))
    p[:2] = xy[i, :n].t().cpu().numpy()
    p[2] = conf[i, :n].cpu().numpy()
    return p


def compare_path(e, gray, sd, heat_tol, kp_frac, cos_min, tag):
    h, w = gray.shape
    img = gray[None, None].contiguous()
    prob_o, desc_o, logits_o = model.forward(img, sd)
    prob, desc, logits = e.forward(img.cuda())
    dh = float((prob.cpu() - prob_o).abs().max())
    pts_o = postproc.get_points(prob_o.numpy())
    dsc_o = postproc.get_descriptors(pts_o, desc_o.numpy(), h, w)
    cap = e.max_keypoints(h, w)
    count, xy, conf, dsc, _ = e.detect(img.cuda(), cap)
    pts = points_from(count, xy, conf)
    inter = len(pset(pts) & pset(pts_o))
    frac = inter / max(len(pset(pts_o)), 1)
    # descriptor cosine at the oracle's keypoints (sample our map at the oracle's points)
    n = pts_o.shape[1]
    xyo = torch.from_numpy(np.ascontiguousarray(pts_o[:2].T.astype(np.int32)))[None].cuda()
    mine = e.sample_descriptors(desc, h, w, torch.tensor([n], dtype=torch.int32, device='cuda'), xyo)[0, :n].t().cpu().numpy()
    cos = float((mine * dsc_o).sum(0).min()) if n else 1.0
    print('[parity %s] heat max-abs %.3e  keypoints %d/%d (%.4f)  desc cos min %.6f  logits max-abs %.3e' %
          (tag, dh, inter, len(pset(pts_o)), frac, cos, float((logits.cpu() - logits_o).abs().max())))
    assert dh <= heat_tol, (tag, dh)
    assert frac >= kp_frac, (tag, frac)
    assert cos >= cos_min, (tag, cos)
    # the fused detect path returns unit-norm descriptors of its own keypoints
    if int(count[0]):
        nrm = dsc[0, :int(count[0])].norm(dim=1)
        assert float((nrm - 1).abs().max()) < 1e-4
    return dh, frac, cos, inter, len(pset(pts_o))




def unexplained_differences(heat_a, heat_b, thresh=0.015, radius=4):
    """Keypoint-set differences between two heatmaps that are NOT ties or threshold-edge cases (the only differences the
    north star allows).

    The greedy NMS (python/src/nms.py:4-53) is run on both maps WITHOUT border removal (a suppressor inside the border is
    invisible in the final lists).  A point kept on one map and not on the other is either a candidate on one map only
    (threshold edge), or was suppressed by a point that is itself kept on one map only: the differing points form chains
    of window neighbours, and a chain ends at a threshold-edge point or at two neighbours whose ORDER differs between the
    maps (a near tie: |difference| <= 2 max|heat_a - heat_b|).  Returns the number of connected components (Chebyshev
    distance <= radius) of the symmetric difference without such a root cause, and the size of the symmetric difference.
    """
    heat_a, heat_b = np.asarray(heat_a, dtype=np.float32), np.asarray(heat_b, dtype=np.float32)
    a = pset(postproc.get_points(heat_a, thresh, radius, 0))
    b = pset(postproc.get_points(heat_b, thresh, radius, 0))
    diff = sorted(a ^ b)
    if not diff:
        return 0, 0
    parent = list(range(len(diff)))

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    for i in range(len(diff)):
        for j in range(i + 1, len(diff)):
            if abs(diff[i][0] - diff[j][0]) <= radius and abs(diff[i][1] - diff[j][1]) <= radius:
                parent[find(i)] = find(j)
    comps = {}
    for i in range(len(diff)):
        comps.setdefault(find(i), []).append(diff[i])
    bad = 0
    for pts in comps.values():
        ok = False
        for (x, y) in pts:
            if (heat_a[y, x] >= thresh) != (heat_b[y, x] >= thresh):
                ok = True
        for i in range(len(pts)):
            for j in range(i + 1, len(pts)):
                (x0, y0), (x1, y1) = pts[i], pts[j]
                if abs(x0 - x1) <= radius and abs(y0 - y1) <= radius:
                    da = float(heat_a[y0, x0]) - float(heat_a[y1, x1])
                    db = float(heat_b[y0, x0]) - float(heat_b[y1, x1])
                    if da * db <= 0:
                        ok = True
        bad += 0 if ok else 1
    return bad, len(diff)
