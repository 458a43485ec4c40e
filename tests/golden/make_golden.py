"""Generate the golden fixtures by RUNNING THE REFERENCE (build container only).

    python tests/golden/make_golden.py

Imports the reference's own modules from /root/reference/python/src (read-only) and writes, next
to this file:

  super_point.pt    a checkpoint written by the reference's save_checkpoint
                    (python/src/saveutils.py:54-63) from the reference's SuperPoint
                    (python/src/superpoint.py:64-72) with the SURVEY.md section 8(d) weight recipe,
                    moderate preset (g=1, d=2), seed 0
  images.npz        input images: 'shapes' family from the reference's synthetic_shapes
                    (python/src/synthetic_shapes.py, as gen_synthetic_dataset.py:79-101 uses it),
                    stored as uint8 (value/255 is the fp32 input); the 'rand' family is seeded
                    (oracle.weights.rand_image) and not stored
  forward_*.npz     outputs of SuperPoint.forward + get_points + get_descriptors of the reference
                    per image: heatmap, logits / descriptor-map checksums and samples, points,
                    descriptors
  nms_kat.npz       corners_nms (python/src/nms.py:4-53) known answers on random point sets,
                    edge cases (0/1 points, borders, dense all-pass)
  label_kat.npz     make_points_labels -> make_prob_map_from_labels -> get_points round trip
                    (the invariant python/tests/synthetic-test.py:27-28 exercises)

This script is the only file in the repository that touches /root/reference; the fixtures travel
to the GPU box, the reference does not.
"""
import os
import sys

import numpy as np
import torch

if not hasattr(np, 'int'):      # the reference predates numpy 1.24 (python/requirements.txt pins 1.21)
    np.int = int

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference/python')
sys.path.insert(0, REPO)

from src.settings import SuperPointSettings          # noqa: E402
from src.superpoint import SuperPoint                # noqa: E402
from src.saveutils import save_checkpoint, load_checkpoint_for_inference   # noqa: E402
from src.netutils import (get_points, get_descriptors, make_points_labels,  # noqa: E402
                          make_prob_map_from_labels)
from src.nms import corners_nms                      # noqa: E402
from src import synthetic_shapes                     # noqa: E402

from oracle import weights as ow                     # noqa: E402  (rand_image only)


def build_reference_model(g, d):
    torch.manual_seed(0)
    settings = SuperPointSettings()
    net = SuperPoint(settings)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = None
            with torch.no_grad():
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
    last = net.detector.layer[1].bn2
    with torch.no_grad():
        last.weight.fill_(g)
        last.bias.zero_()
        last.bias[64] = d
    net.train()
    with torch.no_grad():
        for _ in range(4):
            net(torch.rand(4, 3, 240, 320))
    net.eval()
    return net, settings


def reference_shapes_image(i, h, w):
    import cv2
    synthetic_shapes.set_random_state(np.random.RandomState(2000 + i))
    img = synthetic_shapes.generate_background((h, w))
    prims = ['draw_checkerboard', 'draw_polygon', 'draw_star', 'draw_cube', 'draw_lines',
             'draw_multiple_polygons']
    getattr(synthetic_shapes, prims[i % len(prims)])(img)
    img = cv2.GaussianBlur(img, (5, 5), 0)
    return img.astype(np.uint8)


def main():
    net, settings = build_reference_model(1.0, 2.0)
    opt = torch.optim.AdamW(net.parameters())
    scaler = torch.amp.GradScaler('cuda', enabled=False)
    save_checkpoint('super_point', 0, net, opt, scaler, '/tmp/golden_ck')
    os.replace('/tmp/golden_ck/super_point_0.pt', os.path.join(HERE, 'super_point.pt'))
    # reload through the reference's own loader to make sure the file round-trips
    net2 = SuperPoint(settings)
    assert load_checkpoint_for_inference(os.path.join(HERE, 'super_point.pt'), net2)
    net2.eval()

    # ---- images ------------------------------------------------------------------------------
    shapes = {}
    for i in range(6):
        shapes['s240_%d' % i] = reference_shapes_image(i, 240, 320)
    for i in range(2):
        shapes['s480_%d' % i] = reference_shapes_image(10 + i, 480, 640)
    np.savez_compressed(os.path.join(HERE, 'images.npz'), **shapes)

    cases = []
    for i in range(3):
        cases.append(('shapes240_%d' % i, torch.from_numpy(shapes['s240_%d' % i].astype(np.float32) / 255.)))
    for i in range(2):
        cases.append(('rand240_%d' % i, ow.rand_image(i, 240, 320)))
    cases.append(('shapes480_0', torch.from_numpy(shapes['s480_0'].astype(np.float32) / 255.)))

    with torch.no_grad():
        for name, gray in cases:
            h, w = gray.shape
            img = gray[None, None].repeat(1, 3, 1, 1)
            prob, desc, logits = net2(img)
            pts = get_points(prob, h, w, settings)
            dsc = get_descriptors(pts, desc, h, w, settings)
            out = dict(points=pts, descriptors=dsc.astype(np.float32),
                       logits_sum=np.float64(logits.double().sum().item()),
                       desc_sum=np.float64(desc.double().sum().item()),
                       logits_sample=logits[0, :, ::7, ::9].numpy(),
                       desc_sample=desc[0, :, ::7, ::9].numpy())
            if h == 240:
                out['heatmap'] = prob[0].numpy()
            else:
                out['heatmap_sample'] = prob[0, ::3, ::3].numpy()
                out['descriptors'] = out['descriptors'][:, :512]
            np.savez_compressed(os.path.join(HERE, 'forward_%s.npz' % name), **out)
            print(name, 'points', pts.shape, 'cand', int((prob >= settings.confidence_thresh).sum()),
                  'logit max', float(logits.max()))

    # ---- NMS known answers -------------------------------------------------------------------
    rs = np.random.RandomState(7)
    kat = {}

    def add(tag, pts, h, w, dist):
        kat[tag + '_in'] = pts
        kat[tag + '_hw'] = np.array([h, w, dist])
        kat[tag + '_out'] = corners_nms(pts.copy(), h, w, dist).astype(np.float64)

    for j, (h, w, n, dist) in enumerate([(40, 56, 300, 4), (64, 64, 1500, 4), (48, 80, 800, 2),
                                         (120, 160, 6000, 4), (32, 32, 200, 1)]):
        idx = rs.choice(h * w, n, replace=False)
        conf = rs.permutation(n).astype(np.float64) / n * 0.9 + 0.05      # distinct values
        add('rand%d' % j, np.stack([idx % w, idx // w, conf]), h, w, dist)
    add('empty', np.zeros((3, 0)), 20, 20, 4)
    add('single', np.array([[3.], [5.], [0.5]]), 20, 20, 4)
    yy, xx = np.mgrid[0:24, 0:32]
    conf = rs.permutation(24 * 32).astype(np.float64) / (24 * 32) * 0.9 + 0.05
    add('dense', np.stack([xx.ravel(), yy.ravel(), conf]), 24, 32, 4)
    ramp = np.stack([np.arange(64), np.full(64, 3), 0.1 + np.arange(64) / 100.])
    add('ramp', ramp.astype(np.float64), 8, 64, 4)
    corners = np.array([[0, 31, 0, 31, 2], [0, 0, 23, 23, 1], [0.9, 0.8, 0.7, 0.6, 0.95]], np.float64)
    add('border', corners, 24, 32, 4)
    np.savez_compressed(os.path.join(HERE, 'nms_kat.npz'), **kat)

    # ---- label round trip --------------------------------------------------------------------
    np.random.seed(3)
    h, w = 64, 96
    cells = rs.choice((h // 8) * (w // 8), 40, replace=False)   # at most one corner per 8x8 cell
    lab_pts = np.stack([(cells // (w // 8)) * 8 + rs.randint(0, 8, 40),
                        (cells % (w // 8)) * 8 + rs.randint(0, 8, 40)], 1).astype(np.float64)  # (y, x)
    labels = make_points_labels(lab_pts, h, w, 8)
    prob = make_prob_map_from_labels(labels, h, w, 8)
    s2 = SuperPointSettings()
    got = get_points(prob, h, w, s2)
    np.savez_compressed(os.path.join(HERE, 'label_kat.npz'), label_points=lab_pts, labels=labels,
                        prob_map=prob[0].numpy().astype(np.float32), points=got)
    print('label KAT: %d label points -> %d recovered' % (len(lab_pts), got.shape[1]))


if __name__ == '__main__':
    main()
