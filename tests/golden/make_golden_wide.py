"""Wide parity set: compact outputs of THE REFERENCE for many images (build container only).

    python tests/golden/make_golden_wide.py

SURVEY.md section 8(d) asks for >= 16 images per family at 240x320 and 480x640 and >= 2 at 1088x1920.  For every
image this script runs the reference's SuperPoint.forward (python/src/superpoint.py:91-115), get_points
(python/src/netutils.py:78-100) and get_descriptors (netutils.py:103-121) and stores, in forward_wide.npz:

    <case>_crc     CRC32 of the input image bytes (the images are regenerated from seeds: spb200/synth.py)
    <case>_xy      keypoints, int16 (N, 2) as (x, y), in the reference's order (descending confidence)
    <case>_conf    their confidences, float32
    <case>_cell    the heatmap's maximum over every 8x8 cell, float32 (H/8, W/8)
    <case>_hsum    sum of the heatmap (float64), <case>_lsum sum of the logits, <case>_dsum sum of the descriptor map
    <case>_desc    descriptors of the first 32 keypoints, float32 (128, 32)

Two checkpoints: the moderate preset (tests/golden/super_point.pt, written by the reference; cases m_*) and the harsh
preset (g = 4, d = 8; the state_dict of oracle.weights.make_state_dict(seed=3, preset='harsh') loaded, strict, into the
reference's SuperPoint module; cases h_*, a subset of the images).
"""
import os
import sys
import zlib

import numpy as np
import torch

if not hasattr(np, 'int'):
    np.int = int

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference/python')
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'feature-point-cnn_b200'))

from src.settings import SuperPointSettings          # noqa: E402
from src.superpoint import SuperPoint                # noqa: E402
from src.saveutils import load_checkpoint_for_inference   # noqa: E402
from src.netutils import get_points, get_descriptors  # noqa: E402

from oracle import weights as ow                     # noqa: E402
from spb200 import synth                             # noqa: E402


def cases():
    """(name, family, index, h, w) of the wide set."""
    out = []
    for h, w, n in ((240, 320, 16), (480, 640, 16), (1088, 1920, 2)):
        for fam in ('shapes', 'rand'):
            for i in range(n):
                out.append(('%s%d_%d' % (fam, h, i), fam, 100 + i, h, w))
    return out


def image_of(fam, i, h, w):
    return synth.shapes_image(i, h, w) if fam == 'shapes' else synth.rand_image(i, h, w)


def main():
    settings = SuperPointSettings()
    nets = {}
    net = SuperPoint(settings)
    assert load_checkpoint_for_inference(os.path.join(HERE, 'super_point.pt'), net)
    nets['m'] = net.eval()
    hard = SuperPoint(settings)
    hard.load_state_dict(ow.make_state_dict(seed=3, preset='harsh'), strict=True)
    nets['h'] = hard.eval()
    out = {}
    with torch.no_grad():
        for tag, net in nets.items():
            for name, fam, i, h, w in cases():
                if tag == 'h' and (h != 240 or i >= 104):
                    if not (h == 480 and i < 102):
                        continue                         # harsh preset: 4 + 4 images at 240x320, 2 + 2 at 480x640
                gray = image_of(fam, i, h, w)
                img = gray[None, None].repeat(1, 3, 1, 1)
                prob, desc, logits = net(img)
                pts = get_points(prob, h, w, settings)
                dsc = get_descriptors(pts, desc, h, w, settings)
                key = '%s_%s' % (tag, name)
                out[key + '_crc'] = np.uint32(zlib.crc32(gray.numpy().tobytes()))
                out[key + '_xy'] = pts[:2].T.astype(np.int16)
                out[key + '_conf'] = pts[2].astype(np.float32)
                out[key + '_cell'] = prob[0].reshape(h // 8, 8, w // 8, 8).amax(dim=(1, 3)).numpy()
                out[key + '_hsum'] = np.float64(prob.double().sum().item())
                out[key + '_lsum'] = np.float64(logits.double().sum().item())
                out[key + '_dsum'] = np.float64(desc.double().sum().item())
                out[key + '_desc'] = dsc[:, :32].astype(np.float32)
                print(key, 'points', pts.shape[1], 'logit max %.1f' % float(logits.max()), flush=True)
    np.savez_compressed(os.path.join(HERE, 'forward_wide.npz'), **out)
    print('wrote', len(out) // 8, 'cases,', os.path.getsize(os.path.join(HERE, 'forward_wide.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
